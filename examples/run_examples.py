#!/usr/bin/env python
"""The flow of the reference's two example scripts (examples/3D_ARBInterpExample.py,
examples/4D_ARBInterpExample.py) against the drop-in import path: load a CSV field, build the
interpolator, query a single point and a 20-point line, first on the scalar field, then on the vector
field with mode='both'.  Needs a GPU.  Fields: run examples/make_example_fields.py first."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ARBTools.ARBInterp import quadcubic, tricubic  # noqa: E402  (the reference examples' import line)
from arbinterp_b200.io import load_field_csv  # noqa: E402


def run(cls, d, folder, stem):
    coords = np.zeros((20, d))
    for a in range(3):
        coords[:, a] = np.linspace(-2e-3, 2e-3, 20)
    if d == 4:
        coords[:, 3] = np.linspace(-3e-6, 3e-6, 20)
    out = {}
    print("--- Loading field ---")
    field = load_field_csv(os.path.join(folder, f"Example{stem}ScalarField.csv"))
    run_ = cls(field)                       # mode kword arg is ignored for a scalar input
    out["scalar_single"] = run_.Query(coords[3])
    print("\nSingle point query, scalar field:\n", out["scalar_single"])
    out["scalar_line"] = run_.Query(coords.copy())
    print("\nMulti point query, scalar field:\n", out["scalar_line"])
    print("--- Loading field ---")
    field = load_field_csv(os.path.join(folder, f"Example{stem}VectorField.csv"))
    run_ = cls(field, mode="both")
    out["vector_single"] = run_.Query(coords[3])
    print("Single point query, vector field:\n", out["vector_single"])
    out["vector_line"] = run_.Query(coords.copy())
    print("Multi point query, vector field:\n", out["vector_line"])
    return out


if __name__ == "__main__":
    folder = sys.argv[1] if len(sys.argv) > 1 else "ExampleFields"
    run(tricubic, 3, folder, "3D")
    run(quadcubic, 4, folder, "4D")
