#!/usr/bin/env python
"""Synthetic stand-ins for the reference's example fields (examples/ExampleFields.zip is not part of the
reference checkout): mm-scale coordinates so that the example scripts' query line
``linspace(-2e-3, 2e-3, 20)`` (and ``linspace(-3e-6, 3e-6, 20)`` in time) lies inside the interpolation
volume; a quadrupole-like magnetic field plus bias, written as comma-separated ``%.18e`` rows."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200.io import save_field_csv  # noqa: E402


def field3(n=21):
    ax = np.linspace(-3e-3, 3e-3, n)
    Z, Y, X = [a.ravel() for a in np.meshgrid(ax, ax, ax, indexing="ij")]
    g, b0 = 12.0, 1e-3                                    # T/m gradient, 1 mT bias
    B = np.stack([g * X + 0.3 * b0, g * Y, -2 * g * Z + b0], axis=1)
    return np.column_stack([X, Y, Z, B])


def field4(n=13, nt=9):
    ax = np.linspace(-3e-3, 3e-3, n); t = np.linspace(-5e-6, 5e-6, nt)
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(t, ax, ax, ax, indexing="ij")]
    g, b0, w = 12.0, 1e-3, 2 * np.pi * 4e4
    s = 1 + 0.3 * np.sin(w * T)
    B = np.stack([(g * X + 0.3 * b0) * s, g * Y * s, (-2 * g * Z + b0) * s], axis=1)
    return np.column_stack([X, Y, Z, T, B])


def main(out="ExampleFields"):
    os.makedirs(out, exist_ok=True)
    f3, f4 = field3(), field4()
    save_field_csv(os.path.join(out, "Example3DVectorField.csv"), f3)
    save_field_csv(os.path.join(out, "Example3DScalarField.csv"), np.column_stack([f3[:, :3], np.linalg.norm(f3[:, 3:], axis=1)]))
    save_field_csv(os.path.join(out, "Example4DVectorField.csv"), f4)
    save_field_csv(os.path.join(out, "Example4DScalarField.csv"), np.column_stack([f4[:, :4], np.linalg.norm(f4[:, 4:], axis=1)]))
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main(*sys.argv[1:])
