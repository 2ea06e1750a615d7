#!/usr/bin/env python
"""Beyond the reference's examples: a field that changes between queries, and particles pushed through it.

The reference warns that its lazily filled coefficient cache is "not a good idea ... where you are frequently updating
the field" (CHANGELOG.md:9).  Here a 128^3 table is rebuilt in place in well under a millisecond, so the usual loop

    for every time step:  field values change  ->  interpolate at the particle positions  ->  move the particles

keeps one interpolator alive: `update_values` swaps the values, `push` advances the particles by several
velocity-Verlet steps inside one kernel (dv/dt = kappa * grad|B|).  Needs a B200; prints a few numbers."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic  # noqa: E402


def quadrupole(X, Y, Z, t):
    g, b0 = 1.0 + 0.2 * np.sin(t), 0.3
    return np.stack([g * X, g * Y, -2.0 * g * Z + b0 * np.cos(t)], axis=1)


def main(n=128, particles=200_000, frames=5):
    ax = np.linspace(-1.0, 1.0, n)
    Z, Y, X = [a.ravel() for a in np.meshgrid(ax, ax, ax, indexing="ij")]
    field = np.concatenate([np.stack([X, Y, Z], axis=1), quadrupole(X, Y, Z, 0.0)], axis=1)
    trap = tricubic(field, "quiet", mode="norm")                      # |B| and its gradient
    rng = np.random.default_rng(0)
    pos = torch.from_numpy(rng.normal(0.0, 0.15, (particles, 3))).cuda()
    vel = torch.from_numpy(rng.normal(0.0, 0.05, (particles, 3))).cuda()
    kappa, dt, substeps = -0.5, 0.01, 20
    for frame in range(1, frames + 1):
        t = frame * dt * substeps
        t0 = time.perf_counter()
        trap.update_values(quadrupole(X, Y, Z, t), order="grid")     # rows were generated in grid order
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        lost = trap.push(pos, vel, dt, substeps, kappa)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        bn, _ = trap.Query(pos)                                       # CUDA tensors in, CUDA tensors out
        print(f"t = {t:.2f}: field update {1e3 * (t1 - t0):.1f} ms (host -> table), {substeps} push steps "
              f"{1e3 * (t2 - t1):.2f} ms, lost {lost}, mean |B| at the particles {float(torch.nanmean(bn)):.4f}")


if __name__ == "__main__":
    main()
