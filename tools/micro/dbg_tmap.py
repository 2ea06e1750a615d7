import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["ARB_DEBUG_TMAP"] = "1"
from arbinterp_b200 import tricubic
n = 64
ax = torch.linspace(0, 1, n, dtype=torch.float64, device="cuda")
Z, Y, X = torch.meshgrid(ax, ax, ax, indexing="ij")
f = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1), (X + Y + Z).reshape(-1)], 1)
o = tricubic(f, "quiet")
print("lib build ok", o.nc)
