"""Micro driver used for the ncu capture of profiles/r1_mmd_kernel.md (MMDb of 1e4 x 1e4 x 22 float64 rows)."""
import sys, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nfisam_b200.utils import MMDb
m, d = 10000, 22
xa = torch.randn((m, d), dtype=torch.float64, device='cuda'); xb = torch.randn((m, d), dtype=torch.float64, device='cuda') + .1
for _ in range(3): print(MMDb(xa, xb, float(np.sqrt(d))))
