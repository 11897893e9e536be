"""Wave-quantisation probe: per-kernel time of the 4096^2 SurfaceQG ETDRK4 step as the number of live column
pairs (grid size of the two column kernels) is varied through aliased_fraction.  296 = 148 SMs x 2 resident CTAs."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfb_loader
gf = gfb_loader.load()
from bench import initial_bh

n = 4096
bh = initial_bh(n)
for kcut in (1184, 1366, 1480, 1776):
    a = 1 / 3 if kcut == 1366 else 1 - 2 * kcut / n
    prob = gf.SurfaceQG.Problem(gf.GPU(), nx=n, Lx=2 * np.pi, nu=1e-19 * (256 / n) ** 8, nnu=4, dt=0.03 * 256 / n,
                                stepper="ETDRK4", aliased_fraction=a)
    prob.sol.copy_(torch.from_numpy(bh))
    prob.step_lib(5)
    prob.profile_begin()
    prob.step_lib(20)
    pr = prob.profile_end()
    print(json.dumps({"kcut": int(prob.grid.kr_alias0), "pairs": int(prob.grid.kr_alias0) // 2,
                      "waves": round(int(prob.grid.kr_alias0) / 2 / 296, 2),
                      **{k: round(v["ms"] / max(v["launches"], 1), 4) for k, v in pr.items()}}))
    del prob
