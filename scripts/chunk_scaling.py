"""4096 chains at L = 32 as back-to-back launches of w waves each (development aid): the per-wave time of the resident-chain
kernel grows with the length of a launch (scripts/wave_scaling.py), so shorter launches are faster in total."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
P = ft.Param(beta=4.0, lat=(32, 32), tau=1.0, nstep=10)
B = 4096
x = ((torch.rand(B, 2, 32, 32, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
for w in (28, 14, 7, 4, 2, 1):
    per = 148 * w
    def fn():
        for a in range(0, B, per):
            ft.ft_hmc_batch(P, pf, x[a:a + per], seed=1, chain0=a)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(); fn(); b_.record(); torch.cuda.synchronize(); ts.append(a_.elapsed_time(b_))
    print(f"waves per launch {w:2d}: {min(ts):8.3f} ms for {B} chains -> {B / min(ts) * 1e3:8.1f} traj/s", flush=True)
