"""Where a plain-HMC trajectory's time goes: trajectories/s at nstep = 1, 10, 40 (development aid)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fthmc_b200 as ft
def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
Lx, B = 32, 16384
xb = ((torch.rand(B, 2, Lx, Lx, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
for nstep in (1, 10, 40):
    P = ft.Param(beta=2.0, lat=(Lx, Lx), tau=1.0, nstep=nstep)
    ms = ev(lambda: ft.hmc_run_batch(P, xb, 10, seed=1))
    print(f"nstep={nstep}: {ms / 10:.3f} ms per trajectory of the batch", flush=True)
