import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fthmc_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
import fthmc_b200 as ft
def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for Lx, B in ((8, 65536), (32, 16384)):
    P = ft.Param(beta=2.0, lat=(Lx, Lx), tau=1.0, nstep=10)
    xb = ((torch.rand(B, 2, Lx, Lx, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    ms = ev(lambda: ft.hmc_run_batch(P, xb, 10, seed=1))
    print(f"{sys.argv[1]} plain HMC L={Lx} B={B}: {B * 10 / ms * 1e3:.3e} traj/s")
