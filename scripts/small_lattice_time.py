import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import fthmc_b200 as ft
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
for L, Bs in ((8, (64, 16384)), (16, (64, 4096)), (24, (148, 1184))):
    P = ft.Param(beta=6.0, lat=(L, L), tau=1.0, nstep=10)
    for B in Bs:
        x = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
        t = timeit(lambda: ft.ft_hmc_batch(P, pf, x, seed=1), 2)
        print(f"L={L} B={B}: ft_hmc {t:.2f} ms -> {B / t * 1e3:.0f} traj/s", flush=True)
