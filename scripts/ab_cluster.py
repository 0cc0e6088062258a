"""A/B timing of the cluster path (L=64, L=128) with a given build of the library: python scripts/ab_cluster.py lib.so"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import fthmc_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
import fthmc_b200 as ft
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
for Lx, B in ((64, 33), (128, 7), (48, 37)):
    P = ft.Param(beta=6.0, lat=(Lx, Lx), tau=1.0, nstep=10)
    x = ((torch.rand(B, 2, Lx, Lx, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    t_fwd = timeit(lambda: ft.ft_flow(pf, x)); t_frc = timeit(lambda: ft.ft_force(P, pf, x)); t_trj = timeit(lambda: ft.ft_hmc_batch(P, pf, x, seed=1), 2)
    print(f"{sys.argv[1]} L={Lx} B={B}: flow_fwd {t_fwd:.3f}  ft_force {t_frc:.3f}  ft_hmc {t_trj:.2f} ms -> {B / t_trj * 1e3:.1f} traj/s", flush=True)
