"""Timing across lattice sizes: BASELINE config 2 (L=16, 64 chains) and a full batch at L=16, the cluster path at L=64 and
L=128 (config 4), one wave at L=32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft

def timeit(fn, n=2):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
for L, Bs in ((16, (64, 4096)), (64, (33, 66)), (128, (7, 14)), (32, (148,))):
    P = ft.Param(beta=6.0, lat=(L, L), tau=1.0, nstep=10)
    for B in Bs:
        x = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
        t_frc = timeit(lambda: ft.ft_force(P, pf, x))
        t_trj = timeit(lambda: ft.ft_hmc_batch(P, pf, x, seed=1), n=1)
        print(f"L={L} B={B}: ft_force {t_frc:.2f} ms  ft_hmc {t_trj:.2f} ms -> {B / t_trj * 1e3:.1f} traj/s, "
              f"{B * L * L / t_trj * 1e3 / 1e6:.2f} Msite-traj/s", flush=True)
