"""Quick device timing of the resident-chain kernel (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft
from oracle import fthmc_oracle as O

def raw_of(flow):
    return np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Bs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [148, 592]
flow = O.random_flow(n_layers=24, seed=3647)
pf = ft.PackedFlow(raw_of(flow))
P = ft.Param(beta=4.0, lat=(L, L), tau=1.0, nstep=10)
for B in Bs:
    x = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    t_fwd = timeit(lambda: ft.ft_flow(pf, x))
    t_inv = timeit(lambda: ft.ft_flow_inv(pf, x))
    t_frc = timeit(lambda: ft.ft_force(P, pf, x))
    t_trj = timeit(lambda: ft.ft_hmc_batch(P, pf, x, seed=1), n=2)
    print(f"L={L} B={B}: flow_fwd {t_fwd:.2f} ms  flow_inv {t_inv:.2f} ms  ft_force {t_frc:.2f} ms  ft_hmc {t_trj:.2f} ms "
          f"-> {B / t_trj * 1e3:.1f} traj/s", flush=True)
