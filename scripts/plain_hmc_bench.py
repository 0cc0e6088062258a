"""Plain HMC (hmc_2dU1.py path) on the resident-chain kernel: BASELINE config 1 (L=8, beta=2, one chain, tau=1, nstep=10)
as latency per trajectory through the drop-in `hmc(param, x)`, and batched throughput at L=8 / L=32."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft

def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

P = ft.Param(beta=2.0, lat=(8, 8), tau=1.0, nstep=10)
x = torch.empty(2, 8, 8, dtype=torch.float64).uniform_(-np.pi, np.pi)
for _ in range(5): ft.hmc(P, x)
t0 = time.perf_counter()
for _ in range(200): dH, e, acc, x = ft.hmc(P, x)
t = (time.perf_counter() - t0) / 200
print(f"config 1: hmc(param, x) L=8 single chain, host tensors in and out: {1e3 * t:.3f} ms per trajectory (reference CPU: 9.4 ms)")
xd = x.cuda()[None]
ms = ev(lambda: ft.hmc_run_batch(P, xd, 100, seed=1))
print(f"          resident run loop, 100 trajectories per launch, device tensors: {ms / 100 * 1e3:.1f} us per trajectory")
for L, B in ((8, 65536), (32, 16384)):
    P = ft.Param(beta=2.0, lat=(L, L), tau=1.0, nstep=10)
    xb = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    ms = ev(lambda: ft.hmc_run_batch(P, xb, 10, seed=1), 3)
    print(f"plain HMC L={L} B={B}: {B * 10 / ms * 1e3:.3e} trajectories/s ({B * 10 * L * L * 10 / ms * 1e3:.3e} site-updates/s)")
