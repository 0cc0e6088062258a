"""ft_hmc launch time against the number of chain waves (development aid): B = 148 * w chains at L = 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
P = ft.Param(beta=4.0, lat=(32, 32), tau=1.0, nstep=10)
for w in (1, 2, 4, 8, 16, 27, 28):
    B = 148 * w
    x = ((torch.rand(B, 2, 32, 32, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    fn = lambda: ft.ft_hmc_batch(P, pf, x, seed=1)
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"waves {w:2d}  B={B:5d}  {min(ts):8.3f} ms  -> {min(ts) / w:6.3f} ms per wave", flush=True)
