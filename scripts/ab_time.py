"""A/B timing of two builds of the library on the same box (development aid): python scripts/ab_time.py a.so b.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2:
    for so in sys.argv[1:]:
        print(so, flush=True)
        subprocess.run([sys.executable, __file__, so])
    sys.exit(0)
sys.path.insert(0, ROOT)
import numpy as np, torch
import fthmc_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
import ctypes
_h = ctypes.CDLL(L.LIB_PATH)
for _n in list(L.SIGNATURES):               # older builds lack the newer diagnostic exports
    if not hasattr(_h, _n):
        del L.SIGNATURES[_n]
import fthmc_b200 as ft
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
P = ft.Param(beta=4.0, lat=(32, 32), tau=1.0, nstep=10)
x = ((torch.rand(148, 2, 32, 32, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
print(f"   flow_fwd {timeit(lambda: ft.ft_flow(pf, x)):.3f}  flow_inv {timeit(lambda: ft.ft_flow_inv(pf, x)):.3f}  "
      f"ft_force {timeit(lambda: ft.ft_force(P, pf, x)):.3f}  ft_hmc {timeit(lambda: ft.ft_hmc_batch(P, pf, x, seed=1), 3):.3f} ms", flush=True)
