import os, sys
sys.path.insert(0, os.getcwd())
import torch
import fthmc_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
import fthmc_b200 as ft
for dt, Lx, B in ((torch.float64, 32, 49152), (torch.float64, 128, 3072), (torch.float32, 32, 65535)):
    x = ((torch.rand(B, 2, Lx, Lx, dtype=torch.float64, device="cuda") * 2 - 1) * 3.0).to(dt)
    fn = lambda: ft.topo_charge(x)
    for _ in range(3): fn()
    ts = []
    for _ in range(9):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = sorted(ts)[4]
    print(f"{os.path.basename(sys.argv[1])} batched topo_charge {str(dt)[6:]} L={Lx} {x.numel() * x.element_size() / t / 1e6:.0f} GB/s", flush=True)
