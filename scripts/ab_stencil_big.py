"""A/B timing of the generic (CTA / cluster per chain) reduction scans on large lattices with a given build: python scripts/ab_stencil_big.py lib.so"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import fthmc_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
import fthmc_b200 as ft
for dt in (torch.float64, torch.float32):
    for Lx, B in ((128, 3072), (256, 768), (1024, 48), (64, 600)):
        if dt == torch.float32: B *= 2
        x = ((torch.rand(B, 2, Lx, Lx, dtype=torch.float64, device="cuda") * 2 - 1) * 3.0).to(dt)
        P = ft.Param(beta=2.0, lat=(Lx, Lx))
        for name, fn in (("action", lambda: ft.action(P, x)), ("topo", lambda: ft.topocharge(x))):
            for _ in range(3): fn()
            ts = []
            for _ in range(9):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            t = sorted(ts)[4]
            print(f"{os.path.basename(sys.argv[1])} {name:6s} {str(dt)[6:]:8s} L={Lx:5d} B={B:5d} {x.numel() * x.element_size() / t / 1e6:6.0f} GB/s", flush=True)
