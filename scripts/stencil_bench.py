"""HBM roofline of the streaming Wilson stencils (action / force / topological charge / regularize) through the public
API: algorithmic bytes (16 B/site read; force and regularize also write 16 B/site resp. 8 B/link) over CUDA-event time, against
the measured copy bandwidth in MEASURED_PEAKS.json.  Inputs are larger than L2 (>= 512 MiB)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import fthmc_b200 as ft

try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

out = []
for dtype in (torch.float64, torch.float32):
    es = 8 if dtype == torch.float64 else 4
    for L in (32, 128, 1024):
        B = max(1, (768 << 20) // (2 * L * L * es))
        B = min(B, 65535)
        x = (torch.rand(B, 2, L, L, dtype=dtype, device="cuda") * 2 - 1) * 3.0
        P = ft.Param(beta=4.0, lat=(L, L))
        nbytes = x.numel() * es
        for name, fn, traffic in (("action", lambda: ft.action(P, x), nbytes), ("topocharge", lambda: ft.topocharge(x), nbytes),
                                  ("force", lambda: ft.force(P, x), 2 * nbytes), ("regularize", lambda: ft.regularize(x), 2 * nbytes)):
            ms = timeit(fn)
            gbs = traffic / (ms * 1e-3) / 1e9
            out.append(dict(op=name, dtype=str(dtype).split(".")[-1], L=L, B=B, ms=ms, gbs=gbs, frac=gbs / peak))
            print(f"{name:11s} {str(dtype).split('.')[-1]:8s} L={L:5d} B={B:6d}  {ms:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f} GB/s", flush=True)
json.dump(dict(peak_gbs=peak, results=out), open(os.path.join(ROOT, "gpurun_out", "stencil_bench.json"), "w"), indent=1)
