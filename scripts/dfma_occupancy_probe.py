"""fp64 FMA throughput vs resident warps per SM (development aid): is 8 warps/SM enough to fill the pipe?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fthmc_b200 as ft
lib = ft.lib()
scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
flop = ctypes.c_double()
for blocks in (148, 296, 592, 1184, 2368):
    best = 1e30
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ft._lib.check(lib.fthmc_diag_dfma_probe(scratch.data_ptr(), 40000, blocks, None, ctypes.byref(flop)))
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"blocks={blocks} ({blocks / 148 * 8:.0f} warps/SM): {flop.value / best / 1e9:.2f} TFLOP/s")
