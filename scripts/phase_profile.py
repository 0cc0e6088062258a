"""Per-phase cycle breakdown of the resident-chain kernel (development aid).  Builds a -DFT_PROFILE copy
of the library into gpurun_out/ (never the product .so) and runs one batch of ft_force / ft_hmc."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
out = os.path.join(ROOT, "gpurun_out", "libfthmc_prof.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
import __graft_entry__ as G
if os.environ.get("FT_PROF_SO"):
    out = os.path.abspath(os.environ["FT_PROF_SO"])      # a prebuilt -DFT_PROFILE library (A/B comparisons)
else:
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + G.NVCC_FLAGS + ["-DFT_PROFILE", "-o", out, os.path.join(ROOT, "fthmc_b200/csrc/fthmc_capi.cu")])
import fthmc_b200._lib as L
L.LIB_PATH = out
import fthmc_b200 as ft
lib = ft.lib()
names = ["planes", "conv1", "conv2", "conv3_fwd", "conv3_rev", "outgrad", "conv3T", "conv2T", "conv1T", "scatter", "issue", "wilson_force", "leap", "misc", "  c2_mac", "  c2_act", "  c1_mac", "  c1_act", "  c2T_mac", "  c3_conv_all"]
Lx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
P = ft.Param(beta=4.0, lat=(Lx, Lx), tau=1.0, nstep=10)
x = ((torch.rand(B, 2, Lx, Lx, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
buf = (ctypes.c_ulonglong * 32)()
for what in ("ft_force",):
    fn = (lambda: ft.ft_force(P, pf, x)) if what == "ft_force" else (lambda: ft.ft_hmc_batch(P, pf, x, seed=1))
    fn(); torch.cuda.synchronize()
    lib.fthmc_diag_profile(buf, 1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    lib.fthmc_diag_profile(buf, 1)
    ms = a.elapsed_time(b)
    tot = sum(buf[i] for i in range(14))
    print(f"{what}: {ms:.3f} ms for {B} chains; instrumented cycles/CTA = {tot / min(B, 148):.0f}")
    for i, n in enumerate(names):
        if buf[i]:
            print(f"   {n:13s} {100.0 * buf[i] / tot:5.1f}%   {buf[i] / min(B, 148) / 1e3:9.1f} kcycles/CTA")
