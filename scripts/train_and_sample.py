"""The reference's workflow end to end on the CUDA path (ipynb/ft_hmc.py:519-587 in miniature): train the flow with the
reverse-KL loss, then run FT-HMC through it and compare with plain HMC and with the untrained flow.
    python scripts/train_and_sample.py [L] [beta] [train_steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
beta = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
torch.manual_seed(1331)
raw0 = ft.default_init_raw(24, 3647)
tr = ft.FlowTrainer(raw0, (L, L), beta=beta, lr=1e-3, seed=7)
t0 = time.perf_counter()
for it in range(steps):
    m = tr.train_step(1024)
    if it % max(1, steps // 8) == 0 or it == steps - 1:
        print(f"step {it:4d}  dkl {m['dkl']:10.3f}  ess {m['ess']:.4f}", flush=True)
torch.cuda.synchronize()
print(f"{steps} train steps of 1024 samples: {time.perf_counter() - t0:.1f} s")

B, ntraj = 2048, 60
P = ft.Param(beta=beta, lat=(L, L), tau=1.0, nstep=10)
x0 = torch.zeros(B, 2, L, L, dtype=torch.float64).cuda()                  # cold start
exact = {1.0: 0.44639, 2.0: 0.69777, 3.0: 0.80999, 4.0: 0.86352, 5.0: 0.89338, 6.0: 0.91236}.get(beta)   # fthmc/config.py:37-47
for name, runner in (("plain HMC", lambda: ft.hmc_run_batch(P, x0, ntraj, seed=1)),
                     ("FT-HMC, untrained flow", lambda: ft.ft_hmc_run_batch(P, ft.PackedFlow(raw0), x0, ntraj, seed=1)),
                     ("FT-HMC, trained flow", lambda: ft.ft_hmc_run_batch(P, tr.packed(), x0, ntraj, seed=1))):
    r = runner()
    half = ntraj // 2
    acc = float(r["acc"][half:].double().mean())
    plaq = float(r["plaq"][half:].mean())
    dq2 = ft.stats.batched_topo_change_sqr(r["topo"].cpu().numpy(), dt=1)
    print(f"{name:24s} acc {acc:.3f}  <plaq> {plaq:.5f} (exact {exact})  <dQ^2> per trajectory {dq2[0]:.4f} +- {dq2[1]:.4f}  <|dH|> {float(r['dH'][half:].abs().mean()):.3f}")
