"""Physics at the headline configuration (BASELINE config 3: L=32, beta=4, the 24-layer seed-3647 flow): 4096 device-RNG chains
of FT-HMC at nstep=40 (the setting that accepts about half) and of plain HMC, thermalised from a hot start, against the exact
<cos P> = I1(beta)/I0(beta) (the reference's PLAQ_EXACT table, fthmc/config.py:37-47) and against each other for <Q^2>."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.special import i0, i1
import fthmc_b200 as ft

L, beta, B = 32, 4.0, 4096
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
exact = float(i1(beta) / i0(beta))
gen = torch.Generator().manual_seed(5)
x0 = ((torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
for name, P, flow, ntherm, nmeas in (("FT-HMC nstep=40", ft.Param(beta=beta, lat=(L, L), tau=1.0, nstep=40), pf, 120, 60),
                                     ("plain HMC nstep=20", ft.Param(beta=beta, lat=(L, L), tau=1.0, nstep=20), None, 400, 200)):
    t0 = time.time()
    x = x0.clone()
    run = (lambda x, n, t: ft.ft_hmc_run_batch(P, flow, x, n, seed=11, traj0=t)) if flow is not None else \
          (lambda x, n, t: ft.hmc_run_batch(P, x, n, seed=11, traj0=t))
    r = run(x, ntherm, 0)
    x = r["field"]
    r = run(x, nmeas, ntherm)
    plaq = r["plaq"].double().cpu().numpy()          # (nmeas, B): plaquette after every trajectory
    q = r["topo"].double().cpu().numpy()
    acc = float(r["acc"].double().mean())
    per_chain = plaq.mean(axis=0)
    m, e = per_chain.mean(), per_chain.std(ddof=1) / np.sqrt(B)
    q2 = (q[-1] ** 2)
    print(f"{name:20s} acc {acc:.3f}  <cos P> = {m:.6f} +- {e:.6f}  (exact {exact:.6f}, deviation {(m - exact) / e:+.1f} sigma)  "
          f"<Q^2> = {q2.mean():.3f} +- {q2.std(ddof=1) / np.sqrt(B):.3f}  mean exp(-dH) = {float(r['exp_mdH'].double().mean()):.4f}  "
          f"[{time.time() - t0:.0f} s]", flush=True)
