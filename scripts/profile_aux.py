"""Launches captured by ncu for the auxiliary kernels: the streaming stencils on a batch larger than L2 and one
weight-gradient launch (MODE_FT_GRAD) of 148 samples."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft
what = sys.argv[1]
if what == "stencils":
    L, B = 32, 49152
    x = (torch.rand(B, 2, L, L, dtype=torch.float64, device="cuda") * 2 - 1) * 3.0
    P = ft.Param(beta=4.0, lat=(L, L))
    for _ in range(2):
        ft.action(P, x); ft.topocharge(x); ft.force(P, x); ft.regularize(x)
    torch.cuda.synchronize()
elif what == "grad":
    pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
    P = ft.Param(beta=4.0, lat=(32, 32))
    x = torch.rand(148, 2, 32, 32, dtype=torch.float64, device="cuda") * 2 * np.pi
    ft.ft_action_grad(P, pf, x)
    torch.cuda.synchronize()
if what == "plain":
    P = ft.Param(beta=2.0, lat=(32, 32), tau=1.0, nstep=10)
    xb = ((torch.rand(740 * 4, 2, 32, 32, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    ft.hmc_run_batch(P, xb, 2, seed=1)
    torch.cuda.synchronize()
