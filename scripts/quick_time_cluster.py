"""One ft_hmc batch of 7 chains at L=128 (one 16-CTA cluster each): the launch captured by `ncu --set full -k k_chain_cluster -c 1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fthmc_b200 as ft
B, L = int(os.environ.get("FT_B", "7")), int(os.environ.get("FT_L", "128"))
pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
P = ft.Param(beta=6.0, lat=(L, L), tau=1.0, nstep=10)
x = ((torch.rand(B, 2, L, L, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
r = ft.ft_hmc_batch(P, pf, x, seed=1)
torch.cuda.synchronize()
print("dH mean", float(r["dH"].mean()))
