"""The usage example of README.md, kept runnable."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fthmc_b200 as ft
param = ft.Param(beta=4.0, lat=(32, 32), tau=1.0, nstep=10, ntraj=64, nrun=2)
flow  = ft.PackedFlow(ft.default_init_raw(24, 3647))     # or ft.pack(reference_ModuleList) / ft.load_flow("ckpt-era9-epoch99.tar")
field = param.initializer()                               # (2, L, L), as in the reference
dH, exp_mdH, acc, new = ft.ft_hmc(param, flow, field[None])          # one trajectory, reference signature
field = ft.ft_run(param, flow, field)                     # nrun x ntraj trajectories, chain resident on the SM
r = ft.ft_hmc_batch(param, flow, torch.zeros(4096, 2, 32, 32, dtype=torch.float64).cuda(), seed=1)   # 4096 chains, one launch
trainer = ft.FlowTrainer(ft.default_init_raw(24, 3647), (8, 8), beta=2.0, lr=1e-3)                     # reverse-KL training
metrics = trainer.train_step(1024)
print("readme example ok:", float(dH), bool(acc), tuple(field.shape), float(r["dH"].mean()), metrics)
