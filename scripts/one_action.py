import os, sys
sys.path.insert(0, os.getcwd())
import torch
import fthmc_b200 as ft
x = ((torch.rand(49152, 2, 32, 32, dtype=torch.float64, device="cuda") * 2 - 1) * 3.0)
P = ft.Param(beta=2.0, lat=(32, 32))
for _ in range(3): a = ft.action(P, x)
torch.cuda.synchronize(); print(float(a.sum()))
