"""Flow-training step on N GPUs (development benchmark + NCCL check of the gradient all-reduce):
    python scripts/train_bench.py            |   torchrun --nproc-per-node 2 scripts/train_bench.py
Each rank draws its own prior batch, runs fthmc_ft_action_grad, all-reduces the (n_layers, 955) gradient and takes the
same Adam step.  Prints the weight-gradient kernel time next to a plain ft_force launch on the same batch, the step
time and the loss trace; with N > 1 also checks that every rank ends with identical weights."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.distributed as dist
import fthmc_b200 as ft

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L, B, layers = int(os.environ.get("FT_L", 32)), int(os.environ.get("FT_B", 592)), 24
tr = ft.FlowTrainer(ft.default_init_raw(layers, 3647), (L, L), beta=4.0, lr=1e-3, seed=100 + rank)
P = ft.Param(beta=4.0, lat=(L, L))
xi = tr.sample_prior(B)

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

t_force = timeit(lambda: ft.ft_force(P, tr.packed(), xi))
t_grad = timeit(lambda: ft.ft_action_grad(P, tr.packed(), xi))
t0 = time.perf_counter()
for it in range(10):
    m = tr.train_step(B)
torch.cuda.synchronize()
t_step = (time.perf_counter() - t0) / 10
if rank == 0:
    print(f"L={L} B={B}/GPU x {world} GPU(s): ft_force {t_force:.2f} ms, ft_action_grad (force + weight gradients + D2H/unpack) {t_grad:.2f} ms, "
          f"train_step {1e3 * t_step:.1f} ms  ({B * world / t_step:.0f} samples/s)")
    print("dkl trace:", " ".join(f"{v:.3f}" for v in tr.history["dkl"]), " ess:", f"{m['ess']:.4f}")
if world > 1:
    w = tr.raw.detach().cuda().clone()
    ref = w.clone(); dist.broadcast(ref, 0)
    same = torch.equal(w, ref)
    flag = torch.tensor([1.0 if same else 0.0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("identical weights on all ranks after 10 all-reduced steps:", bool(flag.item()))
    dist.destroy_process_group()
