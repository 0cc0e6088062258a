"""Sharding of independent Markov chains over the GPUs of one box, and the only collective the
trajectory path has: the all-reduce of per-chain observables.

Chains never exchange data inside a trajectory, so ranks take contiguous blocks of the global chain
index and run the same kernels; the counter RNG is keyed by the GLOBAL chain index
(`chain0 + b`), which makes a run independent of how many GPUs it is spread over.  The per-trajectory
run-loop line of the reference (`ipynb/ft_hmc.py:456-467`: dH, exp(-dH), accept, plaq, topo) becomes
one 7-double sum-reduction per measurement.
"""
from dataclasses import dataclass

import torch

OBS_FIELDS = ("plaq", "Q", "Q2", "acc", "dH", "exp_mdH", "count")


def chain_partition(total_chains, rank, world):
    """Contiguous block of the global chain index owned by `rank`: (chain0, count).
    The first `total % world` ranks take one extra chain."""
    if world <= 0 or not (0 <= rank < world) or total_chains < 0:
        raise ValueError(f"bad partition request total={total_chains} rank={rank} world={world}")
    base, extra = divmod(total_chains, world)
    count = base + (1 if rank < extra else 0)
    chain0 = rank * base + min(rank, extra)
    return chain0, count


def local_observable_sums(result):
    """(7,) fp64 sums over this rank's chains of a `ft_hmc_batch` / `hmc_batch` result dict, on the
    device the result lives on: [sum plaq, sum Q, sum Q^2, sum acc, sum dH, sum exp(-dH), count]."""
    q = result["topo"].double()
    rows = torch.stack([result["plaq"].double(), q, q * q, result["acc"].double(), result["dH"].double(),
                        result["exp_mdH"].double(), torch.ones_like(q)])          # (7, B): one reduction instead of six
    return rows.sum(dim=1)


def allreduce_observables(sums, group=None):
    """Sum the (7,) vector over all ranks (NCCL on the GPUs, gloo in the CPU tests).  A single
    process (no initialised process group) returns it unchanged."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


@dataclass
class Observables:
    plaq: float
    Q: float
    Q2: float
    acc_rate: float
    mean_dH: float
    mean_exp_mdH: float
    count: int

    @staticmethod
    def from_sums(sums):
        s = [float(v) for v in sums.detach().cpu()]
        n = s[6]
        if n <= 0:
            raise ValueError("no chains in the observable sums")
        return Observables(plaq=s[0] / n, Q=s[1] / n, Q2=s[2] / n, acc_rate=s[3] / n, mean_dH=s[4] / n,
                           mean_exp_mdH=s[5] / n, count=int(round(n)))


def allreduce_gradient(grad, sums=None, group=None):
    """Sum the flow-training gradient (n_layers, 955) and an optional small vector of loss terms over all ranks
    (BASELINE config 5: NCCL all-reduce of the flow-training gradient).  Host tensors are staged through the device
    when the process group is NCCL.  A single process returns its inputs unchanged."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return grad, sums
    nccl = dist.get_backend(group) == "nccl"
    flat = torch.cat([grad.reshape(-1).double(), sums.reshape(-1).double() if sums is not None else grad.new_zeros(0)])
    buf = flat.cuda() if nccl and not flat.is_cuda else flat
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    buf = buf.to(grad.device)
    g = buf[:grad.numel()].reshape(grad.shape)
    return g, (buf[grad.numel():].reshape(sums.shape) if sums is not None else None)
