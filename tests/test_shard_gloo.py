"""Multi-GPU host logic on CPU: chain partitioning and the observables all-reduce (gloo, world_size 2).
The data path has no collective (chains are independent); what N>1 adds is exactly this."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fthmc_b200 import shard


def test_chain_partition_covers_every_chain_once():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                c0, n = shard.chain_partition(total, r, world)
                seen.extend(range(c0, c0 + n))
            assert seen == list(range(total))
    assert shard.chain_partition(65536, 3, 8) == (3 * 8192, 8192)
    with pytest.raises(ValueError):
        shard.chain_partition(8, 2, 2)


def fake_result(chain0, count):
    """per-chain observables as a deterministic function of the GLOBAL chain index"""
    g = torch.arange(chain0, chain0 + count, dtype=torch.float64)
    return dict(plaq=torch.cos(g) * 0.5, topo=torch.floor(torch.sin(g) * 3), acc=(g.long() % 3 != 0),
                dH=torch.sin(0.1 * g), exp_mdH=torch.exp(-torch.sin(0.1 * g)))


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c0, n = shard.chain_partition(total, rank, world)
        sums = shard.allreduce_observables(shard.local_observable_sums(fake_result(c0, n)))
        if rank == 0:
            torch.save(sums, out)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total", [10, 4097])
def test_allreduce_observables_world2_matches_single_process(tmp_path, total):
    out = str(tmp_path / "sums.pt")
    mp.spawn(_worker, args=(2, _free_port(), total, out), nprocs=2, join=True)
    got = torch.load(out)
    want = shard.local_observable_sums(fake_result(0, total))
    assert torch.allclose(got, want, rtol=1e-13, atol=1e-12)
    assert float(got[6]) == total
    o = shard.Observables.from_sums(got)
    assert o.count == total and abs(o.acc_rate - float(want[3]) / total) < 1e-15


def test_single_process_is_identity():
    s = shard.local_observable_sums(fake_result(0, 5))
    assert torch.equal(shard.allreduce_observables(s.clone()), s)


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.arange(24 * 955, dtype=torch.float64).reshape(24, 955) * (rank + 1)
        sums = torch.tensor([10.0 * (rank + 1), 64.0], dtype=torch.float64)
        g2, s2 = shard.allreduce_gradient(g, sums)
        if rank == 0:
            torch.save((g2, s2), out)
    finally:
        dist.destroy_process_group()


def test_allreduce_gradient_world2(tmp_path):
    """the flow-training gradient (n_layers x 955) and the loss sums travel in one all-reduce"""
    out = str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    g, s = torch.load(out)
    base = torch.arange(24 * 955, dtype=torch.float64).reshape(24, 955)
    assert torch.equal(g, base * 3) and torch.equal(s, torch.tensor([30.0, 128.0], dtype=torch.float64))
    g1, s1 = shard.allreduce_gradient(base, None)
    assert g1 is base and s1 is None
