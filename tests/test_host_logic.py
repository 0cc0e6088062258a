"""Host-side logic that needs no GPU: weight packing order, the random-init replay, mask parameters."""
import numpy as np
import torch

import fthmc_b200 as ft
from fthmc_b200 import flow as F


def test_default_init_replays_reference_stream(golden):
    """default_init_raw(24, 3647) == the reference's torch.manual_seed(3647); make_u1_equiv_layers(...)"""
    g = golden("ft_L32_b4")
    assert np.array_equal(F.default_init_raw(24, 3647), g["weights"])
    g = golden("ft_L16_b6")
    assert np.array_equal(F.default_init_raw(24, 3647), g["weights"])


def test_default_init_does_not_disturb_global_rng():
    torch.manual_seed(5)
    a = torch.rand(3)
    torch.manual_seed(5)
    F.default_init_raw(2, 1)
    assert torch.equal(a, torch.rand(3))


def test_mask_parameters_from_active_mask():
    """pack() reads (mu, off) back from a layer's link mask when it has one (reference layers do)."""
    from oracle import fthmc_oracle as O
    for mu in (0, 1):
        for off in range(4):
            m = O.link_active_mask((8, 12), mu, off)
            got_mu = 0 if bool(m[0].any()) else 1
            line = m[0][0, :] if got_mu == 0 else m[1][:, 0]
            assert (got_mu, int(torch.nonzero(line)[0])) == (mu, off)


def test_block_statistics_match_reference(golden):
    """fthmc_b200.stats against the reference's own functions (ipynb/ft_hmc.py:14-56) on a synthetic charge history."""
    from fthmc_b200 import stats
    g = golden("run_L8")
    hist = list(g["stat_hist"][100:])
    got = np.array(stats.change_sqr_vs_dt(hist, 10), dtype=np.float64)
    assert np.allclose(got, g["stat_change_sqr_vs_dt"], rtol=1e-13, atol=0)     # (numpy's pairwise sums vs the reference's sequential ones)
    assert [len(b) for b in stats.block_list(list(range(37)))] == list(g["stat_block_sizes"])
    assert np.allclose(np.array(stats.topo_change_sqr(list(g["stat_hist"]), 10)), g["stat_change_sqr_vs_dt"], rtol=1e-13, atol=0)
    m2, e2 = stats.mean_and_error(stats.block_means(np.arange(64.0)), standard_error=True)
    assert abs(m2 - 31.5) < 1e-12 and e2 > 0
    q = np.stack([g["stat_hist"], g["stat_hist"][::-1]], axis=1)
    m, e = stats.batched_topo_change_sqr(q, dt=1)
    assert m > 0 and e >= 0


def test_state_dict_and_checkpoint_formats(golden, tmp_path):
    """weights from a flow state_dict / a save_checkpoint-style dict (fthmc/utils/io.py:148-170) come out in the raw
    order the packer expects; mask tensors in the state_dict are ignored; malformed inputs are refused."""
    import torch.nn as nn
    import pytest
    g = golden("ft_L8_n8")
    shapes = [(8, 2, 3, 3), (8,), (8, 8, 3, 3), (8,), (3, 8, 3, 3), (3,)]
    sd = {}
    for i, row in enumerate(g["weights"]):
        pos = 0
        for k, (ws, bs) in zip((0, 2, 4), zip(shapes[0::2], shapes[1::2])):
            n = int(np.prod(ws)); sd[f"{i}.plaq_coupling.net.{k}.weight"] = torch.from_numpy(row[pos:pos + n].reshape(ws).copy()); pos += n
            n = int(np.prod(bs)); sd[f"{i}.plaq_coupling.net.{k}.bias"] = torch.from_numpy(row[pos:pos + n].reshape(bs).copy()); pos += n
        sd[f"{i}.active_mask"] = torch.zeros(2, 8, 8)
        sd[f"{i}.plaq_coupling.active_mask"] = torch.zeros(8, 8)
    assert np.array_equal(F.raw_from_state_dict(sd), g["weights"])
    ck = {"era": 1, "epoch": 2, "model_state_dict": sd, "optimizer_state_dict": {}}
    assert np.array_equal(F.raw_from_state_dict(ck), g["weights"])
    fn = tmp_path / "ckpt-era1-epoch2.tar"
    torch.save(ck, fn)
    assert np.array_equal(F.raw_from_state_dict(torch.load(fn, weights_only=False)), g["weights"])
    with pytest.raises(ft.FthmcError):
        F.raw_from_state_dict({"foo": torch.zeros(1)})
    bad = dict(sd); bad["0.plaq_coupling.net.0.weight"] = torch.zeros(4, 2, 3, 3)
    with pytest.raises(ft.FthmcError):
        F.raw_from_state_dict(bad)
