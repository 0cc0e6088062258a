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
    assert np.array_equal(got, g["stat_change_sqr_vs_dt"])
    assert [len(b) for b in stats.block_list(list(range(37)))] == list(g["stat_block_sizes"])
    assert np.array_equal(np.array(stats.topo_change_sqr(list(g["stat_hist"]), 10)), g["stat_change_sqr_vs_dt"])
    q = np.stack([g["stat_hist"], g["stat_hist"][::-1]], axis=1)
    m, e = stats.batched_topo_change_sqr(q, dt=1)
    assert m > 0 and e >= 0
