"""The CPU oracle (oracle/fthmc_oracle.py) against golden vectors dumped from the real reference
(tests/golden/make_golden.py).  Everything that is a pure torch-op sequence must match bit for bit."""
import numpy as np
import pytest
import torch

from conftest import oracle_flow_from_golden
from oracle import fthmc_oracle as O

T = torch.from_numpy


def test_plain_pointwise(golden):
    g = golden("plain_L8")
    beta = float(g["beta"])
    x = T(g["x0"])
    assert float(O.action(beta, x)) == float(g["action"])
    assert np.array_equal(O.force(beta, x).numpy(), g["force"])
    assert np.allclose(O.force_closed_form(beta, x).numpy(), g["force"], rtol=0, atol=1e-14)
    assert float(O.topocharge(x)) == float(g["topo"])
    assert np.array_equal(O.regularize(T(g["reg_in"])).numpy(), g["reg_out"])
    lx, lp = O.leapfrog(beta, float(g["dt"]), int(g["nstep"]), x, T(g["lf_p"]))
    assert np.array_equal(lx.numpy(), g["lf_x_out"]) and np.array_equal(lp.numpy(), g["lf_p_out"])
    z = torch.zeros(2, 8, 8)
    assert float(O.action(beta, z) / (-beta * 64)) == float(g["cold_plaq"]) == 1.0
    assert float(O.topocharge(z)) == float(g["cold_topo"]) == 0.0


def test_plain_hmc_teacher_forced(golden):
    g = golden("plain_L8")
    beta, dt, nstep = float(g["beta"]), float(g["dt"]), int(g["nstep"])
    for n in range(len(g["traj_u"])):
        dH, e, acc, new = O.hmc(beta, dt, nstep, T(g["traj_x"][n]), p=T(g["traj_p"][n]),
                                u=torch.tensor(g["traj_u"][n]))
        assert float(dH) == g["traj_dH"][n]
        assert bool(acc) == bool(g["traj_acc"][n])
        assert np.array_equal(new.numpy(), g["traj_out"][n])
        assert float(O.topocharge(new)) == g["traj_topo"][n]


def test_plain_hmc_rng_order(golden):
    """Without explicit p/u the oracle must draw from torch's RNG in the reference's order."""
    g = golden("plain_L8")
    torch.manual_seed(5000)
    dH, e, acc, new = O.hmc(float(g["beta"]), float(g["dt"]), int(g["nstep"]), T(g["traj_x"][0]))
    assert float(dH) == g["traj_dH"][0] and bool(acc) == bool(g["traj_acc"][0])


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4"])
def test_flow_pointwise(golden, name):
    g = golden(name)
    flow = oracle_flow_from_golden(g)
    beta = float(g["beta"])
    x = T(g["x"])
    cur = x
    for i, lw in enumerate(flow.layers):
        cur, lj = O.layer_forward(flow, lw, cur)
        assert np.array_equal(lj.numpy(), g["layer_logJ"][i]), f"logJ layer {i}"
        if i < g["layer_out"].shape[0]:
            assert np.array_equal(cur.numpy(), g["layer_out"][i]), f"layer {i}"
    assert np.array_equal(O.ft_flow(flow, x).numpy(), g["flow_fwd"])
    assert np.array_equal(O.ft_action(beta, flow, x).numpy(), g["ft_action"])
    assert np.array_equal(O.ft_force(beta, flow, x).numpy(), g["ft_force"])
    y = T(g["flow_fwd"])
    for b in range(x.shape[0]):
        inv = O.ft_flow_inv(flow, y[b:b + 1])
        assert np.array_equal(inv.numpy()[0], g["flow_inv_of_fwd"][b])
        _, lj = O.layer_reverse(flow, flow.layers[-1], y[b:b + 1])
        assert float(lj) == g["last_layer_reverse_logJ"][b]


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6"])
def test_ft_hmc_teacher_forced(golden, name):
    g = golden(name)
    flow = oracle_flow_from_golden(g)
    beta, dt, nstep = float(g["beta"]), float(g["dt"]), int(g["nstep"])
    for n in range(len(g["traj_u"])):
        dH, e, acc, new = O.ft_hmc(beta, dt, nstep, flow, T(g["traj_x"][n][None]), p=T(g["traj_p"][n][None]),
                                   u=torch.tensor(g["traj_u"][n]))
        assert dH == g["traj_dH"][n]
        assert bool(acc) == bool(g["traj_acc"][n])
        assert np.array_equal(new.numpy()[0], g["traj_out"][n])
        assert float(O.topocharge(new[0])) == g["traj_topo"][n]


def test_ft_hmc_L32_one_traj(golden):
    g = golden("ft_L32_b4")
    flow = oracle_flow_from_golden(g)
    dH, e, acc, new = O.ft_hmc(float(g["beta"]), float(g["dt"]), int(g["nstep"]), flow, T(g["traj_x"][0][None]),
                               p=T(g["traj_p"][0][None]), u=torch.tensor(g["traj_u"][0]))
    assert dH == g["traj_dH"][0] and bool(acc) == bool(g["traj_acc"][0])
    assert np.array_equal(new.numpy()[0], g["traj_out"][0])


@pytest.mark.parametrize("name", ["leaky_L8", "copyB_L8"])
def test_variants(golden, name):
    """LeakyReLU flow of hmc_2dU1.py and the [-pi,pi) convention of the package copy."""
    g = golden(name)
    flow = oracle_flow_from_golden(g)
    x = T(g["x"])
    y, logJ = O.ft_flow_logJ(flow, x)
    assert np.array_equal(y.numpy(), g["flow_fwd"])
    assert np.array_equal(logJ.numpy(), g["logJ"])
    for b in range(x.shape[0]):
        inv = O.ft_flow_inv(flow, y[b:b + 1])
        assert np.array_equal(inv.numpy()[0], g["flow_inv_of_fwd"][b])


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "leaky_L8"])
def test_adjoint_matches_autograd(golden, name):
    """The hand-derived reverse sweep (what the CUDA kernel implements) equals autograd."""
    g = golden(name)
    flow = oracle_flow_from_golden(g)
    x = T(g["x"])
    beta = float(g["beta"]) if "beta" in g else 3.0
    a = O.ft_force(beta, flow, x)
    b = O.ft_force_adjoint(beta, flow, x)
    assert torch.max(torch.abs(a - b)) <= 1e-12 * max(1.0, float(torch.max(torch.abs(a))))


def test_known_answers():
    """Reference inline self-checks: round trip to the bisection tolerance, logJ
    antisymmetry, gauge invariance of action/charge/logJ."""
    flow = O.random_flow(n_layers=8, seed=7, scale=2.0)
    torch.manual_seed(3)
    x = torch.empty(1, 2, 8, 8).uniform_(-np.pi, np.pi)
    y, logJ = O.ft_flow_logJ(flow, x)
    cur, lj_rev = y, 0.0
    for lw in reversed(flow.layers):
        cur, lj = O.layer_reverse(flow, lw, cur)
        lj_rev = lj_rev + lj
    assert torch.max(torch.abs(O.wrap_pi(cur - x))) < 2e-5
    assert abs(float(lj_rev + logJ)) < 1e-4
    alpha = 2 * np.pi * torch.rand(1, 8, 8)
    xg = O.gauge_transform(x, alpha)
    assert abs(float(O.u1_action(2.0, x) - O.u1_action(2.0, xg))) < 1e-10
    assert abs(float(O.topo_charge(x) - O.topo_charge(xg))) < 1e-10
    _, logJg = O.ft_flow_logJ(flow, xg)
    assert abs(float(logJ - logJg)) < 1e-9


def test_run_loops_follow_reference_chain(golden):
    """A chain seeded once (momenta / uniforms from consecutive generator draws, the order of run / ft_run):
    the oracle reproduces the reference's free-running chain, accepts and rejects included."""
    g = golden("run_L8")
    x0 = torch.from_numpy(g["x0"])
    torch.manual_seed(int(g["seed"]))
    f = x0.clone()
    for i in range(int(g["ntraj"])):
        dH, e, acc, f = O.hmc(float(g["plain_beta"]), float(g["plain_tau"]) / int(g["plain_nstep"]), int(g["plain_nstep"]), f)
        assert abs(float(dH) - g["plain_dH"][i]) < 1e-9 and bool(acc) == bool(g["plain_acc"][i])
        assert float(O.topocharge(f)) == g["plain_topo"][i]
    assert np.max(np.abs(f.numpy() - g["plain_final"])) < 1e-10
    flow = oracle_flow_from_golden(dict(weights=g["weights"], activation="silu", convention=0))
    torch.manual_seed(int(g["seed"]))
    f = x0.clone()
    for i in range(int(g["ntraj"])):
        dH, e, acc, fr = O.ft_hmc(float(g["ft_beta"]), float(g["ft_tau"]) / int(g["ft_nstep"]), int(g["ft_nstep"]), flow, f[None])
        f = fr[0]
        assert abs(float(dH) - g["ft_dH"][i]) < 1e-8 and bool(acc) == bool(g["ft_acc"][i])
        assert float(O.topocharge(f)) == g["ft_topo"][i]
    assert np.max(np.abs(f.numpy() - g["ft_final"])) < 1e-8


@pytest.mark.parametrize("name,picks", [("ft_L16_b6_many", (0, 3, 205, 240)), ("ft_L32_b4_many", (1, 230))])
def test_headline_configs_sample(golden, name, picks):
    """A sample of the 260-trajectory reference runs at BASELINE configs 2 and 3 (the GPU suite checks all of them): the
    oracle reproduces the reference's dH bit for bit, its decision and its floored charge."""
    from conftest import thousand_inputs
    g = golden(name)
    flow = oracle_flow_from_golden(g)
    L = int(g["L"])
    x, p, u = thousand_inputs(L, len(g["dH"]), seed=int(g["seed"]))
    for i in picks:
        nstep = int(g["nstep"][i])
        dH, e, acc, new = O.ft_hmc(float(g["beta"]), float(g["tau"]) / nstep, nstep, flow, x[i][None], p=p[i][None], u=u[i])
        assert dH == g["dH"][i] and bool(acc) == bool(g["acc"][i])
        assert float(O.topocharge(new[0])) == g["topo"][i]
        assert float(new.sum()) == g["field_sum"][i]


def test_copyB_physics(golden):
    """Copy B's helpers (fthmc/utils/qed_helpers.py) against the oracle in the package conventions: the oracle is a
    restatement of copy A, so agreement here is to rounding (different term order in the plaquette), not bit for bit."""
    g = golden("copyB_physics_L8")
    flow = oracle_flow_from_golden(g)
    assert flow.convention == 1
    beta, x = float(g["beta"]), T(g["x"])
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(np.asarray(b))))
    assert rel(O.u1_action(beta, x).numpy(), g["batch_action"]) < 1e-13
    assert np.max(np.abs(O.topo_charge(x).numpy() - g["batch_charges"])) < 1e-12
    assert np.max(np.abs(O.topo_charge(x).numpy() - g["topo_charge"])) < 1e-12      # (the two differ in their last ulp: term order)
    assert np.max(np.abs(O.ft_flow(flow, x).numpy() - g["ft_flow"])) < 1e-12
    assert rel(O.ft_action(beta, flow, x).numpy(), g["ft_action"]) < 1e-12
    assert rel(O.ft_force(beta, flow, x).numpy(), g["ft_force"]) < 1e-12
    for b in range(x.shape[0]):
        xi = O.ft_flow_inv(flow, T(g["ft_flow"][b:b + 1]))
        assert np.max(np.abs(xi.numpy()[0] - g["ft_flow_inv_of_fwd"][b])) < 1e-11
    dt, nstep = float(g["dt"]), int(g["nstep"])
    assert abs(float(O.action(beta, x[0])) - float(g["action"])) < 1e-12
    assert np.max(np.abs(O.force(beta, x[0]).numpy() - g["force"])) < 1e-13
    lx, lp = O.leapfrog(beta, dt, nstep, x[0], T(g["lf_p"]))
    assert np.max(np.abs(lx.numpy() - g["lf_x_out"])) < 1e-13 and np.max(np.abs(lp.numpy() - g["lf_p_out"])) < 1e-13
    for n in range(len(g["traj_u"])):
        dH, e, acc, new = O.hmc(beta, dt, nstep, T(g["traj_x"][n]), p=T(g["traj_p"][n]), u=torch.tensor(g["traj_u"][n]))
        assert abs(float(dH) - g["traj_dH"][n]) < 1e-11 and bool(acc) == bool(g["traj_acc"][n])
        assert np.max(np.abs(new.numpy() - g["traj_out"][n])) < 1e-12
