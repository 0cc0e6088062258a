"""The chain engine's device code (fthmc_b200/csrc/chain_engine.cuh), compiled serially for the CPU,
against the golden vectors from the reference.  This checks the algorithm the CUDA kernels run
(canonical stripe geometry, packed weights, pruned convolutions, bisection inverse, hand-written
adjoint, trajectory logic) without a GPU; the `-m gpu` tests check the same through the C ABI."""
import numpy as np
import pytest

import emul_lib as E


def relerr(a, b):
    return np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b)))


def wrap(x):
    return np.remainder(x + np.pi, 2 * np.pi) - np.pi


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4"])
def test_flow_forward_action_force(golden, name):
    g = golden(name)
    x, w, beta = g["x"], g["weights"], float(g["beta"])
    o = E.run("flow_fwd", w, x)
    assert np.max(np.abs(o["field"] - g["flow_fwd"])) < 1e-12
    assert np.max(np.abs(o["layer_logJ"] - g["layer_logJ"].T)) < 1e-11
    assert relerr(o["s"], g["layer_logJ"].sum(axis=0)) < 1e-12
    o = E.run("ft_action", w, x, beta=beta)
    assert relerr(o["s"], g["ft_action"]) < 1e-12
    o = E.run("ft_force", w, x, beta=beta)
    assert relerr(o["field"], g["ft_force"]) < 1e-11


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4"])
def test_flow_inverse_bit_exact_midpoints(golden, name):
    g = golden(name)
    o = E.run("flow_inv", g["weights"], g["flow_fwd"])
    # the bisection returns dyadic midpoints: same decisions => agreement to rounding, not to 1e-6
    assert np.max(np.abs(o["field"] - g["flow_inv_of_fwd"])) < 1e-11
    assert np.max(np.abs(wrap(o["field"] - g["x"]))) < 2e-5
    assert np.all(o["iters"] > 10) and np.all(o["iters"] < 40)


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4", "ft_L32_b4_n40"])
def test_ft_hmc_teacher_forced(golden, name):
    g = golden(name)
    n = len(g["traj_u"])
    o = E.run("ft_hmc", g["weights"], g["traj_x"], beta=float(g["beta"]), dt=float(g["dt"]), nstep=int(g["nstep"]),
              p=g["traj_p"], u=g["traj_u"])
    assert np.max(np.abs(o["s"] - g["traj_dH"])) < 1e-8
    assert np.array_equal(o["acc"].astype(bool), g["traj_acc"])
    assert np.array_equal(o["topo"], g["traj_topo"])
    assert np.max(np.abs(o["plaq"] - g["traj_plaq"])) < 1e-12
    assert np.max(np.abs(o["field"] - g["traj_out"])) < 1e-8
    assert n >= 1


def test_plain_hmc_teacher_forced(golden):
    g = golden("plain_L8")
    o = E.run("hmc", None, g["traj_x"], beta=float(g["beta"]), dt=float(g["dt"]), nstep=int(g["nstep"]),
              p=g["traj_p"], u=g["traj_u"])
    assert np.max(np.abs(o["s"] - g["traj_dH"])) < 1e-11
    assert np.array_equal(o["acc"].astype(bool), g["traj_acc"])
    assert np.array_equal(o["topo"], g["traj_topo"])
    assert np.max(np.abs(o["field"] - g["traj_out"])) < 1e-12
    o = E.run("leapfrog", None, g["x0"][None], beta=float(g["beta"]), dt=float(g["dt"]), nstep=int(g["nstep"]),
              p=g["lf_p"][None])
    assert np.max(np.abs(o["field"][0] - g["lf_x_out"])) < 1e-13
    assert np.max(np.abs(o["p"][0] - g["lf_p_out"])) < 1e-13


@pytest.mark.parametrize("name", ["leaky_L8", "copyB_L8"])
def test_variants(golden, name):
    g = golden(name)
    kw = dict(act=str(g["activation"]), conv=int(g["convention"]))
    o = E.run("flow_fwd", g["weights"], g["x"], **kw)
    assert np.max(np.abs(o["field"] - g["flow_fwd"])) < 1e-12
    assert relerr(o["s"], g["logJ"]) < 1e-12
    o = E.run("flow_inv", g["weights"], g["flow_fwd"], **kw)
    assert np.max(np.abs(o["field"] - g["flow_inv_of_fwd"])) < 1e-11


def test_rectangular_lattice_against_oracle():
    """L0 != L1 exercises the row/column bookkeeping of the two mask orientations."""
    import torch
    from oracle import fthmc_oracle as O
    flow = O.random_flow(n_layers=8, seed=11, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()])
                                    for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    torch.manual_seed(5)
    x = torch.empty(2, 2, 8, 12).uniform_(-np.pi, np.pi)
    y, lj = O.ft_flow_logJ(flow, x)
    o = E.run("flow_fwd", raw, x.numpy())
    assert np.max(np.abs(o["field"] - y.numpy())) < 1e-12
    assert relerr(o["s"], lj.numpy()) < 1e-12
    f = O.ft_force(2.5, flow, x)
    o = E.run("ft_force", raw, x.numpy(), beta=2.5)
    assert relerr(o["field"], f.numpy()) < 1e-11


def test_thousand_trajectories_bit_exact_decisions(golden):
    """north_star: accept/reject decisions and integer topological charges bit-exact over 1000
    trajectories with the reference's momenta and uniforms; dH within 1e-8."""
    from conftest import thousand_inputs
    x, p, u = thousand_inputs(8)
    g = golden("plain_L8_1000")
    o = E.run("hmc", None, x.numpy(), beta=float(g["beta"]), dt=float(g["dt"]), nstep=int(g["nstep"]), p=p.numpy(), u=u.numpy())
    assert np.max(np.abs(o["s"] - g["dH"])) < 1e-8
    assert np.array_equal(o["acc"].astype(bool), g["acc"]) and np.array_equal(o["topo"], g["topo"])
    assert np.max(np.abs(o["field"].sum(axis=(1, 2, 3)) - g["field_sum"])) < 1e-9
    g = golden("ft_L8_1000")
    o = E.run("ft_hmc", g["weights"], x.numpy(), beta=float(g["beta"]), dt=float(g["dt"]), nstep=int(g["nstep"]),
              p=p.numpy(), u=u.numpy())
    assert np.max(np.abs(o["s"] - g["dH"])) < 1e-8
    assert np.array_equal(o["acc"].astype(bool), g["acc"]) and np.array_equal(o["topo"], g["topo"])
    assert np.max(np.abs(o["field"].sum(axis=(1, 2, 3)) - g["field_sum"])) < 1e-7


# ---------------------------------------------------------------- cluster decomposition (L = 64 .. 128 path)
@pytest.mark.parametrize("name,nranks", [("ft_L8_n8", 2), ("ft_L16_b6", 2), ("ft_L16_b6", 4), ("ft_L32_b4", 8)])
def test_cluster_ranks_match_golden(golden, name, nranks):
    """The kCluster code path (row-block X/GR, column-block planes, halo pushes, cluster-wide reductions) with host
    threads standing in for the CTAs of a cluster: same golden vectors, same tolerances as the single-CTA path."""
    g = golden(name)
    x, w, beta = g["x"], g["weights"], float(g["beta"])
    o = E.run("flow_fwd", w, x, nranks=nranks)
    assert np.max(np.abs(o["field"] - g["flow_fwd"])) < 1e-12
    assert np.max(np.abs(o["layer_logJ"] - g["layer_logJ"].T)) < 1e-11
    assert relerr(E.run("ft_action", w, x, beta=beta, nranks=nranks)["s"], g["ft_action"]) < 1e-12
    assert relerr(E.run("ft_force", w, x, beta=beta, nranks=nranks)["field"], g["ft_force"]) < 1e-11
    o = E.run("flow_inv", w, g["flow_fwd"], nranks=nranks)
    assert np.max(np.abs(o["field"] - g["flow_inv_of_fwd"])) < 1e-11
    o = E.run("ft_hmc", w, g["traj_x"], beta=beta, dt=float(g["dt"]), nstep=int(g["nstep"]), p=g["traj_p"], u=g["traj_u"],
              nranks=nranks)
    assert np.max(np.abs(o["s"] - g["traj_dH"])) < 1e-8
    assert np.array_equal(o["acc"].astype(bool), g["traj_acc"]) and np.array_equal(o["topo"], g["traj_topo"])
    assert np.max(np.abs(o["field"] - g["traj_out"])) < 1e-8


def test_cluster_decomposition_is_invisible():
    """Same chain, 1 / 2 / 3 ranks, rectangular lattice, device-RNG mode: fields bit-identical (only the order of
    the action sums changes), plain HMC included."""
    import torch
    from oracle import fthmc_oracle as O
    flow = O.random_flow(n_layers=6, seed=2, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()])
                                    for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    torch.manual_seed(9)
    x = torch.empty(2, 2, 24, 12).uniform_(-np.pi, np.pi).numpy()
    ref = E.run("ft_hmc", raw, x, beta=3.0, dt=0.05, nstep=4, seed=5, traj=1)
    refp = E.run("hmc", None, x, beta=3.0, dt=0.05, nstep=4, seed=5, traj=1)
    for nr in (1, 3):
        o = E.run("ft_hmc", raw, x, beta=3.0, dt=0.05, nstep=4, seed=5, traj=1, nranks=nr)
        assert np.array_equal(o["field"], ref["field"]) and np.max(np.abs(o["s"] - ref["s"])) < 1e-10
        assert np.array_equal(o["topo"], ref["topo"]) and np.array_equal(o["acc"], ref["acc"])
        o = E.run("hmc", None, x, beta=3.0, dt=0.05, nstep=4, seed=5, traj=1, nranks=nr)
        assert np.array_equal(o["field"], refp["field"]) and np.max(np.abs(o["s"] - refp["s"])) < 1e-10


def replay_draws(seed, shape, n):
    """the momenta and uniforms a run loop seeded once would draw: randn_like(field), rand([]) per trajectory"""
    import torch
    st = torch.get_rng_state()
    torch.manual_seed(seed)
    ps, us = [], []
    for _ in range(n):
        ps.append(torch.randn(shape, dtype=torch.float64).numpy())
        us.append(float(torch.rand([], dtype=torch.float64)))
    torch.set_rng_state(st)
    return np.stack(ps), np.array(us)


@pytest.mark.parametrize("nranks", [0, 2])
def test_run_loops_resident_chain(golden, nranks):
    """ntraj trajectories in ONE call with the field resident (run / ft_run): the reference's free-running chain,
    accepts and rejects included; and identical to calling trajectory by trajectory."""
    g = golden("run_L8")
    n, seed = int(g["ntraj"]), int(g["seed"])
    p, u = replay_draws(seed, (2, 8, 8), n)
    x0 = g["x0"][None]
    kw = dict(beta=float(g["plain_beta"]), dt=float(g["plain_tau"]) / int(g["plain_nstep"]), nstep=int(g["plain_nstep"]))
    o = E.run("hmc", None, x0, p=p[:, None], u=u[:, None], ntraj=n, nranks=nranks, **kw)
    assert np.max(np.abs(o["s"] - g["plain_dH"])) < 1e-9 and np.array_equal(o["acc"].astype(bool), g["plain_acc"])
    assert np.array_equal(o["topo"], g["plain_topo"]) and np.max(np.abs(o["plaq"] - g["plain_plaq"])) < 1e-12
    assert np.max(np.abs(o["field"][0] - g["plain_final"])) < 1e-10
    kw = dict(beta=float(g["ft_beta"]), dt=float(g["ft_tau"]) / int(g["ft_nstep"]), nstep=int(g["ft_nstep"]))
    o = E.run("ft_hmc", g["weights"], x0, p=p[:, None], u=u[:, None], ntraj=n, nranks=nranks, **kw)
    assert np.max(np.abs(o["s"] - g["ft_dH"])) < 1e-8 and np.array_equal(o["acc"].astype(bool), g["ft_acc"])
    assert np.array_equal(o["topo"], g["ft_topo"]) and np.max(np.abs(o["plaq"] - g["ft_plaq"])) < 1e-10
    assert np.max(np.abs(o["field"][0] - g["ft_final"])) < 1e-8
    cur = x0
    for t in range(3):
        r = E.run("ft_hmc", g["weights"], cur, p=p[t][None], u=u[t:t + 1], nranks=nranks, **kw)
        assert r["s"][0] == o["s"][t] and r["acc"][0] == o["acc"][t]
        cur = r["field"]


@pytest.mark.parametrize("shape,act", [((2, 8, 8), "silu"), ((1, 8, 16), "silu"), ((1, 16, 8), "leaky_relu")])
def test_weight_gradient_matches_autograd(shape, act):
    """MODE_FT_GRAD (the flow-training gradient: input-gradient sweep + weight-gradient GEMMs, unpacked through the
    adjoint of the weight packing) against torch.autograd on the oracle, both mask orientations, constant-input terms
    of conv1 included."""
    import torch
    from oracle import fthmc_oracle as O
    B, L0, L1 = shape
    flow = O.random_flow(n_layers=6, seed=B + L0, activation=act, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()])
                                    for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    torch.manual_seed(3)
    x = torch.empty(B, 2, L0, L1).uniform_(0, 2 * np.pi)
    a_ref, g_ref = O.ft_action_weight_grad(2.5, flow, x)
    o = E.grad(raw, x.numpy(), beta=2.5, act=act)
    assert relerr(o["action"], a_ref.numpy()) < 1e-12
    assert relerr(o["force"], O.ft_force(2.5, flow, x).numpy()) < 1e-10
    g = g_ref.numpy()
    for l in range(g.shape[0]):
        assert relerr(o["grad"][l], g[l]) < 1e-9, l


def _vjp_reference(flow, x, gy, glj):
    """d/dx and d/dweights of sum_b [<gy_b, F(x_b)> + glj_b logJ_b] by torch.autograd on the oracle."""
    import torch
    from oracle import fthmc_oracle as O
    leaves = [t.requires_grad_(True) for lw in flow.layers for pair in zip(lw.w, lw.b) for t in pair]
    xl = x.clone().requires_grad_(True)
    y, lj = O.ft_flow_logJ(flow, xl)
    obj = (gy * y).sum() + (glj * lj).sum()
    g = torch.autograd.grad(obj, [xl] + leaves)
    per = len(leaves) // len(flow.layers)
    rows = [torch.cat([t.reshape(-1) for t in g[1 + i * per:1 + (i + 1) * per]]) for i in range(len(flow.layers))]
    for t in leaves:
        t.requires_grad_(False)
    return g[0], torch.stack(rows)


@pytest.mark.parametrize("shape,act", [((2, 8, 8), "silu"), ((1, 8, 16), "leaky_relu")])
def test_flow_vjp_matches_autograd(shape, act):
    """The vector-Jacobian mode of the adjoint sweep (fthmc_flow_vjp: an external d/dy seeds the sweep, the log-Jacobian terms
    carry a per-chain weight) against torch.autograd on the oracle."""
    import torch
    from oracle import fthmc_oracle as O
    B, L0, L1 = shape
    flow = O.random_flow(n_layers=6, seed=3 + L1, activation=act, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()])
                                    for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(B, 2, L0, L1, generator=gen) * 2 * np.pi
    gy = torch.randn(B, 2, L0, L1, generator=gen)
    glj = torch.randn(B, generator=gen)
    gx_ref, gw_ref = _vjp_reference(flow, x, gy, glj)
    o = E.grad(raw, x.numpy(), act=act, vjp_seed=gy.numpy(), vjp_wlj=glj.numpy())
    assert relerr(o["force"], gx_ref.numpy()) < 1e-10
    for l in range(gw_ref.shape[0]):
        assert relerr(o["grad"][l], gw_ref[l].numpy()) < 1e-9, l


@pytest.mark.parametrize("shape", [(2, 4, 4), (1, 4, 8), (1, 8, 4)])
def test_minimal_lattices(shape):
    """L = 4: a single stripe group per orientation, every neighbour access wraps onto the group itself."""
    import torch
    from oracle import fthmc_oracle as O
    B, L0, L1 = shape
    flow = O.random_flow(n_layers=8, seed=3, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()])
                                    for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    torch.manual_seed(8)
    x = torch.empty(B, 2, L0, L1).uniform_(-np.pi, np.pi)
    y, lj = O.ft_flow_logJ(flow, x)
    o = E.run("flow_fwd", raw, x.numpy())
    assert np.max(np.abs(o["field"] - y.numpy())) < 1e-12 and relerr(o["s"], lj.numpy()) < 1e-11
    assert relerr(E.run("ft_force", raw, x.numpy(), beta=2.0)["field"], O.ft_force(2.0, flow, x).numpy()) < 1e-10
    inv = E.run("flow_inv", raw, y.numpy())["field"]
    for b in range(B):                   # the reference's stop test spans the whole tensor: compare chain by chain
        assert np.max(np.abs(inv[b] - O.ft_flow_inv(flow, y[b:b + 1])[0].numpy())) < 1e-10


@pytest.mark.parametrize("name", ["ft_L16_b6", "ft_L32_b4"])
def test_tensor_core_winograd_phases(golden, name, monkeypatch):
    """FT_EMUL_MMA switches the CPU build to the warp-fragment (DMMA) form of conv2 / conv2^T, which for stripes of
    16k rows is the Winograd F(2,3) variant: same parity bars as the scalar form, and a teacher-forced trajectory."""
    monkeypatch.setenv("FT_EMUL_MMA", "1")
    g = golden(name)
    x, w, beta = g["x"], g["weights"], float(g["beta"])
    o = E.run("flow_fwd", w, x)
    assert np.max(np.abs(o["field"] - g["flow_fwd"])) < 1e-12
    assert np.max(np.abs(o["layer_logJ"] - g["layer_logJ"].T)) < 1e-11
    o = E.run("ft_force", w, x, beta=beta)
    assert relerr(o["field"], g["ft_force"]) < 1e-11
    o = E.run("flow_inv", w, g["flow_fwd"])
    assert np.max(np.abs(o["field"] - g["flow_inv_of_fwd"])) < 1e-11
    o = E.run("ft_hmc", w, g["traj_x"], beta=beta, dt=float(g["dt"]), nstep=int(g["nstep"]), p=g["traj_p"], u=g["traj_u"])
    assert np.max(np.abs(o["s"] - g["traj_dH"])) < 1e-8
    assert np.array_equal(o["acc"].astype(bool), g["traj_acc"])
    assert np.array_equal(o["topo"], g["traj_topo"])


@pytest.mark.parametrize("name,nranks", [("ft_L16_b6", 2), ("ft_L32_b4", 4)])
def test_tensor_core_winograd_phases_cluster(golden, name, nranks, monkeypatch):
    """The same with the lattice split over the ranks of a cluster: the Winograd source columns of a rank's first /
    last stripe group come from the halo buffers."""
    monkeypatch.setenv("FT_EMUL_MMA", "1")
    g = golden(name)
    x, w, beta = g["x"], g["weights"], float(g["beta"])
    o = E.run("flow_fwd", w, x, nranks=nranks)
    assert np.max(np.abs(o["field"] - g["flow_fwd"])) < 1e-12
    assert relerr(E.run("ft_force", w, x, beta=beta, nranks=nranks)["field"], g["ft_force"]) < 1e-11
    o = E.run("ft_hmc", w, g["traj_x"], beta=beta, dt=float(g["dt"]), nstep=int(g["nstep"]), p=g["traj_p"], u=g["traj_u"],
              nranks=nranks)
    assert np.max(np.abs(o["s"] - g["traj_dH"])) < 1e-8
    assert np.array_equal(o["acc"].astype(bool), g["traj_acc"]) and np.array_equal(o["topo"], g["traj_topo"])


def test_bisection_replay_equals_plain_loop(golden, tmp_path):
    """The inverse replays the reference's bisection from a Newton root and evaluates the mixture map only where its bounds do
    not decide a step.  Against a build of the same engine with the plain loop (-DFT_BISECT_REPLAY=0): bit-identical fields
    and iteration counts for tolerances from 1e-6 down to 1e-15 (where every late step takes the exact-evaluation
    fallback), in both angle conventions."""
    import subprocess
    plain = str(tmp_path / "libfthmc_emul_plain.so")
    subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                           "-DFT_BISECT_REPLAY=0", "-o", plain, E.SRC])
    g = golden("ft_L16_b6")
    y0 = g["flow_fwd"]
    cases = [(tol, conv) for tol in (1e-6, 3e-7, 1e-10, 1e-13, 1e-15) for conv in (0, 1)]
    out = {}
    saved = (E._lib, E.build)
    try:
        for name, lib in (("replay", None), ("plain", plain)):
            E._lib = None
            if lib is not None:
                E.build = lambda force=False, lib=lib: lib
            for tol, conv in cases:
                y = y0 if conv == 0 else wrap(y0)
                o = E.run("flow_inv", g["weights"], y, tol=tol, max_iter=200, conv=conv)
                out[(name, tol, conv)] = (o["field"].copy(), o["iters"].copy())
    finally:
        E._lib, E.build = None, saved[1]
    for tol, conv in cases:
        a, b = out[("replay", tol, conv)], out[("plain", tol, conv)]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (tol, conv)
    assert out[("replay", 1e-6, 0)][1].max() == 23 and out[("replay", 1e-13, 0)][1].min() > 40


@pytest.mark.parametrize("name,picks", [("ft_L16_b6_many", (0, 3, 205, 240)), ("ft_L32_b4_many", (1, 230))])
def test_headline_configs_sample(golden, name, picks):
    """The engine (the code the GPU runs, compiled for the host) on a sample of the reference's 260-trajectory runs at
    BASELINE configs 2 and 3: dH to 1e-8, decision and floored charge exact."""
    from conftest import thousand_inputs
    g = golden(name)
    L = int(g["L"])
    x, p, u = thousand_inputs(L, len(g["dH"]), seed=int(g["seed"]))
    for i in picks:
        nstep = int(g["nstep"][i])
        o = E.run("ft_hmc", g["weights"], x[i:i + 1].numpy(), beta=float(g["beta"]), dt=float(g["tau"]) / nstep, nstep=nstep,
                  p=p[i:i + 1].numpy(), u=u[i:i + 1].numpy())
        assert abs(o["s"][0] - g["dH"][i]) < 1e-8
        assert bool(o["acc"][0]) == bool(g["acc"][i]) and o["topo"][0] == g["topo"][i]
        assert abs(o["field"].sum() - g["field_sum"][i]) < 2e-6
