"""Parity of the CUDA path (through the Python mirror -> C ABI -> sm_100a kernels) with
  (a) golden vectors dumped from the real reference (tests/golden/*.npz), and
  (b) the CPU oracle on fresh seeded inputs.
Tolerances (north_star): action / force / log-det 1e-10 relative, dH 1e-8 absolute, accept/reject
and integer topological charge bit-exact."""
import numpy as np
import pytest
import torch

import fthmc_b200 as ft
from conftest import oracle_flow_from_golden, thousand_inputs
from oracle import fthmc_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy
REL = 1e-10


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def wrap(x):
    return np.remainder(x + np.pi, 2 * np.pi) - np.pi


def packed(g):
    return ft.PackedFlow(g["weights"], activation=str(g["activation"]), convention=int(g["convention"]))


# ---------------------------------------------------------------- plain stencils
def test_plain_pointwise_golden(golden):
    g = golden("plain_L8")
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8), tau=1.0, nstep=10)
    x = T(g["x0"])
    assert abs(float(ft.action(P, x)) - float(g["action"])) <= REL * abs(float(g["action"]))
    assert relerr(ft.force(P, x).numpy(), g["force"]) < REL
    assert float(ft.topocharge(x)) == float(g["topo"])
    assert np.max(np.abs(ft.regularize(T(g["reg_in"])).numpy() - g["reg_out"])) < 1e-15
    lx, lp = ft.leapfrog(P, x, T(g["lf_p"]))
    assert np.max(np.abs(lx.numpy() - g["lf_x_out"])) < 1e-12 and np.max(np.abs(lp.numpy() - g["lf_p_out"])) < 1e-12
    z = torch.zeros(2, 8, 8)
    assert float(ft.action(P, z)) / (-P.beta * P.volume) == 1.0 and float(ft.topocharge(z)) == 0.0


# (the last two shapes have rows of >= 4 KB: the column-chunked force tiles, k_force_tiled; few large chains also take the
# cluster-per-chain, software-pipelined form of the reduction scans)
@pytest.mark.parametrize("shape", [(1, 4, 4), (7, 8, 12), (5, 32, 32), (3, 64, 64), (2, 256, 128), (2, 24, 512), (1, 64, 1024)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_stencils_vs_oracle(shape, dtype):
    B, L0, L1 = shape
    gen = torch.Generator().manual_seed(B * 1000 + L0)
    x = (torch.rand(B, 2, L0, L1, generator=gen, dtype=torch.float64) * 2 - 1) * 3.0
    xd = x.to(dtype).cuda()
    P = ft.Param(beta=2.5, lat=(L0, L1))
    tol = REL if dtype == torch.float64 else 2e-5
    x64 = xd.double().cpu()
    act = torch.stack([O.action(2.5, x64[b]) for b in range(B)])
    assert relerr(ft.action(P, xd).double().cpu().numpy(), act.numpy()) < tol
    assert relerr(ft.u1_action(2.5, xd).double().cpu().numpy(), O.u1_action(2.5, x64).numpy()) < tol
    frc = torch.stack([O.force_closed_form(2.5, x64[b]) for b in range(B)])
    assert relerr(ft.force(P, xd).double().cpu().numpy(), frc.numpy()) < (tol if dtype == torch.float64 else 1e-4)
    if dtype == torch.float64:
        q = torch.stack([O.topocharge(x64[b]) for b in range(B)])
        assert np.array_equal(ft.topocharge(xd).cpu().numpy(), q.numpy())
        assert np.max(np.abs(ft.topo_charge(xd).cpu().numpy() - O.topo_charge(x64).numpy())) < 1e-9
        assert np.array_equal(ft.regularize(xd).cpu().numpy(), O.regularize(x64).numpy())


@pytest.mark.parametrize("shape", [(4096, 32, 32), (2500, 16, 24), (3000, 64, 64)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_reduction_stencils_large_batch(shape, dtype):
    """Batches of >= 2368 small lattices take the warp-per-chain form of the action / charge scans (k_action_topo_warp):
    same parity bars as the CTA-per-chain form, against the batched oracle (one torch call for the whole batch)."""
    B, L0, L1 = shape
    gen = torch.Generator().manual_seed(B + L0)
    x = (torch.rand(B, 2, L0, L1, generator=gen, dtype=torch.float64) * 2 - 1) * 3.0
    xd = x.to(dtype).cuda()
    x64 = xd.double().cpu()
    P = ft.Param(beta=2.5, lat=(L0, L1))
    tol = REL if dtype == torch.float64 else 2e-5
    ref = O.u1_action(2.5, x64).numpy()
    assert relerr(ft.u1_action(2.5, xd).double().cpu().numpy(), ref) < tol
    # hmc_2dU1's plaquette term order differs from u1_plaq's in the last ulp only
    assert relerr(ft.action(P, xd).double().cpu().numpy(), ref) < (1e-9 if dtype == torch.float64 else 2e-5)
    # the force of the whole batch (large batches of small fp32 chains go two chains per CTA), against the closed form
    # hmc_2dU1.py:104-111 differentiates: F0 = beta [sin P(n) - sin P(n - e1)], F1 = beta [sin P(n - e0) - sin P(n)]
    sp = torch.sin(x64[:, 0] - x64[:, 1] - torch.roll(x64[:, 0], -1, 2) + torch.roll(x64[:, 1], -1, 1))
    fref = 2.5 * torch.stack([sp - torch.roll(sp, 1, 2), torch.roll(sp, 1, 1) - sp], dim=1)
    assert relerr(fref[:3].numpy(), torch.stack([O.force_closed_form(2.5, x64[b]) for b in range(3)]).numpy()) < 1e-12
    assert relerr(ft.force(P, xd).double().cpu().numpy(), fref.numpy()) < (REL if dtype == torch.float64 else 1e-4)
    odd = ft.force(P, xd[:2501]).double().cpu().numpy()          # an odd number of chains: the last CTA holds one
    assert relerr(odd, fref[:2501].numpy()) < (REL if dtype == torch.float64 else 1e-4)
    if dtype == torch.float64:
        assert np.max(np.abs(ft.topo_charge(xd).cpu().numpy() - O.topo_charge(x64).numpy())) < 1e-9
        q = ft.topocharge(xd).cpu().numpy()
        assert np.array_equal(q, np.floor(0.1 + O.topo_charge(x64).numpy()))
        small = ft.topocharge(xd[:5]).cpu().numpy()          # CTA-per-chain form on the same chains
        assert np.array_equal(q[:5], small)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("scale", [40.0, 3.0e5, 4.0e6])
def test_action_on_unwrapped_links(scale, dtype):
    """HMC never wraps the links between Metropolis steps, so the plaquette angles the action scans see are unbounded: the
    scans' cosine reduces by multiples of pi up to 2^19 (2^15 in fp32) and hands larger arguments to the library.  Every
    kernel form (TMA ring, warp per chain, CTA per chain, cluster per chain) against the oracle on links of magnitude
    `scale`; one chain carries an Inf and must come back NaN, as torch.cos gives."""
    gen = torch.Generator().manual_seed(int(scale))
    for B, L0, L1 in [(2400, 32, 32), (2400, 8, 12), (2400, 8, 8), (5, 32, 32), (2, 256, 128)]:
        x = (torch.rand(B, 2, L0, L1, generator=gen, dtype=torch.float64) * 2 - 1) * scale
        xd = x.to(dtype).cuda()
        # the plaquette angle in the caller's precision and the reference's term order (field_transformation.py:118-119),
        # then an exact cosine of THAT angle summed in fp64: what the scan's cosine is measured against
        xc = xd.cpu()
        pl = ((xc[:, 0] + torch.roll(xc[:, 1], -1, 1)) - torch.roll(xc[:, 0], -1, 2)) - xc[:, 1]
        ref = (-2.5 * torch.cos(pl.double()).sum(dim=(1, 2))).numpy()
        if dtype == torch.float64:
            assert relerr(ref, O.u1_action(2.5, xc).numpy()) < 1e-12
        got = ft.u1_action(2.5, xd).double().cpu().numpy()
        bar = 2.5 * L0 * L1 * (4e-16 if dtype == torch.float64 else 3e-7)      # beta * V * (error of one cosine)
        assert np.max(np.abs(got - ref)) < bar, (B, L0, L1, np.max(np.abs(got - ref)))
    x = torch.zeros(3, 2, 32, 32, dtype=dtype)
    x[1, 0, 3, 4] = float("inf")
    a = ft.u1_action(2.5, x.cuda()).cpu()
    assert bool(torch.isnan(a[1])) and float(a[0]) == -2.5 * 1024 and float(a[2]) == -2.5 * 1024


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_regularize_is_bit_identical_without_the_division(dtype):
    """hmc_2dU1.regularize (hmc_2dU1.py:127-129) divides by 2 pi; the kernels form that quotient by a multiply and two fused
    multiply-adds (Markstein's correction), which must give the correctly rounded quotient, i.e. the SAME bits as torch's
    division, on every argument: 2^24 random values per magnitude class, in the caller's precision."""
    gen = torch.Generator().manual_seed(77)
    for scale in (3.2, 40.0, 1.0e4, 1.0e-3):
        f = ((torch.rand(4096, 2, 32, 64, generator=gen, dtype=torch.float64) * 2 - 1) * scale).to(dtype)
        pi, tp = torch.tensor(np.pi, dtype=dtype), torch.tensor(2 * np.pi, dtype=dtype)
        g = (f - pi) / tp
        ref = tp * (g - torch.floor(g) - 0.5)
        got = ft.regularize(f.cuda()).cpu()
        assert torch.equal(got, ref), (dtype, scale, float((got - ref).abs().max()))


# ---------------------------------------------------------------- plain HMC
def test_plain_hmc_teacher_forced_golden(golden):
    g = golden("plain_L8")
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8), tau=1.0, nstep=int(g["nstep"]))
    r = ft.hmc_batch(P, T(g["traj_x"]), T(g["traj_p"]), T(g["traj_u"]))
    assert np.max(np.abs(r["dH"].numpy() - g["traj_dH"])) < 1e-8
    assert np.array_equal(r["acc"].numpy(), g["traj_acc"])
    assert np.array_equal(r["topo"].numpy(), g["traj_topo"])
    assert np.max(np.abs(r["field"].numpy() - g["traj_out"])) < 1e-10
    assert np.max(np.abs(r["plaq"].numpy() - g["traj_plaq"])) < 1e-12


def test_plain_hmc_dropin_reproduces_reference_chain(golden):
    """hmc(param,x) draws p,u from torch's generator in the reference's order: same seed, same chain."""
    g = golden("plain_L8")
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8), tau=1.0, nstep=int(g["nstep"]))
    cur = T(g["traj_x"][0]).clone()
    for n in range(12):
        torch.manual_seed(5000 + n)
        dH, e, acc, cur = ft.hmc(P, cur)
        assert abs(float(dH) - g["traj_dH"][n]) < 1e-8 and bool(acc) == bool(g["traj_acc"][n])
        assert abs(float(e) - np.exp(-g["traj_dH"][n])) < 1e-8
    assert np.max(np.abs(cur.numpy() - g["traj_out"][11])) < 1e-9


def test_thousand_trajectories_plain(golden):
    g = golden("plain_L8_1000")
    x, p, u = thousand_inputs(8)
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8), tau=1.0, nstep=int(g["nstep"]))
    r = ft.hmc_batch(P, x, p, u)
    assert np.max(np.abs(r["dH"].numpy() - g["dH"])) < 1e-8
    assert np.array_equal(r["acc"].numpy(), g["acc"])
    assert np.array_equal(r["topo"].numpy(), g["topo"])
    assert np.max(np.abs(r["field"].sum(dim=(1, 2, 3)).numpy() - g["field_sum"])) < 1e-9


# ---------------------------------------------------------------- flow
@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4"])
def test_flow_forward_action_force_golden(golden, name):
    g = golden(name)
    pf = packed(g)
    L = int(g["L"])
    P = ft.Param(beta=float(g["beta"]), lat=(L, L))
    x = T(g["x"])
    y, lj = ft.ft_flow(pf, x, with_logJ=True)
    assert np.max(np.abs(y.numpy() - g["flow_fwd"])) < 1e-11
    assert relerr(lj.numpy(), g["layer_logJ"].sum(axis=0)) < REL
    assert relerr(ft.ft_action(P, pf, x).numpy(), g["ft_action"]) < REL
    assert relerr(ft.ft_force(P, pf, x).numpy(), g["ft_force"]) < REL


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4"])
def test_flow_inverse_golden(golden, name):
    g = golden(name)
    pf = packed(g)
    xi = ft.ft_flow_inv(pf, T(g["flow_fwd"]))
    # same bisection decisions => the same dyadic midpoints (agreement far below the 1e-6 tolerance)
    assert np.max(np.abs(xi.numpy() - g["flow_inv_of_fwd"])) < 1e-10
    assert np.max(np.abs(wrap(xi.numpy() - g["x"]))) < 2e-5


@pytest.mark.parametrize("name", ["leaky_L8", "copyB_L8"])
def test_variants_golden(golden, name):
    g = golden(name)
    pf = packed(g)
    y, lj = ft.ft_flow(pf, T(g["x"]), with_logJ=True)
    assert np.max(np.abs(y.numpy() - g["flow_fwd"])) < 1e-11 and relerr(lj.numpy(), g["logJ"]) < REL
    xi = ft.ft_flow_inv(pf, T(g["flow_fwd"]))
    assert np.max(np.abs(xi.numpy() - g["flow_inv_of_fwd"])) < 1e-10


@pytest.mark.parametrize("name", ["ft_L8_n8", "ft_L16_b6", "ft_L32_b4", "ft_L32_b4_n40"])
def test_ft_hmc_teacher_forced_golden(golden, name):
    g = golden(name)
    pf = packed(g)
    L = int(g["L"])
    P = ft.Param(beta=float(g["beta"]), lat=(L, L), tau=float(g["dt"]) * int(g["nstep"]), nstep=int(g["nstep"]))
    r = ft.ft_hmc_batch(P, pf, T(g["traj_x"]), T(g["traj_p"]), T(g["traj_u"]))
    assert np.max(np.abs(r["dH"].numpy() - g["traj_dH"])) < 1e-8
    assert np.array_equal(r["acc"].numpy(), g["traj_acc"])
    assert np.array_equal(r["topo"].numpy(), g["traj_topo"])
    assert np.max(np.abs(r["plaq"].numpy() - g["traj_plaq"])) < 1e-11
    assert np.max(np.abs(r["field"].numpy() - g["traj_out"])) < 1e-8


def test_thousand_trajectories_ft(golden):
    g = golden("ft_L8_1000")
    pf = packed(g)
    x, p, u = thousand_inputs(8)
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8), tau=float(g["dt"]) * int(g["nstep"]), nstep=int(g["nstep"]))
    r = ft.ft_hmc_batch(P, pf, x, p, u)
    assert np.max(np.abs(r["dH"].numpy() - g["dH"])) < 1e-8
    assert np.array_equal(r["acc"].numpy(), g["acc"])
    assert np.array_equal(r["topo"].numpy(), g["topo"])
    assert np.max(np.abs(r["field"].sum(dim=(1, 2, 3)).numpy() - g["field_sum"])) < 1e-7


@pytest.mark.parametrize("name", ["ft_L16_b6_many", "ft_L32_b4_many"])
def test_headline_configs_many_trajectories(golden, name):
    """BASELINE configs 2 (L=16, beta=6) and 3 (L=32, beta=4) at depth: 260 teacher-forced trajectories of the REFERENCE
    each (24-layer seed-3647 flow; 200 at nstep=40, where about half are accepted, the rest at nstep=20 and at the bench's
    nstep=10), every one from its own field / momentum / uniform.  dH to 1e-8, accept/reject and floored charge bit-exact."""
    g = golden(name)
    pf = packed(g)
    L, n = int(g["L"]), len(g["dH"])
    x, p, u = thousand_inputs(L, n, seed=int(g["seed"]))
    acc = g["acc"].astype(bool)
    assert 0.25 < acc[g["nstep"] == 40].mean() < 0.75          # the decision check is not vacuous
    assert len(np.unique(g["topo"])) > 3
    for nstep in np.unique(g["nstep"]):
        sel = np.nonzero(g["nstep"] == nstep)[0]
        P = ft.Param(beta=float(g["beta"]), lat=(L, L), tau=float(g["tau"]), nstep=int(nstep))
        r = ft.ft_hmc_batch(P, pf, x[sel], p[sel], u[sel])
        assert np.max(np.abs(r["dH"].numpy() - g["dH"][sel])) < 1e-8, nstep
        assert np.array_equal(r["acc"].numpy(), acc[sel]), nstep
        assert np.array_equal(r["topo"].numpy(), g["topo"][sel]), nstep
        assert np.max(np.abs(r["field"].sum(dim=(1, 2, 3)).numpy() - g["field_sum"][sel])) < 2e-6, nstep


def test_config4_L128_reference_golden(golden):
    """BASELINE config 4 pinned by the reference itself: L=128, beta=6, the L=16 flow transferred with the reference's
    flow_resize (ipynb/ft_hmc.py:489-513), on the 16-CTA cluster path: ft_action, ft_force, the flowed field and one
    teacher-forced trajectory (nstep=40)."""
    g = golden("ft_L128_b6")
    pf = ft.flow_resize(packed(g), (128, 128))
    x, p, u = thousand_inputs(128, 1, seed=int(g["seed"]))
    P = ft.Param(beta=6.0, lat=(128, 128), tau=float(g["dt"]) * int(g["nstep"]), nstep=int(g["nstep"]))
    assert relerr(ft.ft_action(P, pf, x).numpy(), g["ft_action"]) < REL
    f = ft.ft_force(P, pf, x).numpy()
    assert relerr(np.array([f.sum(), np.abs(f).sum(), (f * f).sum()])[1:], g["ft_force_sum"][1:]) < REL
    assert abs(f.sum() - g["ft_force_sum"][0]) < 1e-9 * g["ft_force_sum"][1]
    assert relerr(f[0, :, 0, :], g["ft_force_row0"]) < REL and relerr(f[0, :, :, 5], g["ft_force_col5"]) < REL
    y = ft.ft_flow(pf, x)
    assert abs(float((y * y).sum()) - g["flow_fwd_sum"][1]) < 1e-10 * g["flow_fwd_sum"][1]
    yw = torch.remainder(y + np.pi, 2 * np.pi) - np.pi
    r = ft.ft_hmc_batch(P, pf, yw, p, u)
    assert abs(float(r["dH"][0]) - float(g["traj_dH"])) < 1e-8
    assert bool(r["acc"][0]) == bool(g["traj_acc"]) and float(r["topo"][0]) == float(g["traj_topo"])
    assert abs(float(r["plaq"][0]) - float(g["traj_plaq"])) < 1e-11
    assert abs(float(r["field"].sum()) - float(g["traj_field_sum"])) < 1e-5


def test_ft_hmc_dropin_signature(golden):
    """ft_hmc(param, flow, field) on the reference's own kind of flow object (an nn.ModuleList-like
    container exposing layer.plaq_coupling.net), CPU tensors in, CPU tensors out."""
    g = golden("ft_L8_n8")
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8), tau=float(g["dt"]) * int(g["nstep"]), nstep=int(g["nstep"]))
    flow = module_like(g)
    torch.manual_seed(7000)
    dH, e, acc, new = ft.ft_hmc(P, flow, T(g["traj_x"][0][None]))
    assert isinstance(dH, float) and isinstance(e, float) and acc.dtype == torch.bool and new.shape == (1, 2, 8, 8)
    assert abs(dH - g["traj_dH"][0]) < 1e-8 and bool(acc) == bool(g["traj_acc"][0])
    assert np.max(np.abs(new.numpy()[0] - g["traj_out"][0])) < 1e-8


def module_like(g):
    """Stand-in for the reference's ModuleList of GaugeEquivCouplingLayer (same attribute names)."""
    import torch.nn as nn
    shapes = [(8, 2, 3, 3), (8,), (8, 8, 3, 3), (8,), (3, 8, 3, 3), (3,)]

    class PlaqCoupling(nn.Module):
        def __init__(self, net):
            super().__init__()
            self.net, self.inv_prec, self.inv_max_iter = net, 1e-6, 1000

    class Layer(nn.Module):
        def __init__(self, net):
            super().__init__()
            self.plaq_coupling = PlaqCoupling(net)

    layers = []
    for row in g["weights"]:
        convs = [nn.Conv2d(2, 8, 3, padding=1, padding_mode="circular"), nn.Conv2d(8, 8, 3, padding=1, padding_mode="circular"),
                 nn.Conv2d(8, 3, 3, padding=1, padding_mode="circular")]
        pos = 0
        with torch.no_grad():
            for c, (ws, bs) in zip(convs, zip(shapes[0::2], shapes[1::2])):
                n = int(np.prod(ws)); c.weight.copy_(T(row[pos:pos + n].reshape(ws))); pos += n
                n = int(np.prod(bs)); c.bias.copy_(T(row[pos:pos + n].reshape(bs))); pos += n
        layers.append(Layer(nn.Sequential(convs[0], nn.SiLU(), convs[1], nn.SiLU(), convs[2])))
    return nn.ModuleList(layers).double()


# ---------------------------------------------------------------- fresh inputs vs the oracle
@pytest.mark.parametrize("shape", [(2, 4, 4), (1, 4, 8), (3, 8, 8), (2, 8, 12), (2, 16, 16), (1, 12, 20), (1, 24, 40)])
@pytest.mark.parametrize("act", ["silu", "leaky_relu", "relu"])
def test_flow_vs_oracle_fresh(shape, act):
    B, L0, L1 = shape
    flow = O.random_flow(n_layers=8, seed=B + L0 + L1, activation=act, scale=2.5)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)])
                    for lw in flow.layers])
    pf = ft.PackedFlow(raw, activation=act)
    gen = torch.Generator().manual_seed(17)
    x = (torch.rand(B, 2, L0, L1, generator=gen, dtype=torch.float64) * 2 - 1) * 4.0     # un-wrapped angles
    P = ft.Param(beta=3.0, lat=(L0, L1))
    y, lj = O.ft_flow_logJ(flow, x)
    yg, ljg = ft.ft_flow(pf, x.cuda(), with_logJ=True)
    assert yg.is_cuda and np.max(np.abs(yg.cpu().numpy() - y.numpy())) < 1e-11
    assert relerr(ljg.cpu().numpy(), lj.numpy()) < REL
    assert relerr(ft.ft_action(P, pf, x).numpy(), O.ft_action(3.0, flow, x).numpy()) < REL
    assert relerr(ft.ft_force(P, pf, x).numpy(), O.ft_force(3.0, flow, x).numpy()) < REL
    for b in range(B):
        xi = O.ft_flow_inv(flow, y[b:b + 1])
        assert np.max(np.abs(ft.ft_flow_inv(pf, y[b:b + 1]).numpy() - xi.numpy())) < 1e-10


def test_ft_leapfrog_vs_oracle():
    flow = O.random_flow(n_layers=8, seed=5, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)])
                    for lw in flow.layers])
    pf = ft.PackedFlow(raw)
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(2, 2, 8, 8, generator=gen, dtype=torch.float64) * 6 - 3
    p = torch.randn(2, 2, 8, 8, generator=gen, dtype=torch.float64)
    P = ft.Param(beta=2.0, lat=(8, 8), tau=0.5, nstep=5)
    xo, po = O.ft_leapfrog(2.0, P.dt, 5, flow, x, p)
    xg, pg = ft.ft_leapfrog(P, pf, x, p)
    assert np.max(np.abs(xg.numpy() - xo.numpy())) < 1e-10 and np.max(np.abs(pg.numpy() - po.numpy())) < 1e-10


# ---------------------------------------------------------------- size-independent properties at full size
def test_full_size_properties_L32():
    """BASELINE config 3 shape (L=32, beta=4, 24 layers) at B=296 (two waves of the persistent grid):
    round trip to the bisection tolerance, logJ antisymmetry, gauge invariance, chain independence of
    batching, Philox determinism, reversibility of the integrator."""
    flow = O.random_flow(n_layers=24, seed=3647)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)])
                    for lw in flow.layers])
    pf = ft.PackedFlow(raw)
    B, L = 296, 32
    gen = torch.Generator().manual_seed(1331)
    x = ((torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    P = ft.Param(beta=4.0, lat=(L, L), tau=1.0, nstep=10)
    y, lj = ft.ft_flow(pf, x, with_logJ=True)
    xi, lji = ft.ft_flow_inv(pf, y, with_logJ=True)
    assert float(torch.max(torch.abs(torch.remainder(xi - x + np.pi, 2 * np.pi) - np.pi))) < 5e-5
    assert float(torch.max(torch.abs(lj + lji))) < 1e-3
    # gauge invariance of action / charge / logJ, equivariance of the flow
    alpha = torch.rand(B, L, L, generator=gen, dtype=torch.float64).cuda() * 2 * np.pi
    xg = x.clone()
    xg[:, 0] = alpha + x[:, 0] - torch.roll(alpha, -1, 1)
    xg[:, 1] = alpha + x[:, 1] - torch.roll(alpha, -1, 2)
    assert float(torch.max(torch.abs(ft.u1_action(4.0, x) - ft.u1_action(4.0, xg)))) < 1e-8
    assert torch.equal(ft.topocharge(x), ft.topocharge(xg))
    yg, ljg = ft.ft_flow(pf, xg, with_logJ=True)
    assert float(torch.max(torch.abs(lj - ljg))) < 1e-8
    assert float(torch.max(torch.abs(ft.ft_action(P, pf, x) - ft.ft_action(P, pf, xg)))) < 1e-7
    # a chain's result does not depend on which CTA / batch position ran it
    f_all = ft.ft_force(P, pf, x)
    f_one = ft.ft_force(P, pf, x[200:201])
    assert torch.equal(f_all[200:201], f_one)
    # oracle spot-check of one chain at full size
    fo = O.ft_force(4.0, flow, x[5:6].cpu())
    assert relerr(f_all[5:6].cpu().numpy(), fo.numpy()) < REL
    # Philox throughput mode: deterministic in (seed, chain, traj), different across chains
    r1 = ft.ft_hmc_batch(P, pf, x[:150], seed=11, traj=3)
    r2 = ft.ft_hmc_batch(P, pf, x[:150], seed=11, traj=3)
    assert torch.equal(r1["field"], r2["field"]) and torch.equal(r1["dH"], r2["dH"])
    r3 = ft.ft_hmc_batch(P, pf, x[100:150], seed=11, traj=3, chain0=100)
    assert torch.equal(r1["field"][100:150], r3["field"]) and torch.equal(r1["acc"][100:150], r3["acc"])
    assert float(torch.std(r1["dH"])) > 0
    q = r1["topo"]
    assert torch.equal(q, torch.round(q)) and torch.equal(q, ft.topocharge(r1["field"]))
    # leapfrog reversibility: integrate, flip momenta, integrate back
    p = torch.randn(8, 2, L, L, generator=gen, dtype=torch.float64).cuda()
    xa, pa = ft.ft_leapfrog(P, pf, x[:8], p)
    xb, pb = ft.ft_leapfrog(P, pf, xa, -pa)
    assert float(torch.max(torch.abs(xb - x[:8]))) < 1e-9 and float(torch.max(torch.abs(pb + p))) < 1e-9


# ---------------------------------------------------------------- cluster / DSMEM path (lattices beyond one SM)
def _raw_of(flow):
    return np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)])
                     for lw in flow.layers])


@pytest.mark.parametrize("L,layers,B", [(64, 8, 3), (48, 4, 2), (128, 24, 2)])
def test_cluster_path_vs_oracle(L, layers, B):
    """BASELINE config 4 (L=128, 24 layers; one 16-CTA cluster per chain) and smaller cluster sizes against the oracle:
    flow, log-det, ft_action, ft_force to 1e-10, inverse decision-for-decision, a teacher-forced trajectory."""
    flow = O.random_flow(n_layers=layers, seed=3647)
    pf = ft.PackedFlow(_raw_of(flow))
    gen = torch.Generator().manual_seed(L)
    x = (torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi
    P = ft.Param(beta=6.0, lat=(L, L), tau=0.3, nstep=3)
    y, lj = O.ft_flow_logJ(flow, x)
    yg, ljg = ft.ft_flow(pf, x.cuda(), with_logJ=True)
    assert np.max(np.abs(yg.cpu().numpy() - y.numpy())) < 1e-11
    assert relerr(ljg.cpu().numpy(), lj.numpy()) < REL
    assert relerr(ft.ft_action(P, pf, x).numpy(), O.ft_action(6.0, flow, x).numpy()) < REL
    assert relerr(ft.ft_force(P, pf, x).numpy(), O.ft_force(6.0, flow, x).numpy()) < REL
    xi = O.ft_flow_inv(flow, y[:1])
    assert np.max(np.abs(ft.ft_flow_inv(pf, y[:1]).numpy() - xi.numpy())) < 1e-10
    p = torch.randn(1, 2, L, L, generator=gen, dtype=torch.float64)
    u = torch.rand(1, generator=gen, dtype=torch.float64)
    dH, e, acc, new = O.ft_hmc(6.0, P.dt, P.nstep, flow, y[:1], p=p, u=u[0])
    r = ft.ft_hmc_batch(P, pf, y[:1], p, u)
    assert abs(float(r["dH"][0]) - dH) < 1e-8 and bool(r["acc"][0]) == bool(acc)
    assert np.max(np.abs(r["field"].numpy() - new.numpy())) < 1e-8
    assert float(r["topo"][0]) == float(O.topocharge(new[0]))
    # plain HMC on the same lattice (cluster for L=128) and batch > resident clusters
    r1 = ft.hmc_batch(P, x, seed=3, traj=1)
    xs = torch.cat([x] * 6)
    r2 = ft.hmc_batch(P, xs, seed=3, traj=1)
    assert torch.equal(r2["field"][:B], r1["field"])
    xo, po = O.leapfrog(6.0, P.dt, P.nstep, x[0], p[0])
    xg, pg = ft.leapfrog(P, x[0], p[0])
    assert np.max(np.abs(xg.numpy() - xo.numpy())) < 1e-11 and np.max(np.abs(pg.numpy() - po.numpy())) < 1e-11


# ---------------------------------------------------------------- run loops (many trajectories per launch)
def test_run_loops_reproduce_reference_chain(golden):
    """run / ft_run seeded once: the reference's free-running chain (12 trajectories, accepts and rejects), every
    printed observable, from ONE kernel launch per nrun block with the chain resident in shared memory."""
    import io
    g = golden("run_L8")
    n, x0 = int(g["ntraj"]), T(g["x0"])
    P = ft.Param(beta=float(g["plain_beta"]), lat=(8, 8), tau=float(g["plain_tau"]), nstep=int(g["plain_nstep"]), ntraj=n // 2, nrun=2)
    torch.manual_seed(int(g["seed"]))
    buf = io.StringIO()
    f = ft.run(P, x0.clone(), out=buf)
    assert np.max(np.abs(f.numpy() - g["plain_final"])) < 1e-10
    assert np.array_equal(np.array(ft.topo_history), g["plain_topo"])
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("Traj:")]
    assert len(lines) == n and [("ACCEPT" in l) for l in lines] == list(g["plain_acc"])
    flow = module_like(g)
    Pf = ft.Param(beta=float(g["ft_beta"]), lat=(8, 8), tau=float(g["ft_tau"]), nstep=int(g["ft_nstep"]), ntraj=n, nrun=1)
    torch.manual_seed(int(g["seed"]))
    f = ft.ft_run(Pf, flow, x0.clone())
    assert np.max(np.abs(f.numpy() - g["ft_final"])) < 1e-8
    assert np.array_equal(np.array(ft.topo_history), g["ft_topo"])


def test_run_batch_equals_trajectory_by_trajectory():
    """ft_hmc_run_batch(ntraj) == ntraj calls of ft_hmc_batch (bit for bit), explicit momenta and device-RNG mode,
    single-CTA (L=16) and cluster (L=64) paths; same for plain HMC."""
    flow = O.random_flow(n_layers=6, seed=4, scale=2.0)
    pf = ft.PackedFlow(_raw_of(flow))
    for L, B in ((16, 5), (64, 2)):
        gen = torch.Generator().manual_seed(L)
        x = ((torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
        P = ft.Param(beta=3.0, lat=(L, L), tau=0.4, nstep=3)
        n = 3
        r = ft.ft_hmc_run_batch(P, pf, x, n, seed=9, traj0=5, chain0=2)
        cur = x
        for t in range(n):
            s1 = ft.ft_hmc_batch(P, pf, cur, seed=9, traj=5 + t, chain0=2)
            assert torch.equal(s1["dH"], r["dH"][t]) and torch.equal(s1["acc"], r["acc"][t])
            assert torch.equal(s1["topo"], r["topo"][t]) and torch.equal(s1["plaq"], r["plaq"][t])
            cur = s1["field"]
        assert torch.equal(cur, r["field"])
        p = torch.randn(n, B, 2, L, L, generator=gen, dtype=torch.float64).cuda()
        u = torch.rand(n, B, generator=gen, dtype=torch.float64).cuda()
        r = ft.hmc_run_batch(P, x, n, p, u)
        cur = x
        for t in range(n):
            s1 = ft.hmc_batch(P, cur, p[t], u[t])
            assert torch.equal(s1["dH"], r["dH"][t]) and torch.equal(s1["acc"], r["acc"][t])
            cur = s1["field"]
        assert torch.equal(cur, r["field"])
        assert r["dH"].shape == (n, B) and r["acc"].dtype == torch.bool


def test_checkpoint_load_and_lattice_transfer(golden, tmp_path):
    """A checkpoint written in the reference's format drives the kernels; the packed flow of an L=16 run applies
    unchanged to L=32 and L=128 (flow_resize / transfer_to_new_lattice: same CNN weights, masks for the new lattice)."""
    g16, g32 = golden("ft_L16_b6"), golden("ft_L32_b4")
    flow16 = module_like(g16)
    fn = tmp_path / "ckpt-era0-epoch0.tar"
    torch.save({"era": 0, "epoch": 0, "model_state_dict": flow16.state_dict(), "optimizer_state_dict": {}}, fn)
    pf = ft.load_flow(str(fn))
    assert np.max(np.abs(ft.ft_flow(pf, T(g16["x"])).numpy() - g16["flow_fwd"])) < 1e-12
    big = ft.flow_resize(flow16, (32, 32))                      # the reference's flow at L=32 has the same weights
    assert np.array_equal(g16["weights"], g32["weights"])
    assert np.max(np.abs(ft.ft_flow(big, T(g32["x"])).numpy() - g32["flow_fwd"])) < 1e-12
    assert relerr(ft.ft_force(ft.Param(beta=float(g32["beta"]), lat=(32, 32)), big, T(g32["x"])).numpy(), g32["ft_force"]) < REL
    x = (torch.rand(1, 2, 128, 128, dtype=torch.float64) * 2 - 1) * np.pi
    y, lj = ft.ft_flow(pf, x, with_logJ=True)
    xi, lji = ft.ft_flow_inv(pf, y, with_logJ=True)
    assert float(torch.max(torch.abs(torch.remainder(xi - x + np.pi, 2 * np.pi) - np.pi))) < 5e-5 and abs(float(lj + lji)) < 1e-2


def test_load_flow_pickled_modulelist(golden, tmp_path):
    """The reference's own way of keeping a trained flow, `torch.save(flow, "flow_b{beta}_l{L}x{L}.dat")` (a pickled
    ModuleList, ipynb/ft_hmc.py:356-373): load_flow unpickles it (its classes importable, as for the reference) and packs it."""
    import importlib, sys, textwrap
    (tmp_path / "ref_like_layers.py").write_text(textwrap.dedent("""
        import torch.nn as nn
        class PlaqCoupling(nn.Module):
            def __init__(self, net):
                super().__init__()
                self.net, self.inv_prec, self.inv_max_iter = net, 1e-6, 1000
        class Layer(nn.Module):
            def __init__(self, net):
                super().__init__()
                self.plaq_coupling = PlaqCoupling(net)
        def make(weights):
            import torch
            shapes = [(8, 2, 3, 3), (8,), (8, 8, 3, 3), (8,), (3, 8, 3, 3), (3,)]
            layers = []
            for row in weights:
                convs = [nn.Conv2d(2, 8, 3, padding=1, padding_mode="circular"), nn.Conv2d(8, 8, 3, padding=1, padding_mode="circular"),
                         nn.Conv2d(8, 3, 3, padding=1, padding_mode="circular")]
                pos = 0
                for c in convs:
                    for prm in (c.weight, c.bias):
                        n = prm.numel()
                        prm.data = torch.from_numpy(row[pos:pos + n].reshape(tuple(prm.shape)).copy()); pos += n
                layers.append(Layer(nn.Sequential(convs[0], nn.SiLU(), convs[1], nn.SiLU(), convs[2])))
            return nn.ModuleList(layers)
    """))
    sys.path.insert(0, str(tmp_path))
    try:
        mod = importlib.import_module("ref_like_layers")
        g = golden("ft_L8_n8")
        fn = tmp_path / "flow_b2.0_l8x8.dat"
        torch.save(mod.make(g["weights"]), fn)
        pf = ft.load_flow(str(fn))
        assert pf.n_layers == int(g["n_layers"])
        assert np.max(np.abs(ft.ft_flow(pf, T(g["x"])).numpy() - g["flow_fwd"])) < 1e-12
    finally:
        sys.path.remove(str(tmp_path))
        sys.modules.pop("ref_like_layers", None)


# ---------------------------------------------------------------- flow training gradient
@pytest.mark.parametrize("L,layers,B", [(8, 6, 5), (16, 8, 3), (32, 24, 2)])
def test_weight_gradient_vs_autograd(L, layers, B):
    """fthmc_ft_action_grad (weight-gradient GEMMs on the tensor path + host-side unpacking) against torch.autograd on
    the oracle: the gradient the reference's reverse-KL train_step back-propagates."""
    flow = O.random_flow(n_layers=layers, seed=L, scale=1.5 if layers < 24 else 1.0)
    pf = ft.PackedFlow(_raw_of(flow))
    gen = torch.Generator().manual_seed(2 * L)
    x = torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 * np.pi
    P = ft.Param(beta=3.0, lat=(L, L))
    a_ref, g_ref = O.ft_action_weight_grad(3.0, flow, x)
    act, g, frc = ft.ft_action_grad(P, pf, x, want_force=True)
    assert relerr(act.numpy(), a_ref.numpy()) < REL
    assert relerr(frc.numpy(), O.ft_force(3.0, flow, x).numpy()) < REL
    for l in range(layers):
        assert relerr(g[l].numpy(), g_ref[l].numpy()) < 1e-9, l
    # more chains than resident CTAs: the per-CTA accumulators add up over the persistent loop
    xb = torch.cat([x] * 80)[:200] if L == 8 else x
    if L == 8:
        act2, g2 = ft.ft_action_grad(P, pf, xb)
        reps = torch.tensor([(200 - i + B - 1) // B for i in range(B)], dtype=torch.float64)      # copies of each chain
        a3, g3 = O.ft_action_weight_grad(3.0, flow, x)
        want = sum(float(reps[i]) * O.ft_action_weight_grad(3.0, flow, x[i:i + 1])[1] for i in range(B))
        assert relerr(g2.numpy(), want.numpy()) < 1e-9


def test_train_step_follows_autograd_adam():
    """FlowTrainer.train_step == the reference's step (reverse-KL loss, loss.backward(), Adam) done with autograd on the
    oracle, for the same prior batch; a few steps reduce the loss."""
    L, layers = 8, 4
    raw0 = ft.default_init_raw(layers, 11)
    tr = ft.FlowTrainer(raw0, (L, L), beta=2.0, lr=1e-3, seed=5)
    xi = tr.sample_prior(16).cpu()
    # autograd twin
    flow = oracle_flow_from_golden(dict(weights=raw0, activation="silu", convention=0))
    params = [t.requires_grad_(True) for lw in flow.layers for pair in zip(lw.w, lw.b) for t in pair]
    opt = torch.optim.Adam(params, lr=1e-3)
    loss = torch.mean(O.ft_action(2.0, flow, xi)) + tr.log_prior()
    opt.zero_grad(); loss.backward(); opt.step()
    m = tr.train_step(16, xi=xi)
    assert abs(m["dkl"] - float(loss.detach())) < 1e-9 * abs(float(loss.detach()))
    twin = np.stack([np.concatenate([np.concatenate([w.detach().numpy().ravel(), b.detach().numpy().ravel()])
                                     for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    assert np.max(np.abs(tr.raw.detach().numpy() - twin)) < 1e-9
    first = m["dkl"]
    for _ in range(30):
        m = tr.train_step(64)
    assert np.mean(tr.history["dkl"][-5:]) < first and 0 < m["ess"] <= 1.0


def _force_norm_autograd(beta, flow, xi):
    """loss = sum |ft_force(xi)|^2 and d loss / d weights the reference's way: ft_force(..., create_graph=True), backward."""
    leaves = [t.requires_grad_(True) for lw in flow.layers for pair in zip(lw.w, lw.b) for t in pair]
    x = xi.clone().requires_grad_(True)
    f, = torch.autograd.grad(O.ft_action(beta, flow, x).sum(), x, create_graph=True)
    loss = (f * f).sum()
    g = torch.autograd.grad(loss, leaves)
    per = len(leaves) // len(flow.layers)
    rows = [torch.cat([t.reshape(-1) for t in g[i * per:(i + 1) * per]]) for i in range(len(flow.layers))]
    for t in leaves:
        t.requires_grad_(False)
    return float(loss.detach()), torch.stack(rows)


@pytest.mark.parametrize("L,layers,B,scale", [(8, 8, 4, 2.0), (16, 24, 2, 1.0)])
def test_force_norm_gradient_vs_autograd(L, layers, B, scale):
    """ft_force_norm_grad (the reference's second-stage loss, ipynb/ft_hmc.py:266-269, 367) against
    torch.autograd.grad(..., create_graph=True) on the oracle: loss and d loss / d weights to 1e-8."""
    flow = O.random_flow(n_layers=layers, seed=3647 + L, scale=scale)
    pf = ft.PackedFlow(_raw_of(flow))
    gen = torch.Generator().manual_seed(L)
    xi = torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 * np.pi
    P = ft.Param(beta=2.0, lat=(L, L))
    loss_ref, g_ref = _force_norm_autograd(2.0, flow, xi)
    loss, g, F = ft.ft_force_norm_grad(P, pf, xi)
    assert abs(float(loss) - loss_ref) < 1e-10 * loss_ref
    assert relerr(g.numpy(), g_ref.numpy()) < 1e-8
    for l in range(layers):
        assert relerr(g[l].numpy(), g_ref[l].numpy()) < 1e-7, l          # (layer by layer: small-gradient layers included)
    with pytest.raises(NotImplementedError, match="ft_force_norm_grad"):
        ft.ft_force(P, pf, xi, create_graph=True)


def test_train_step_with_force_and_pre_model():
    """train_step(with_force=True, pre_model=...) (ipynb/ft_hmc.py:253-276): xi = F^-1(F_pre(xi_pre)) held fixed, loss =
    sum |ft_force(xi)|^2, Adam step -- against the same step done with autograd on the oracle from the same latent batch."""
    L, layers = 8, 4
    raw_pre, raw0 = ft.default_init_raw(layers, 21), ft.default_init_raw(layers, 11)
    pre = ft.FlowTrainer(raw_pre, (L, L), beta=2.0, seed=1)
    tr = ft.FlowTrainer(raw0, (L, L), beta=2.0, lr=1e-5, seed=5)
    xi = ft.ft_flow_inv(tr.packed(), ft.ft_flow(pre.packed(), tr.sample_prior(8))).cpu()
    flow = oracle_flow_from_golden(dict(weights=raw0, activation="silu", convention=0))
    loss_ref, g_ref = _force_norm_autograd(2.0, flow, xi)
    params = [t.requires_grad_(True) for lw in flow.layers for pair in zip(lw.w, lw.b) for t in pair]
    opt = torch.optim.Adam(params, lr=1e-5)
    opt.zero_grad()
    pos = 0
    for lw, row in zip(flow.layers, g_ref):
        for w, b in zip(lw.w, lw.b):
            for t in (w, b):
                t.grad = row[pos:pos + t.numel()].reshape(t.shape).clone(); pos += t.numel()
        pos = 0
    opt.step()
    m = tr.train_step(8, xi=xi, with_force=True)
    assert abs(m["force"] - loss_ref) < 1e-9 * loss_ref and m["loss"] == m["force"]
    twin = np.stack([np.concatenate([np.concatenate([w.detach().numpy().ravel(), b.detach().numpy().ravel()])
                                     for w, b in zip(lw.w, lw.b)]) for lw in flow.layers])
    assert np.max(np.abs(tr.raw.detach().numpy() - twin)) < 1e-9
    # the sampling path through a pre-trained flow, both losses
    m1 = tr.train_step(16, pre_model=pre)
    m2 = tr.train_step(16, pre_model=pre, with_force=True)
    assert np.isfinite(m1["dkl"]) and m2["force"] > 0 and len(tr.history["force"]) == 3


def test_flow_train_and_eval_drivers():
    """flow_train / flow_eval (ipynb/ft_hmc.py:297-354) on the CUDA entry points: a short two-stage training run at L=8
    (reverse-KL, then reverse-KL + force-norm steps through the first stage's flow) lowers the loss, and the
    independence-Metropolis evaluation of the trained flow returns an accept rate and <Q^2> with an error."""
    import io
    out = io.StringIO()
    pre = ft.flow_train((8, 8), 2.0, n_layers=4, n_era=2, n_epoch=20, batch_size=64, base_lr=1e-3, seed=11, out=out)
    assert np.mean(pre.history["dkl"][-5:]) < np.mean(pre.history["dkl"][:5]) and out.getvalue().count("== Era") == 2
    tr = ft.flow_train((8, 8), 2.0, n_layers=4, n_era=1, n_epoch=3, batch_size=16, base_lr=1e-3, with_force=True, pre_model=pre,
                       raw_weights=pre.raw.detach().numpy(), seed=12)
    assert len(tr.history["loss"]) == 6 and sum(f > 0 for f in tr.history["force"]) == 3
    torch.manual_seed(3)
    ev = ft.flow_eval(pre, 2.0, (8, 8), ensemble_size=128, batch_size=32, rng=np.random.default_rng(1))
    assert 0 < ev["accept_rate"] <= 1 and ev["Q2"] >= 0 and ev["Q2_err"] >= 0 and len(ev["ensemble"]["x"]) == 128


def test_flow_vjp_and_differentiable_flow():
    """fthmc_flow_vjp against torch.autograd on the oracle (random d/dy and per-chain logJ weights), and the flow as a
    differentiable torch operation: the reference's reverse-KL training loss written with ordinary torch code on
    differentiable_flow / differentiable_u1_action gives, through loss.backward(), the gradients autograd gives on the oracle
    -- in the Conv2d parameters of a reference-style ModuleList and in the latent field."""
    from test_engine_emul import _vjp_reference
    L, layers, B = 16, 8, 3
    flow = O.random_flow(n_layers=layers, seed=41, scale=1.5)
    raw = _raw_of(flow)
    gen = torch.Generator().manual_seed(12)
    x = torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 * np.pi
    gy = torch.randn(B, 2, L, L, generator=gen, dtype=torch.float64)
    glj = torch.randn(B, generator=gen, dtype=torch.float64)
    gx_ref, gw_ref = _vjp_reference(flow, x, gy, glj)
    gw, gx = ft.flow_vjp(ft.PackedFlow(raw), x.cuda(), gy.cuda(), glj.cuda())
    assert relerr(gx.cpu().numpy(), gx_ref.numpy()) < REL
    assert relerr(gw.numpy(), gw_ref.numpy()) < 1e-9
    # reverse-KL loss through autograd on the kernels: mean(logq - logp), logq = const - logJ, logp = -S(x)
    beta = 2.0
    mod = module_like(dict(weights=raw))
    xi = x.cuda().requires_grad_(True)
    y, lj = ft.differentiable_flow(mod, xi)
    loss = torch.mean(-lj + ft.differentiable_u1_action(beta, y))
    loss.backward()
    a_ref, g_ref = O.ft_action_weight_grad(beta, flow, x)
    assert abs(float(loss.detach()) - float(a_ref.mean())) < 1e-10 * abs(float(a_ref.mean()))
    got = np.stack([np.concatenate([np.concatenate([c.weight.grad.numpy().ravel(), c.bias.grad.numpy().ravel()])
                                    for c in layer.plaq_coupling.net if hasattr(c, "weight")]) for layer in mod])
    assert relerr(got, (g_ref / B).numpy()) < 1e-9
    assert relerr(xi.grad.cpu().numpy(), (O.ft_force(beta, flow, x) / B).numpy()) < REL


def test_flow_independence_sampler():
    """apply_flow_to_prior / make_mcmc_ensemble (ipynb/field_transformation.py:37-83) on the forward-flow kernel: logq
    and logp of the proposals against the oracle, and the accept/reject chain against a replay of the reference's loop."""
    from fthmc_b200 import sampler
    flow = O.random_flow(n_layers=8, seed=6, scale=1.5)
    pf = ft.PackedFlow(_raw_of(flow))
    L, beta = 8, 2.0
    gen = torch.Generator(device="cuda"); gen.manual_seed(9)
    xi, x, logq = sampler.apply_flow_to_prior(pf, (L, L), 16, generator=gen)
    y, lj = O.ft_flow_logJ(flow, xi.cpu())
    assert np.max(np.abs(x.cpu().numpy() - y.numpy())) < 1e-11
    assert relerr(logq.cpu().numpy(), (-2 * L * L * np.log(2 * np.pi) - lj).numpy()) < REL
    gen.manual_seed(9)
    torch.manual_seed(4)
    h = sampler.make_mcmc_ensemble(pf, beta, (L, L), 16, 40, generator=gen)
    assert len(h["x"]) == 40 and h["accepted"][0] is True and 0 < sum(h["accepted"]) <= 40
    # replay: same proposals (same device generator), same uniforms (same host generator)
    gen.manual_seed(9)
    torch.manual_seed(4)
    last = None
    for i in range(40):
        if i % 16 == 0:
            _, xb, lq = sampler.apply_flow_to_prior(pf, (L, L), 16, generator=gen)
            lp = -O.u1_action(beta, xb.cpu()); lq = lq.cpu()
        cur = (float(lp[i % 16]), float(lq[i % 16]))
        if last is None:
            acc = True
        else:
            acc = bool(torch.rand(1) < min(1.0, float(np.exp((cur[0] - cur[1]) - (last[0] - last[1])))))
        if acc:
            last = cur
        assert acc == h["accepted"][i] and abs(float(h["logp"][i]) - last[0]) < 1e-9 * abs(last[0])
    assert 0 < float(sampler.compute_ess(torch.stack(h["logp"]), torch.stack(h["logq"]))) <= 1


def test_field_transformation_class(golden):
    """The package's class form (fthmc/ft_hmc.py:108-257) in the package conventions ([-pi,pi) torch_mod): action / force /
    flow_forward / flow_backward against the oracle with convention 1, and a latent-space trajectory against the same
    steps done with the oracle (the reference's buggy leapfrog / calc_energy are deliberately not reproduced)."""
    import types
    g = golden("copyB_L8")
    flow_o = oracle_flow_from_golden(g)
    assert flow_o.convention == 1
    cfg = types.SimpleNamespace(beta=2.0, volume=64, lat=[8, 8], nd=2)
    lf = types.SimpleNamespace(dt=0.05, tau=0.2, nstep=4)
    FT = ft.FieldTransformation(ft.PackedFlow(g["weights"], activation=str(g["activation"]), convention=1), cfg, lf)
    x = T(g["x"])
    assert relerr(FT.action(x).numpy(), O.ft_action(cfg.beta, flow_o, x).numpy()) < REL
    assert relerr(FT.force(x).numpy(), O.ft_force(cfg.beta, flow_o, x).numpy()) < REL
    y, lj = FT.flow_forward(x)
    assert np.max(np.abs(y.numpy() - g["flow_fwd"])) < 1e-12 and relerr(lj.numpy(), g["logJ"]) < 1e-12
    xi, lji = FT.flow_backward(y)
    assert np.max(np.abs(xi.numpy() - g["flow_inv_of_fwd"])) < 1e-11
    torch.manual_seed(21)
    xn, m = FT.hmc(x)
    torch.manual_seed(21)
    xc = x.cuda()
    v = torch.randn_like(xc).cpu()
    h0 = O.ft_action(cfg.beta, flow_o, x) + 0.5 * (v * v).flatten(1).sum(-1)
    x_, v_ = O.ft_leapfrog(cfg.beta, lf.dt, lf.nstep, flow_o, x, v)
    x_ = torch.remainder(x_ + np.pi, 2 * np.pi) - np.pi
    h1 = O.ft_action(cfg.beta, flow_o, x_) + 0.5 * (v_ * v_).flatten(1).sum(-1)
    assert np.max(np.abs(m["dh"].cpu().numpy() - (h1 - h0).numpy())) < 1e-8
    acc = m["acc"].cpu()
    assert np.max(np.abs(xn.cpu().numpy() - torch.where(acc[:, None, None, None], x_, x).numpy())) < 1e-9
    lm = FT.lattice_metrics(xn, torch.zeros(x.shape[0], dtype=torch.float64, device=xn.device))
    assert lm["plaq"].shape == (x.shape[0],) and float(lm["plaq"].abs().max()) <= 1.0


def test_stencils_beyond_grid_y_limit():
    """BASELINE config 5 asks for 65 536 chains: the stencil entry points take batches beyond the 65 535 a CUDA grid.y holds."""
    B, L = 65536 + 300, 8
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi
    P = ft.Param(beta=4.0, lat=(L, L))
    xd = x.cuda()
    sel = torch.tensor([0, 1, 65534, 65535, 65536, B - 1])
    assert relerr(ft.u1_action(4.0, xd).cpu()[sel].numpy(), O.u1_action(4.0, x[sel]).numpy()) < REL
    assert np.array_equal(ft.topocharge(xd).cpu()[sel].numpy(), np.array([float(O.topocharge(x[i])) for i in sel]))
    f = ft.force(P, xd).cpu()
    for i in sel:
        assert relerr(f[i].numpy(), O.force_closed_form(4.0, x[i]).numpy()) < REL


def test_resident_path_selection():
    """Which lattices run with the whole chain in ONE SM's shared memory and which on a thread-block cluster: a guard for the
    L=32 shared-memory budget (232 384 of 232 448 bytes; one more 128-byte step of the engine object would silently move the
    headline configuration onto the 2-CTA cluster path)."""
    L = ft.lib()
    assert [L.fthmc_chain_ranks(n, n, 1) for n in (8, 16, 32, 48, 64, 128)] == [1, 1, 1, 3, 4, 16]
    assert L.fthmc_chain_ranks(32, 32, 0) == 1 and L.fthmc_chain_ranks(256, 256, 1) == 0


def test_copyB_physics_golden(golden):
    """Reference-generated goldens of the package copy's physics helpers (fthmc/utils/qed_helpers.py:73-116 batch_charges /
    topo_charge, :166-186 BatchAction, :191-242 ft_flow / ft_flow_inv / ft_action / ft_force on (B,2,L,L), :261-311 action /
    force / leapfrog / hmc) against the CUDA entry points and the FieldTransformation class, [-pi,pi) convention."""
    import types
    g = golden("copyB_physics_L8")
    pf = packed(g)
    assert pf.convention == 1
    beta, x = float(g["beta"]), T(g["x"])
    assert relerr(ft.u1_action(beta, x).numpy(), g["batch_action"]) < REL
    assert np.max(np.abs(ft.topo_charge(x).numpy() - g["batch_charges"])) < 1e-11
    cfg = types.SimpleNamespace(beta=beta, volume=64, lat=[8, 8], nd=2)
    lf = types.SimpleNamespace(dt=float(g["dt"]), tau=float(g["dt"]) * int(g["nstep"]), nstep=int(g["nstep"]))
    FT = ft.FieldTransformation(pf, cfg, lf)
    assert relerr(FT.action(x).numpy(), g["ft_action"]) < REL
    assert relerr(FT.force(x).numpy(), g["ft_force"]) < REL
    y, _ = FT.flow_forward(x)
    assert np.max(np.abs(y.numpy() - g["ft_flow"])) < 1e-12
    for b in range(x.shape[0]):
        xi, _ = FT.flow_backward(T(g["ft_flow"][b:b + 1]))
        assert np.max(np.abs(xi.numpy()[0] - g["ft_flow_inv_of_fwd"][b])) < 1e-10
    lm = FT.lattice_metrics(x, torch.zeros(x.shape[0]))
    assert relerr(lm["plaq"].numpy(), -g["batch_action"] / (beta * 64)) < REL
    assert np.max(np.abs(lm["q"].numpy() - g["batch_charges"])) < 1e-11
    # plain single-chain helpers
    P = ft.Param(beta=beta, lat=(8, 8), tau=lf.tau, nstep=lf.nstep)
    assert abs(float(ft.action(P, x[0])) - float(g["action"])) < 1e-10 * abs(float(g["action"]))
    assert relerr(ft.force(P, x[0], order=0).numpy(), g["force"]) < REL
    lx, lp = ft.leapfrog(P, x[0], T(g["lf_p"]))
    assert np.max(np.abs(lx.numpy() - g["lf_x_out"])) < 1e-11 and np.max(np.abs(lp.numpy() - g["lf_p_out"])) < 1e-11
    r = ft.hmc_batch(P, T(g["traj_x"]), T(g["traj_p"]), T(g["traj_u"]))
    assert np.max(np.abs(r["dH"].numpy() - g["traj_dH"])) < 1e-8
    assert np.array_equal(r["acc"].numpy(), g["traj_acc"])
    assert np.max(np.abs(r["field"].numpy() - g["traj_out"])) < 1e-9


def test_run_history_mirrors(golden):
    """run_hmc (fthmc/hmc.py:57-175) and FieldTransformation.run (fthmc/ft_hmc.py:272-346) without their plots and dumps:
    the history dicts carry the reference's keys, one entry per trajectory, and run_hmc seeded like the reference's loop
    follows the reference's free-running chain (the run_L8 golden)."""
    import io, types
    g = golden("run_L8")
    n = int(g["ntraj"])
    P = ft.Param(beta=float(g["plain_beta"]), lat=(8, 8), tau=float(g["plain_tau"]), nstep=int(g["plain_nstep"]), ntraj=n, nrun=1)
    torch.manual_seed(int(g["seed"]))
    out = io.StringIO()
    fields, hist = ft.run_hmc(P, T(g["x0"]), out=out)
    h = hist[0]
    assert set(h) == {"traj", "dt", "acc", "dH", "plaq", "q", "dq"} and len(h["dH"]) == n and out.getvalue().count("Traj:") == n
    assert np.max(np.abs(np.array(h["dH"]) - g["plain_dH"])) < 1e-8
    assert [bool(a) for a in h["acc"]] == [bool(a) for a in g["plain_acc"]]
    assert np.max(np.abs(fields[0][0].numpy() - g["plain_final"])) < 1e-8
    gb = golden("copyB_L8")
    cfg = types.SimpleNamespace(beta=2.0, volume=64, lat=[8, 8], nd=2)
    lf = types.SimpleNamespace(dt=0.05, tau=0.2, nstep=4)
    FT = ft.FieldTransformation(ft.PackedFlow(gb["weights"], activation=str(gb["activation"]), convention=1), cfg, lf)
    hist = FT.run(T(gb["x"]), num_trajs=5, nprint=2, out=out)
    assert {"traj", "dt", "acc", "dh", "exp_mdh", "plaq", "q", "dq"} <= set(hist) and all(len(v) == 5 for v in hist.values())
    assert float(hist["plaq"][-1].abs().max()) <= 1.0


def test_two_streams_do_not_share_scratch(golden):
    """The kernels keep per-CTA state (momenta, layer blocks of the adjoint) in the workspace the Python mirror hands them:
    calls in flight on two CUDA streams must get different workspaces.  Two different batches run concurrently on two
    streams, repeatedly; each must reproduce its single-stream result bit for bit."""
    g = golden("ft_L16_b6")
    pf = packed(g)
    P = ft.Param(beta=6.0, lat=(16, 16), tau=1.0, nstep=10)
    gen = torch.Generator().manual_seed(77)
    xa = ((torch.rand(300, 2, 16, 16, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    xb = ((torch.rand(300, 2, 16, 16, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
    ra = ft.ft_hmc_batch(P, pf, xa, seed=5)
    rb = ft.ft_hmc_batch(P, pf, xb, seed=6)
    fa, fb = ft.ft_force(P, pf, xa), ft.ft_force(P, pf, xb)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(3):
        with torch.cuda.stream(s1):
            ra2 = ft.ft_hmc_batch(P, pf, xa, seed=5)
            fa2 = ft.ft_force(P, pf, xa)
        with torch.cuda.stream(s2):
            rb2 = ft.ft_hmc_batch(P, pf, xb, seed=6)
            fb2 = ft.ft_force(P, pf, xb)
        torch.cuda.synchronize()
        assert torch.equal(ra2["field"], ra["field"]) and torch.equal(rb2["field"], rb["field"])
        assert torch.equal(ra2["dH"], ra["dH"]) and torch.equal(rb2["dH"], rb["dH"])
        assert torch.equal(fa2, fa) and torch.equal(fb2, fb)


def test_pack_sees_inplace_weight_edits(golden):
    """pack(flow) validates its cached device copy with a digest of the weights: an in-place `.data` edit (which does not
    bump the tensor version; the reference's set_weights uses one) must be seen by the next call."""
    g = golden("ft_L8_n8")
    flow = module_like(g)
    x = T(g["x"])
    P = ft.Param(beta=float(g["beta"]), lat=(8, 8))
    a0 = ft.ft_action(P, flow, x)
    assert relerr(a0.numpy(), g["ft_action"]) < REL
    assert ft.pack(flow) is ft.pack(flow)
    flow[0].plaq_coupling.net[0].bias.data.fill_(0.25)
    a1 = ft.ft_action(P, flow, x)
    assert float((a1 - a0).abs().max()) > 1e-6


def test_statistical_known_answers():
    """The reference's recorded physics: <cos P> = I1(beta)/I0(beta) (PLAQ_EXACT, fthmc/config.py:37-47:
    0.69777 at beta=2) and <Q^2> = 1.23 +- 0.02 at L=8, beta=2 (hmc_2dU1.py:661), from 2048 device-RNG chains of plain HMC
    and of FT-HMC through the random-init flow.  Tolerances are > 5 sigma of the chain-to-chain spread."""
    L, beta, B, ntraj = 8, 2.0, 2048, 80
    P = ft.Param(beta=beta, lat=(L, L), tau=1.0, nstep=10)
    x0 = torch.zeros(B, 2, L, L, dtype=torch.float64).cuda()
    pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
    for r in (ft.hmc_run_batch(P, x0, ntraj, seed=11), ft.ft_hmc_run_batch(P, pf, x0, ntraj, seed=12)):
        tail = slice(ntraj // 2, None)
        assert float(r["acc"][tail].double().mean()) > 0.8
        assert abs(float(r["plaq"][tail].mean()) - 0.69777) < 2.5e-3
        q = r["topo"][tail]
        assert torch.equal(q, torch.round(q))
        assert abs(float((q * q).mean()) - 1.23) < 0.08
        assert abs(float(torch.exp(-r["dH"][tail]).mean()) - 1.0) < 0.02          # <exp(-dH)> = 1


def test_host_batches_pipelined_equal_single_launch():
    """Host batches of >= 8 device waves run as chunks of whole waves with their copies on side streams: same fields,
    decisions and charges as the single launch on device tensors (sums to rounding), in teacher-forced (p, u given) and
    device-RNG mode."""
    flow = O.random_flow(n_layers=4, seed=5, scale=2.0)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)])
                    for lw in flow.layers])
    pf = ft.PackedFlow(raw)
    B = 8 * torch.cuda.get_device_properties(0).multi_processor_count + 37
    gen = torch.Generator().manual_seed(99)
    x = (torch.rand(B, 2, 8, 8, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi
    p = torch.randn(B, 2, 8, 8, generator=gen, dtype=torch.float64)
    u = torch.rand(B, generator=gen, dtype=torch.float64)
    P = ft.Param(beta=2.0, lat=(8, 8), tau=0.4, nstep=4)
    for kw_host, kw_dev in ((dict(p=p, u=u), dict(p=p.cuda(), u=u.cuda())), (dict(seed=7, traj=3, chain0=11),) * 2):
        rh = ft.ft_hmc_batch(P, pf, x, want_h=True, **kw_host)
        rd = ft.ft_hmc_batch(P, pf, x.cuda(), want_h=True, **kw_dev)
        assert not rh["field"].is_cuda and rh["field"].is_pinned()
        for k in ("field", "acc", "topo"):
            assert torch.equal(rh[k], rd[k].cpu()), k
        for k in ("dH", "exp_mdH", "plaq", "h0", "h1"):          # per-chain sums: the CTA width (hence the order of a chain's
            assert float(torch.max(torch.abs(rh[k] - rd[k].cpu()))) < 1e-11, k      # reduction) may depend on the launch's batch


def test_bitwise_reproducibility():
    """Same inputs, same bits: the resident-chain kernels have fixed reduction orders (no atomics), on the single-CTA
    path, the cluster path (DSMEM halos, cluster-wide reductions) and in the weight-gradient mode.  A missing barrier
    shows up here as run-to-run differences."""
    pf = ft.PackedFlow(ft.default_init_raw(24, 3647))
    for L, B in ((32, 300), (64, 40), (128, 9)):
        gen = torch.Generator().manual_seed(L)
        x = ((torch.rand(B, 2, L, L, generator=gen, dtype=torch.float64) * 2 - 1) * np.pi).cuda()
        P = ft.Param(beta=4.0, lat=(L, L), tau=0.5, nstep=3)
        ref = ft.ft_hmc_batch(P, pf, x, seed=5, traj=2)
        for _ in range(3):
            r = ft.ft_hmc_batch(P, pf, x, seed=5, traj=2)
            assert torch.equal(r["field"], ref["field"]) and torch.equal(r["dH"], ref["dH"]) and torch.equal(r["topo"], ref["topo"])
    x = torch.rand(200, 2, 32, 32, dtype=torch.float64, device="cuda") * 2 * np.pi
    P = ft.Param(beta=4.0, lat=(32, 32))
    a0, g0 = ft.ft_action_grad(P, pf, x)
    for _ in range(3):
        a1, g1 = ft.ft_action_grad(P, pf, x)
        assert torch.equal(a0, a1) and torch.equal(g0, g1)


def test_float32_callers_get_float32_back():
    """The package copy runs in fp32 by default (fthmc/config.py:25-31): fp32 tensors are accepted by every flow entry
    point, computed in fp64 on the device and returned as fp32 (agreement with the fp64 result at fp32 resolution)."""
    flow = O.random_flow(n_layers=4, seed=9)
    pf = ft.PackedFlow(_raw_of(flow))
    x64 = (torch.rand(2, 2, 8, 8, dtype=torch.float64) * 2 - 1) * np.pi
    x32 = x64.float()
    P = ft.Param(beta=2.0, lat=(8, 8), tau=0.2, nstep=2)
    y32 = ft.ft_flow(pf, x32)
    assert y32.dtype == torch.float32
    assert float((y32.double() - ft.ft_flow(pf, x32.double())).abs().max()) < 1e-6
    assert ft.ft_force(P, pf, x32).dtype == torch.float32 and ft.ft_action(P, pf, x32).dtype == torch.float32
    r = ft.ft_hmc_batch(P, pf, x32.cuda(), seed=1)
    assert r["field"].dtype == torch.float32 and r["field"].is_cuda and r["acc"].dtype == torch.bool
    assert ft.action(P, x32[0]).dtype == torch.float32        # the stencils have native fp32 kernels


def test_errors_are_loud():
    P = ft.Param(beta=1.0, lat=(6, 6))
    with pytest.raises(ft.FthmcError) as e:
        ft.hmc_batch(P, torch.zeros(1, 2, 6, 6))
    assert e.value.code == -2
    flow = O.random_flow(n_layers=2, seed=1)
    raw = np.stack([np.concatenate([np.concatenate([w.numpy().ravel(), b.numpy().ravel()]) for w, b in zip(lw.w, lw.b)])
                    for lw in flow.layers])
    pf = ft.PackedFlow(raw)
    with pytest.raises(ft.FthmcError) as e:
        ft.ft_flow(pf, torch.zeros(1, 2, 256, 256))     # beyond a 16-CTA cluster's shared memory
    assert e.value.code == -2
    with pytest.raises(ft.FthmcError):
        ft.PackedFlow(np.zeros((2, 900)))
