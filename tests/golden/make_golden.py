#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REAL reference (nftqcd/fthmc at /root/reference).

Runs only in the build container (the reference cannot travel to the GPU box).  The reference is
imported unmodified:
  * hmc_2dU1.py                       (plain HMC, LeakyReLU flow copy)
  * ipynb/field_transformation.py     (flow library, copy A)
  * ipynb/ft_hmc.py lines 1-514       (everything before its module-level experiment block)
  * fthmc/utils/layers.py             (copy B: [-pi,pi) convention)
Momenta and Metropolis uniforms are the reference's own torch-RNG draws: they are recovered by
replaying the generator from the same seed in the same order (randn_like, then rand([],float64)).

usage:  python tests/golden/make_golden.py [--ref /root/reference]
"""
import argparse
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference(ref):
    sys.path.insert(0, ref)
    sys.path.insert(0, os.path.join(ref, "ipynb"))
    import hmc_2dU1 as plain                      # noqa
    import field_transformation as ftlib          # noqa
    src = open(os.path.join(ref, "ipynb", "ft_hmc.py")).read().split("\n")
    cut = next(i for i, l in enumerate(src) if l.startswith("# set param"))
    fthmc_mod = types.ModuleType("ref_ft_hmc")
    fthmc_mod.__file__ = os.path.join(ref, "ipynb", "ft_hmc.py")
    exec(compile("\n".join(src[:cut]), fthmc_mod.__file__, "exec"), fthmc_mod.__dict__)
    torch.set_default_dtype(torch.float64)
    return plain, ftlib, fthmc_mod


def flat_weights(flow):
    rows = []
    for layer in flow:
        convs = [m for m in layer.plaq_coupling.net if hasattr(m, "weight")]
        rows.append(np.concatenate([np.concatenate([c.weight.detach().numpy().ravel(),
                                                    c.bias.detach().numpy().ravel()]) for c in convs]))
    return np.stack(rows)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def replay_p_u(seed, shape):
    st = torch.get_rng_state()
    torch.manual_seed(seed)
    p = torch.randn(shape, dtype=torch.float64)
    u = torch.rand([], dtype=torch.float64)
    torch.set_rng_state(st)
    return p, u


def gen_plain(plain, out):
    """BASELINE config 1: plain HMC L=8 beta=2 tau=1 nstep=10 (hmc_2dU1.py on CPU, fp64)."""
    P = plain.Param(beta=2.0, lat=(8, 8), tau=1.0, nstep=10)
    torch.manual_seed(1331)
    x = torch.empty((2, 8, 8)).uniform_(-np.pi, np.pi)
    rec = dict(beta=2.0, dt=P.dt, nstep=10, x0=x.numpy().copy())
    rec["action"] = float(plain.action(P, x))
    rec["force"] = plain.force(P, x.clone()).numpy().copy()
    rec["topo"] = float(plain.topocharge(x))
    big = x * 7.3 + 1.1
    rec["reg_in"] = big.numpy().copy()
    rec["reg_out"] = plain.regularize(big).numpy().copy()
    p = torch.randn_like(x)
    lx, lp = plain.leapfrog(P, x.clone(), p)
    rec["lf_p"], rec["lf_x_out"], rec["lf_p_out"] = p.numpy().copy(), lx.numpy().copy(), lp.numpy().copy()
    # free-running chain of 40 trajectories; every trajectory's input, p, u and outputs are stored
    xs, ps, us, dHs, accs, outs, topos, plaqs = [], [], [], [], [], [], [], []
    cur = x.clone()
    for n in range(40):
        seed = 5000 + n
        p, u = replay_p_u(seed, cur.shape)
        torch.manual_seed(seed)
        dH, e, acc, new = plain.hmc(P, cur.clone())
        xs.append(cur.numpy().copy()); ps.append(p.numpy().copy()); us.append(float(u))
        dHs.append(float(dH)); accs.append(bool(acc)); outs.append(new.numpy().copy())
        topos.append(float(plain.topocharge(new)))
        plaqs.append(float(plain.action(P, new) / (-P.beta * P.volume)))
        cur = new.detach().clone()
    rec.update(traj_x=np.stack(xs), traj_p=np.stack(ps), traj_u=np.array(us), traj_dH=np.array(dHs),
               traj_acc=np.array(accs), traj_out=np.stack(outs), traj_topo=np.array(topos),
               traj_plaq=np.array(plaqs))
    # cold start known answers (hmc_2dU1.py:692-693)
    z = torch.zeros((2, 8, 8))
    rec["cold_plaq"] = float(plain.action(P, z) / (-P.beta * P.volume))
    rec["cold_topo"] = float(plain.topocharge(z))
    np.savez_compressed(os.path.join(out, "plain_L8.npz"), **rec)
    print("plain_L8: acc rate", np.mean(accs), "dH", dHs[:4])


def gen_flow_case(ftlib, ref, name, out, *, L, beta, n_layers, B, nstep, ntraj, scale=1.0, seed_w=3647,
                  tau=1.0, hot=True):
    torch.manual_seed(seed_w)
    flow = ftlib.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=n_layers, n_mixture_comps=2,
                                      hidden_sizes=[8, 8], kernel_size=3)
    flow.eval()
    if scale != 1.0:
        with torch.no_grad():
            for prm in flow.parameters():
                prm.mul_(scale)
    for prm in flow.parameters():
        prm.requires_grad_(False)
    P = ref.Param(beta=beta, lat=(L, L), tau=tau, nstep=nstep)
    torch.manual_seed(1331)
    if hot:
        x = torch.empty((B, 2, L, L)).uniform_(-np.pi, np.pi)
    else:
        x = 0.3 * torch.randn((B, 2, L, L))
    rec = dict(L=L, beta=beta, dt=P.dt, nstep=nstep, n_layers=n_layers, weights=flat_weights(flow),
               activation="silu", convention=0, x=x.numpy().copy())
    # per-layer forward of layer 0..: outputs and logJ
    cur = x.clone()
    lay_out, lay_logJ = [], []
    for layer in flow:
        cur, lj = layer.forward(cur)
        lay_out.append(cur.numpy().copy()); lay_logJ.append(lj.numpy().copy())
    rec["layer_out"] = np.stack(lay_out[:4])          # first 4 layers in full
    rec["layer_logJ"] = np.stack(lay_logJ)            # all layers
    y = ref.ft_flow(flow, x.clone())
    rec["flow_fwd"] = y.numpy().copy()
    rec["ft_action"] = ref.ft_action(P, flow, x.clone()).detach().numpy().copy()
    rec["ft_force"] = ref.ft_force(P, flow, x.clone()).numpy().copy()
    # reverse of single chains (the bisection stop test is tensor-global, so the reference path is B=1)
    inv, inv_logJ0 = [], []
    for b in range(B):
        xi = quiet(ref.ft_flow_inv, flow, y[b:b + 1].clone())
        inv.append(xi.numpy()[0].copy())
        _, lj = flow[-1].reverse(y[b:b + 1].clone())
        inv_logJ0.append(float(lj))
    rec["flow_inv_of_fwd"] = np.stack(inv)
    rec["last_layer_reverse_logJ"] = np.array(inv_logJ0)
    # teacher-forced + free-running ft_hmc on chain 0
    cur = wrap(y[0:1].clone())
    xs, ps, us, dHs, accs, outs, topos, plaqs = [], [], [], [], [], [], [], []
    for n in range(ntraj):
        seed = 7000 + n
        p, u = replay_p_u(seed, cur.shape)
        torch.manual_seed(seed)
        dH, e, acc, new = quiet(ref.ft_hmc, P, flow, cur.clone())
        xs.append(cur.numpy()[0].copy()); ps.append(p.numpy()[0].copy()); us.append(float(u))
        dHs.append(dH); accs.append(bool(acc)); outs.append(new.numpy()[0].copy())
        topos.append(float(ref.topocharge(new[0])))
        plaqs.append(float(ref.action(P, new[0]) / (-P.beta * P.volume)))
        cur = new.detach().clone()
    rec.update(traj_x=np.stack(xs), traj_p=np.stack(ps), traj_u=np.array(us), traj_dH=np.array(dHs),
               traj_acc=np.array(accs), traj_out=np.stack(outs), traj_topo=np.array(topos),
               traj_plaq=np.array(plaqs))
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print(name, "acc", accs, "dH", [f"{d:.4g}" for d in dHs])


def wrap(x):
    return torch.remainder(x + np.pi, 2 * np.pi) - np.pi


def gen_leaky(plain, out):
    """hmc_2dU1.py's embedded flow copy (LeakyReLU, :260) and its round-trip self check (:719-745)."""
    L = 8
    torch.manual_seed(99)
    flow = plain.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=16, n_mixture_comps=2,
                                      hidden_sizes=[8, 8], kernel_size=3)
    flow.eval()
    with torch.no_grad():
        for prm in flow.parameters():
            prm.mul_(2.0)
    torch.manual_seed(1331)
    x = torch.empty((2, 2, L, L)).uniform_(-np.pi, np.pi)
    cur, lj_tot = x.clone(), 0.0
    for layer in flow:
        cur, lj = layer.forward(cur)
        lj_tot = lj_tot + lj
    rec = dict(L=L, n_layers=16, weights=flat_weights(flow), activation="leaky_relu", convention=0,
               x=x.numpy().copy(), flow_fwd=cur.detach().numpy().copy(), logJ=lj_tot.detach().numpy().copy())
    invs = []
    for b in range(2):
        c = cur[b:b + 1].detach().clone()
        for layer in reversed(flow):
            c, _ = layer.reverse(c)
        invs.append(c.detach().numpy()[0].copy())
    rec["flow_inv_of_fwd"] = np.stack(invs)
    np.savez_compressed(os.path.join(out, "leaky_L8.npz"), **rec)
    print("leaky_L8 done; roundtrip err", np.abs(wrap(torch.from_numpy(np.stack(invs)) - x)).max().item())


def gen_copyB(refroot, out):
    """Copy B (fthmc/utils/layers.py): torch_mod in [-pi,pi), bisection on [-pi,pi]."""
    st = io.StringIO()
    with contextlib.redirect_stdout(st), contextlib.redirect_stderr(st):
        import fthmc.utils.layers as LB
    torch.set_default_dtype(torch.float64)
    torch.set_default_tensor_type(torch.DoubleTensor)
    L = 8
    torch.manual_seed(4242)
    try:
        flow = LB.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=8, n_mixture_comps=2,
                                       hidden_sizes=[8, 8], kernel_size=3)
    except TypeError:
        flow = LB.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=8, n_mixture_comps=2,
                                       hidden_sizes=[8, 8], kernel_size=3, activation_fn="silu")
    flow.eval()
    flow = flow.double()
    for layer in flow:
        layer.active_mask = layer.active_mask.double()
    with torch.no_grad():
        for prm in flow.parameters():
            prm.mul_(2.0)
    torch.manual_seed(1331)
    x = torch.empty((2, 2, L, L)).uniform_(-np.pi, np.pi)
    cur, lj_tot = x.clone(), 0.0
    for layer in flow:
        cur, lj = layer.forward(cur)
        lj_tot = lj_tot + lj
    rec = dict(L=L, n_layers=8, weights=flat_weights(flow), activation="silu", convention=1,
               x=x.numpy().copy(), flow_fwd=cur.detach().numpy().copy(), logJ=lj_tot.detach().numpy().copy())
    invs = []
    for b in range(2):
        c = cur[b:b + 1].detach().clone()
        for layer in reversed(flow):
            c, _ = layer.reverse(c)
        invs.append(c.detach().numpy()[0].copy())
    rec["flow_inv_of_fwd"] = np.stack(invs)
    np.savez_compressed(os.path.join(out, "copyB_L8.npz"), **rec)
    print("copyB_L8 done; roundtrip err", np.abs(wrap(torch.from_numpy(np.stack(invs)) - x)).max().item())


def gen_copyB_physics(refroot, out):
    """Copy B's physics helpers, fthmc/utils/qed_helpers.py: BatchAction (:166-186), batch_charges / topo_charge (:73-77,
    :108-116), ft_flow / ft_flow_inv / ft_action / ft_force on (B,2,L,L) (:191-242) through a copy-B flow
    (fthmc/utils/layers.py: [-pi,pi) convention), and the plain action / force / leapfrog / hmc (:261-311)."""
    st = io.StringIO()
    with contextlib.redirect_stdout(st), contextlib.redirect_stderr(st):
        import fthmc.utils.layers as LB
        import fthmc.utils.qed_helpers as Q
    torch.set_default_dtype(torch.float64)
    torch.set_default_tensor_type(torch.DoubleTensor)
    L, B = 8, 3
    torch.manual_seed(4243)
    flow = LB.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=8, n_mixture_comps=2, hidden_sizes=[8, 8], kernel_size=3)
    flow.eval()
    flow = flow.double()
    for layer in flow:
        layer.active_mask = layer.active_mask.double()
    with torch.no_grad():
        for prm in flow.parameters():
            prm.mul_(2.0)
    for prm in flow.parameters():
        prm.requires_grad_(False)
    P = types.SimpleNamespace(beta=2.5, dt=0.1, nstep=5, volume=L * L)
    torch.manual_seed(1332)
    x = torch.empty((B, 2, L, L)).uniform_(-np.pi, np.pi)
    rec = dict(L=L, n_layers=8, weights=flat_weights(flow), activation="silu", convention=1, beta=P.beta, dt=P.dt, nstep=P.nstep,
               x=x.numpy().copy())
    rec["batch_action"] = Q.BatchAction(P.beta)(x).numpy().copy()
    rec["batch_charges"] = Q.batch_charges(x).numpy().copy()
    rec["topo_charge"] = Q.topo_charge(x).numpy().copy()
    rec["ft_flow"] = Q.ft_flow(flow, x.clone()).numpy().copy()
    rec["ft_action"] = Q.ft_action(P, flow, x.clone()).detach().numpy().copy()
    rec["ft_force"] = Q.ft_force(P, flow, x.clone()).numpy().copy()
    inv = [Q.ft_flow_inv(flow, torch.from_numpy(rec["ft_flow"][b:b + 1]).clone()).numpy()[0].copy() for b in range(B)]
    rec["ft_flow_inv_of_fwd"] = np.stack(inv)
    # plain single-chain functions
    x1 = x[0].clone()
    rec["action"] = float(Q.action(P, x1))
    rec["force"] = Q.force(P, x1.clone()).numpy().copy()
    p1 = torch.randn_like(x1)
    lx, lp = Q.leapfrog(P, x1.clone(), p1, verbose=False)
    rec.update(lf_p=p1.numpy().copy(), lf_x_out=lx.numpy().copy(), lf_p_out=lp.numpy().copy())
    xs, ps, us, dHs, accs, outs = [], [], [], [], [], []
    cur = x1.clone()
    for n in range(8):
        seed = 9100 + n
        p, u = replay_p_u(seed, cur.shape)
        torch.manual_seed(seed)
        dH, e, acc, new = Q.hmc(P, cur.clone(), verbose=False)
        xs.append(cur.numpy().copy()); ps.append(p.numpy().copy()); us.append(float(u))
        dHs.append(float(dH)); accs.append(bool(acc)); outs.append(new.numpy().copy())
        cur = new.detach().clone()
    rec.update(traj_x=np.stack(xs), traj_p=np.stack(ps), traj_u=np.array(us), traj_dH=np.array(dHs), traj_acc=np.array(accs),
               traj_out=np.stack(outs))
    np.savez_compressed(os.path.join(out, "copyB_physics_L8.npz"), **rec)
    print("copyB_physics_L8: ft_action", rec["ft_action"], "hmc acc", accs, "dH", dHs[:3])


def thousand_inputs(L, n=1000, seed=20261018):
    """Deterministic inputs for the 1000-trajectory parity runs; regenerated (not stored) by the tests.
    torch's CPU generator is bit-reproducible across machines."""
    g = torch.Generator().manual_seed(seed)
    amp = 0.2 + 1.3 * torch.rand(n, 1, 1, 1, generator=g, dtype=torch.float64)
    x = amp * torch.randn(n, 2, L, L, generator=g, dtype=torch.float64)
    p = torch.randn(n, 2, L, L, generator=g, dtype=torch.float64)
    u = torch.rand(n, generator=g, dtype=torch.float64)
    return x, p, u


def patched_rng(p_next, u_next):
    """Feed the reference's own randn_like / rand calls with prescribed draws."""
    class Ctx:
        def __enter__(self):
            self.rl, self.r = torch.randn_like, torch.rand
            torch.randn_like = lambda t, *a, **k: p_next.reshape(t.shape).clone()
            torch.rand = lambda *a, **k: u_next.clone()
        def __exit__(self, *e):
            torch.randn_like, torch.rand = self.rl, self.r
    return Ctx()


def gen_thousand(plain, ftlib, ref, out):
    """north_star: accept/reject and integer topological charge bit-exact over 1000 trajectories,
    momenta and uniforms supplied from the reference run."""
    L, n = 8, 1000
    x, p, u = thousand_inputs(L, n)
    # plain HMC, BASELINE config 1 parameters
    P = plain.Param(beta=2.0, lat=(L, L), tau=1.0, nstep=10)
    dH, acc, topo, chk = [], [], [], []
    for i in range(n):
        with patched_rng(p[i], u[i]):
            d, e, a, new = plain.hmc(P, x[i].clone())
        dH.append(float(d)); acc.append(bool(a)); topo.append(float(plain.topocharge(new))); chk.append(float(new.sum()))
    np.savez_compressed(os.path.join(out, "plain_L8_1000.npz"), beta=2.0, dt=P.dt, nstep=10, L=L, dH=np.array(dH),
                        acc=np.array(acc), topo=np.array(topo), field_sum=np.array(chk))
    print("plain 1000: acc rate", np.mean(acc))
    # FT-HMC, 8-layer flow (all mask arrangements), weights x2
    torch.manual_seed(3647)
    flow = ftlib.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=8, n_mixture_comps=2, hidden_sizes=[8, 8], kernel_size=3)
    flow.eval()
    with torch.no_grad():
        for prm in flow.parameters():
            prm.mul_(2.0)
    for prm in flow.parameters():
        prm.requires_grad_(False)
    P = ref.Param(beta=2.0, lat=(L, L), tau=0.6, nstep=6)
    dH, acc, topo, chk = [], [], [], []
    for i in range(n):
        with patched_rng(p[i][None], u[i]):
            d, e, a, new = quiet(ref.ft_hmc, P, flow, x[i][None].clone())
        dH.append(d); acc.append(bool(a)); topo.append(float(ref.topocharge(new[0]))); chk.append(float(new.sum()))
    np.savez_compressed(os.path.join(out, "ft_L8_1000.npz"), beta=2.0, dt=P.dt, nstep=6, L=L, n_layers=8,
                        weights=flat_weights(flow), activation="silu", convention=0, dH=np.array(dH), acc=np.array(acc),
                        topo=np.array(topo), field_sum=np.array(chk))
    print("ft 1000: acc rate", np.mean(acc), "dH range", np.min(dH), np.max(dH))


def headline_flow(ftlib, L, n_layers=24, seed_w=3647):
    """The bench's flow: ipynb/ft_hmc.py:519 seeds 3647, :310-315 builds 24 layers, default Conv2d init."""
    torch.manual_seed(seed_w)
    flow = ftlib.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=n_layers, n_mixture_comps=2, hidden_sizes=[8, 8],
                                      kernel_size=3)
    flow.eval()
    for prm in flow.parameters():
        prm.requires_grad_(False)
    return flow


def gen_many(ftlib, ref, name, out, *, L, beta, legs, seed):
    """BASELINE configs 2 and 3 at depth: teacher-forced ft_hmc (ipynb/ft_hmc.py:420-435) with the 24-layer seed-3647 flow,
    every trajectory from its own input field / momentum / uniform (thousand_inputs(L, n, seed): regenerated by the tests,
    not stored).  legs = [(nstep, n), ...]: consecutive slices of the inputs run at different nstep (tau = 1) so that the
    set holds accepted AND rejected trajectories as well as the headline nstep=10."""
    flow = headline_flow(ftlib, L)
    ntot = sum(n for _, n in legs)
    x, p, u = thousand_inputs(L, ntot, seed=seed)
    dH, acc, topo, chk, nst = [], [], [], [], []
    i = 0
    for nstep, n in legs:
        P = ref.Param(beta=beta, lat=(L, L), tau=1.0, nstep=nstep)
        for _ in range(n):
            with patched_rng(p[i][None], u[i]):
                d, e, a, new = quiet(ref.ft_hmc, P, flow, x[i][None].clone())
            dH.append(d); acc.append(bool(a)); topo.append(float(ref.topocharge(new[0]))); chk.append(float(new.sum()))
            nst.append(nstep)
            i += 1
            if i % 20 == 0:
                print(name, i, "/", ntot, "acc so far", np.mean(acc), flush=True)
    np.savez_compressed(os.path.join(out, name + ".npz"), L=L, beta=beta, tau=1.0, n_layers=24, seed=seed,
                        weights=flat_weights(flow), activation="silu", convention=0, nstep=np.array(nst),
                        dH=np.array(dH), acc=np.array(acc), topo=np.array(topo), field_sum=np.array(chk))
    print(name, "acc rate", np.mean(acc), "dH range", np.min(dH), np.max(dH))


def gen_c4(ftlib, ref, out):
    """BASELINE config 4: L=128 beta=6 with the L=16 flow transferred by the reference's own flow_resize
    (ipynb/ft_hmc.py:489-513): one ft_action / ft_force and one teacher-forced trajectory."""
    L = 128
    flow16 = headline_flow(ftlib, 16)
    flow = ref.flow_resize(flow16, (L, L))
    flow.eval()
    x, p, u = thousand_inputs(L, 1, seed=1284)
    P = ref.Param(beta=6.0, lat=(L, L), tau=1.0, nstep=40)
    rec = dict(L=L, beta=6.0, dt=P.dt, nstep=40, n_layers=24, weights=flat_weights(flow), activation="silu", convention=0,
               seed=1284)
    rec["ft_action"] = ref.ft_action(P, flow, x.clone()).detach().numpy().copy()
    f = ref.ft_force(P, flow, x.clone()).numpy()
    rec["ft_force_sum"] = np.array([f.sum(), np.abs(f).sum(), (f * f).sum()])
    rec["ft_force_row0"] = f[0, :, 0, :].copy()
    rec["ft_force_col5"] = f[0, :, :, 5].copy()
    y = ref.ft_flow(flow, x.clone())
    rec["flow_fwd_sum"] = np.array([float(y.sum()), float((y * y).sum())])
    with patched_rng(p[0][None], u[0]):
        d, e, a, new = quiet(ref.ft_hmc, P, flow, wrap(y).clone())
    rec.update(traj_dH=d, traj_acc=bool(a), traj_topo=float(ref.topocharge(new[0])), traj_field_sum=float(new.sum()),
               traj_plaq=float(ref.action(P, new[0]) / (-P.beta * P.volume)))
    np.savez_compressed(os.path.join(out, "ft_L128_b6.npz"), **rec)
    print("ft_L128_b6: dH", d, "acc", bool(a), "Q", rec["traj_topo"])


def gen_run(plain, ftlib, ref, out):
    """The trajectory loops of run / ft_run (ipynb/ft_hmc.py:199-208, 454-467) seeded ONCE: a free-running chain whose
    momenta and uniforms come from consecutive draws of the torch generator, plus the block statistics of a synthetic
    charge history (ipynb/ft_hmc.py:14-56).  The loops are replayed here without the reference's result-file handling."""
    L, ntraj, seed = 8, 12, 781          # seed picked so that both chains contain rejected trajectories
    rec = dict(L=L, ntraj=ntraj, seed=seed)
    # plain
    P = ref.Param(beta=2.0, lat=(L, L), tau=1.6, nstep=6)
    torch.manual_seed(1331)
    x0 = torch.empty((2, L, L)).uniform_(-np.pi, np.pi)
    rec.update(x0=x0.numpy().copy(), plain_beta=2.0, plain_tau=1.6, plain_nstep=6)
    torch.manual_seed(seed)
    field = x0.clone()
    dHs, accs, plaqs, topos = [], [], [], []
    for i in range(ntraj):
        dH, e, acc, field = ref.hmc(P, field)
        dHs.append(float(dH)); accs.append(bool(acc))
        plaqs.append(float(ref.action(P, field) / (-P.beta * P.volume))); topos.append(float(ref.topocharge(field)))
    rec.update(plain_dH=np.array(dHs), plain_acc=np.array(accs), plain_plaq=np.array(plaqs), plain_topo=np.array(topos),
               plain_final=field.numpy().copy())
    # field transformed (8 layers, weights x2 so that the flow is far from the identity)
    torch.manual_seed(3647)
    flow = ftlib.make_u1_equiv_layers(lattice_shape=(L, L), n_layers=8, n_mixture_comps=2, hidden_sizes=[8, 8], kernel_size=3)
    flow.eval()
    with torch.no_grad():
        for prm in flow.parameters():
            prm.mul_(2.0)
    for prm in flow.parameters():
        prm.requires_grad_(False)
    Pf = ref.Param(beta=2.0, lat=(L, L), tau=0.9, nstep=6)
    rec.update(weights=flat_weights(flow), ft_beta=2.0, ft_tau=0.9, ft_nstep=6)
    torch.manual_seed(seed)
    field = x0.clone()
    dHs, accs, plaqs, topos = [], [], [], []
    for i in range(ntraj):
        fr = torch.reshape(field, (1,) + field.shape)
        dH, e, acc, fr = quiet(ref.ft_hmc, Pf, flow, fr)
        field = fr[0]
        dHs.append(float(dH)); accs.append(bool(acc))
        plaqs.append(float(ref.action(Pf, field) / (-Pf.beta * Pf.volume))); topos.append(float(ref.topocharge(field)))
    rec.update(ft_dH=np.array(dHs), ft_acc=np.array(accs), ft_plaq=np.array(plaqs), ft_topo=np.array(topos),
               ft_final=field.numpy().copy())
    # statistics
    rng = np.random.default_rng(5)
    hist = np.cumsum(rng.integers(-1, 2, size=300)).astype(np.float64)
    rec["stat_hist"] = hist
    rec["stat_change_sqr_vs_dt"] = np.array(ref.change_sqr_vs_dt(list(hist[100:]), 10), dtype=np.float64)
    rec["stat_block_sizes"] = np.array([len(b) for b in ref.block_list(list(range(37)))])
    np.savez_compressed(os.path.join(out, "run_L8.npz"), **rec)
    print("run_L8: plain acc", np.mean(rec["plain_acc"]), "ft acc", np.mean(rec["ft_acc"]), "ft dH", rec["ft_dH"][:4])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--only", default="", help="generate only this fixture (e.g. run)")
    a = ap.parse_args()
    plain, ftlib, ref = load_reference(a.ref)
    if a.only == "run":
        gen_run(plain, ftlib, ref, a.out)
        return
    if a.only == "c2":
        gen_many(ftlib, ref, "ft_L16_b6_many", a.out, L=16, beta=6.0, legs=[(40, 200), (20, 30), (10, 30)], seed=20261016)
        return
    if a.only == "c3":
        gen_many(ftlib, ref, "ft_L32_b4_many", a.out, L=32, beta=4.0, legs=[(40, 200), (20, 20), (10, 40)], seed=20261032)
        return
    if a.only == "copyB":
        gen_copyB_physics(a.ref, a.out)
        return
    if a.only == "c4":
        gen_c4(ftlib, ref, a.out)
        return
    gen_run(plain, ftlib, ref, a.out)
    gen_plain(plain, a.out)
    gen_flow_case(ftlib, ref, "ft_L8_n8", a.out, L=8, beta=2.0, n_layers=8, B=3, nstep=6, ntraj=6, scale=2.0)
    gen_flow_case(ftlib, ref, "ft_L16_b6", a.out, L=16, beta=6.0, n_layers=24, B=2, nstep=10, ntraj=3)
    gen_flow_case(ftlib, ref, "ft_L32_b4", a.out, L=32, beta=4.0, n_layers=24, B=1, nstep=10, ntraj=2)
    gen_flow_case(ftlib, ref, "ft_L32_b4_n40", a.out, L=32, beta=4.0, n_layers=24, B=1, nstep=40, ntraj=1,
                  tau=1.0)
    gen_leaky(plain, a.out)
    gen_thousand(plain, ftlib, ref, a.out)
    gen_many(ftlib, ref, "ft_L16_b6_many", a.out, L=16, beta=6.0, legs=[(40, 200), (20, 30), (10, 30)], seed=20261016)
    gen_many(ftlib, ref, "ft_L32_b4_many", a.out, L=32, beta=4.0, legs=[(40, 200), (20, 20), (10, 40)], seed=20261032)
    gen_c4(ftlib, ref, a.out)
    try:
        gen_copyB_physics(a.ref, a.out)
        gen_copyB(a.ref, a.out)
    except Exception as e:  # copy B is optional (it drags in the package's logger/config)
        print("copyB generation failed:", repr(e))


if __name__ == "__main__":
    main()
