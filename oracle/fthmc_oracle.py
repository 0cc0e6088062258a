"""CPU oracle for the FT-HMC trajectory path of nftqcd/fthmc  --  TEST INFRASTRUCTURE ONLY.

This module is a plain torch-CPU fp64 restatement of the reference's "copy A" (the scripts
`hmc_2dU1.py`, `ipynb/field_transformation.py`, `ipynb/ft_hmc.py`).  It exists so that the CUDA
path in `fthmc_b200/` can be checked on a GPU box where `/root/reference` is absent.  Only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import it; the product package never does (and fails loudly without its CUDA library).

Parity pin: the reference ships no tests or golden vectors for this path ("parity unpinned" by
the reference itself).  The pin used here is the reference itself, imported in the build container
by `tests/golden/make_golden.py`, which dumps `tests/golden/*.npz`; `tests/test_oracle_golden.py`
replays those files against this module (bit-for-bit for everything that is a pure torch-op
sequence; the oracle deliberately issues the same torch ops in the same order).

Every function cites the reference file:line it follows (paths relative to the reference root).
The arithmetic lives in PyTorch (third-party, un-vendored; reference pins no version; 2.11.0 here).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as _dc_field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

TWO_PI = 2 * np.pi  # the reference writes 2*np.pi (field_transformation.py:18) / 2*math.pi (ft_hmc.py:107)

# ----------------------------------------------------------------------------------------------
# angle helpers
# ----------------------------------------------------------------------------------------------

def mod_2pi(x: torch.Tensor, convention: int = 0) -> torch.Tensor:
    """`torch_mod`.  convention 0 = copy A, [0,2pi)  (ipynb/field_transformation.py:17-18);
    convention 1 = copy B, [-pi,pi)  (fthmc/utils/layers.py:41-43)."""
    if convention == 0:
        return torch.remainder(x, TWO_PI)
    return torch.remainder(x + np.pi, TWO_PI) - np.pi


def wrap_pi(x: torch.Tensor) -> torch.Tensor:
    """`torch_wrap` (ipynb/field_transformation.py:19-20): [-pi,pi)."""
    return torch.remainder(x + np.pi, TWO_PI) - np.pi


def regularize(f: torch.Tensor) -> torch.Tensor:
    """hmc_2dU1.py:127-129 / ipynb/ft_hmc.py:109-112."""
    g = (f - math.pi) / (2 * math.pi)
    return (2 * math.pi) * (g - torch.floor(g) - 0.5)


# ----------------------------------------------------------------------------------------------
# plain Wilson action path (single chain, links (2,L0,L1))
# ----------------------------------------------------------------------------------------------

def plaqphase(f: torch.Tensor) -> torch.Tensor:
    """hmc_2dU1.py:114-120 / ipynb/ft_hmc.py:105.  Term order: ((t0 - t1) - t0(n+e1)) + t1(n+e0)."""
    t0, t1 = f[0, :], f[1, :]
    return t0 - t1 - torch.roll(t0, shifts=-1, dims=1) + torch.roll(t1, shifts=-1, dims=0)


def action(beta: float, f: torch.Tensor) -> torch.Tensor:
    """hmc_2dU1.py:100-101 / ipynb/ft_hmc.py:93-94."""
    return (-beta) * torch.sum(torch.cos(plaqphase(f)))


def force(beta: float, f: torch.Tensor) -> torch.Tensor:
    """hmc_2dU1.py:104-111 (autograd of `action`).  Works on a private leaf so the caller's
    tensor is not toggled (the reference flips `requires_grad` on its argument)."""
    leaf = f.detach().clone().requires_grad_(True)
    action(beta, leaf).backward()
    return leaf.grad.detach()


def force_closed_form(beta: float, f: torch.Tensor) -> torch.Tensor:
    """Closed form of the same derivative:
    F0 = beta[sin P(n) - sin P(n-e1)],  F1 = beta[-sin P(n) + sin P(n-e0)]."""
    s = torch.sin(plaqphase(f))
    f0 = beta * (s - torch.roll(s, shifts=1, dims=1))
    f1 = beta * (-s + torch.roll(s, shifts=1, dims=0))
    return torch.stack((f0, f1), dim=0)


def topocharge(f: torch.Tensor) -> torch.Tensor:
    """hmc_2dU1.py:123-124 / ipynb/ft_hmc.py:107: floor(0.1 + sum(regularize(P))/2pi)."""
    return torch.floor(0.1 + torch.sum(regularize(plaqphase(f))) / (2 * math.pi))


def leapfrog(beta: float, dt: float, nstep: int, x: torch.Tensor, p: torch.Tensor):
    """hmc_2dU1.py:132-141: position-first leapfrog with `nstep` force evaluations."""
    xx = x + 0.5 * dt * p
    pp = p + (-dt) * force(beta, xx)
    for _ in range(nstep - 1):
        xx = xx + dt * pp
        pp = pp + (-dt) * force(beta, xx)
    xx = xx + 0.5 * dt * pp
    return xx, pp


def hmc(beta: float, dt: float, nstep: int, x: torch.Tensor,
        p: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None):
    """hmc_2dU1.py:144-155.  `p`/`u` default to the same torch-RNG draws, in the same order, as the
    reference (randn_like before the leapfrog, rand([],float64) after it)."""
    if p is None:
        p = torch.randn_like(x)
    h0 = action(beta, x) + 0.5 * torch.sum(p * p)
    xx, pp = leapfrog(beta, dt, nstep, x, p)
    xr = regularize(xx)
    h1 = action(beta, xr) + 0.5 * torch.sum(pp * pp)
    if u is None:
        u = torch.rand([], dtype=torch.float64)
    dH = h1 - h0
    exp_mdH = torch.exp(-dH)
    acc = u < exp_mdH
    return dH, exp_mdH, acc, (xr if acc else x)


# ----------------------------------------------------------------------------------------------
# batched gauge helpers (links (B,2,L0,L1))
# ----------------------------------------------------------------------------------------------

def u1_plaq(links: torch.Tensor, order: int = 0) -> torch.Tensor:
    """`compute_u1_plaq(links,0,1)`.  order 0 = copy A (ipynb/field_transformation.py:116-119),
    term order ((t0 + t1(n+e0)) - t0(n+e1)) - t1;  order 1 = copy B (fthmc/utils/qed_helpers.py:80-86),
    term order ((t0 - t1) - t0(n+e1)) + t1(n+e0).  Same plaquette, last-ulp different."""
    if order == 1:
        return (links[:, 0] - links[:, 1]
                - torch.roll(links[:, 0], -1, 2) + torch.roll(links[:, 1], -1, 1))
    return (links[:, 0] + torch.roll(links[:, 1], -1, 1)
            - torch.roll(links[:, 0], -1, 2) - links[:, 1])


def u1_action(beta: float, cfgs: torch.Tensor) -> torch.Tensor:
    """`U1GaugeAction.__call__` (ipynb/field_transformation.py:120-130) for Nd=2 -> (B,)."""
    dens = 0 + torch.cos(u1_plaq(cfgs))
    return -beta * torch.sum(dens, dim=(1, 2))


def topo_charge(x: torch.Tensor) -> torch.Tensor:
    """Batched, un-rounded charge (ipynb/field_transformation.py:138-141)."""
    return torch.sum(wrap_pi(u1_plaq(x)), dim=(1, 2)) / TWO_PI


def gauge_transform(links: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    """ipynb/field_transformation.py:131-134 (out-of-place here)."""
    out = links.clone()
    for mu in range(2):
        out[:, mu] = alpha + links[:, mu] - torch.roll(alpha, -1, mu + 1)
    return out


# ----------------------------------------------------------------------------------------------
# flow description
# ----------------------------------------------------------------------------------------------

@dataclass
class LayerWeights:
    """One coupling layer's CNN: Conv2d(2,h0,k) act Conv2d(h0,h1,k) act Conv2d(h1,K+1,k)
    (ipynb/field_transformation.py:84-99), plus its mask parameters (:343-344)."""
    w: List[torch.Tensor]
    b: List[torch.Tensor]
    mu: int
    off: int


@dataclass
class Flow:
    layers: List[LayerWeights]
    activation: str = "silu"          # ipynb copy: SiLU (field_transformation.py:95); hmc_2dU1.py:260: LeakyReLU
    convention: int = 0               # 0: [0,2pi) copy A; 1: [-pi,pi) + copy-B plaquette term order
    inv_prec: float = 1e-6            # field_transformation.py:290
    inv_max_iter: int = 1000
    bisect_iters: List[int] = _dc_field(default_factory=list)  # diagnostics: iterations used per reverse call

    def __len__(self):
        return len(self.layers)


def flow_from_modulelist(flow_module, activation: str = "silu", convention: int = 0) -> Flow:
    """Read the weights out of a reference `nn.ModuleList` of `GaugeEquivCouplingLayer`
    (layer.plaq_coupling.net[0|2|4], ipynb/field_transformation.py:339-356)."""
    layers = []
    for i, layer in enumerate(flow_module):
        convs = [m for m in layer.plaq_coupling.net if hasattr(m, "weight")]
        layers.append(LayerWeights(
            w=[c.weight.detach().to(torch.float64).clone() for c in convs],
            b=[c.bias.detach().to(torch.float64).clone() for c in convs],
            mu=i % 2, off=(i // 2) % 4))
    return Flow(layers=layers, activation=activation, convention=convention)


def random_flow(n_layers: int = 24, n_mix: int = 2, hidden: Sequence[int] = (8, 8), ksize: int = 3,
                seed: int = 3647, activation: str = "silu", convention: int = 0,
                scale: float = 1.0) -> Flow:
    """Random-init flow with PyTorch's default Conv2d init (the reference's `set_weights(layers)` is a
    no-op on a ModuleList): weight, bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)).
    NOT the same stream as `torch.manual_seed(3647); make_u1_equiv_layers(...)`; golden files carry
    the reference-generated weights where that matters."""
    g = torch.Generator().manual_seed(seed)
    sizes = [2] + list(hidden) + [n_mix + 1]
    layers = []
    for i in range(n_layers):
        ws, bs = [], []
        for cin, cout in zip(sizes[:-1], sizes[1:]):
            bound = 1.0 / math.sqrt(cin * ksize * ksize)
            ws.append(scale * (torch.rand(cout, cin, ksize, ksize, generator=g, dtype=torch.float64) * 2 - 1) * bound)
            bs.append(scale * (torch.rand(cout, generator=g, dtype=torch.float64) * 2 - 1) * bound)
        layers.append(LayerWeights(w=ws, b=bs, mu=i % 2, off=(i // 2) % 4))
    return Flow(layers=layers, activation=activation, convention=convention)


# ----------------------------------------------------------------------------------------------
# masks (ipynb/field_transformation.py:175-248)
# ----------------------------------------------------------------------------------------------

def link_active_mask(shape: Tuple[int, int], mu: int, off: int) -> torch.Tensor:
    """`make_2d_link_active_stripes` (:175-197): (2,L0,L1) fp64, ones on the mu-links of every 4th
    line perpendicular to mu, shifted by `off`."""
    m = np.zeros((2,) + tuple(shape), dtype=np.uint8)
    if mu == 0:
        m[0, :, 0::4] = 1
    else:
        m[1, 0::4, :] = 1
    m = np.roll(m, off, axis=(1 - mu) + 1)
    return torch.from_numpy(m.astype(np.float64))


def _stripes(shape, mu, off, width):
    m = np.zeros(tuple(shape), dtype=np.uint8)
    for k in range(width):
        if mu == 0:
            m[:, k::4] = 1
        else:
            m[k::4, :] = 1
    return np.roll(m, off, axis=1 - mu)


def plaq_masks(shape: Tuple[int, int], mu: int, off: int):
    """`make_plaq_masks` (:243-248): frozen = double stripes at off+1, active = single stripes at
    off, passive = the rest.  uint8 tensors, like the reference."""
    frozen = _stripes(shape, mu, off + 1, 2)
    active = _stripes(shape, mu, off, 1)
    passive = 1 - frozen - active
    return (torch.from_numpy(active), torch.from_numpy(frozen), torch.from_numpy(passive))


# ----------------------------------------------------------------------------------------------
# coupling layer (ipynb/field_transformation.py:249-338, 152-174)
# ----------------------------------------------------------------------------------------------

def _activation(name: str, z: torch.Tensor) -> torch.Tensor:
    if name in ("silu", "swish"):
        return F.silu(z)
    if name == "leaky_relu":
        return F.leaky_relu(z, 0.01)
    if name == "relu":
        return F.relu(z)
    raise ValueError(name)


def cnn(lw: LayerWeights, activation: str, inp: torch.Tensor) -> torch.Tensor:
    """`make_conv_net` (:84-99): circular-padded 3x3 cross-correlations with the activation between.
    F.pad(circular)+conv2d is exactly what nn.Conv2d(padding_mode='circular') executes."""
    h = inp
    n = len(lw.w)
    for i in range(n):
        pad = lw.w[i].shape[-1] // 2
        h = F.conv2d(F.pad(h, (pad, pad, pad, pad), mode="circular"), lw.w[i], lw.b[i])
        if i != n - 1:
            h = _activation(activation, h)
    return h


def _tan_transform(x, s, conv):
    """:249-250"""
    return mod_2pi(2 * torch.atan(torch.exp(s) * torch.tan(x / 2)), conv)


def _tan_transform_logJ(x, s):
    """:252-253"""
    return -torch.log(torch.exp(-s) * torch.cos(x / 2) ** 2 + torch.exp(s) * torch.sin(x / 2) ** 2)


def _mixture(x, s, conv):
    """:254-257"""
    return torch.mean(_tan_transform(x, s, conv), dim=1, keepdim=True)


def _mixture_logJ(x, s):
    """:259-262"""
    return torch.logsumexp(_tan_transform_logJ(x, s), dim=1) - np.log(s.shape[1])


def _bisect(y, f, tol, max_iter, a, b):
    """`invert_transform_bisect` (:263-285).  Returns (mid_x, iterations_used).  The stop test is a
    max over the WHOLE tensor (all chains stop together in a batched call)."""
    lo = a * torch.ones_like(y)
    hi = b * torch.ones_like(y)
    lo_val = f(lo)
    hi_val = f(hi)
    mid = lo
    for it in range(max_iter):
        mid = (lo + hi) / 2
        mid_val = f(mid)
        gt = (y > mid_val).int().float()
        err = torch.max(torch.abs(y - mid_val))
        if err < tol:
            return mid, it + 1
        if torch.all((mid == lo) + (mid == hi)):
            return mid, it + 1
        lo = gt * mid + (1 - gt) * lo
        lo_val = gt * mid_val + (1 - gt) * lo_val
        hi = (1 - gt) * mid + gt * hi
        hi_val = (1 - gt) * mid_val + gt * hi_val
    return mid, max_iter


def plaq_coupling_forward(flow: Flow, lw: LayerWeights, x: torch.Tensor):
    """`NCPPlaqCouplingLayer.forward` (:300-317).  x: (B,L0,L1) plaquettes -> (new plaq, logJ (B,))."""
    act, frz, pas = plaq_masks(x.shape[1:], lw.mu, lw.off)
    x2 = frz * x
    net_out = cnn(lw, flow.activation, torch.stack((torch.cos(x2), torch.sin(x2)), dim=1))
    s, t = net_out[:, :-1], net_out[:, -1]
    x1 = (act * x).unsqueeze(1)
    local = act * _mixture_logJ(x1, s)
    logJ = torch.sum(local, dim=(1, 2))
    fx1 = act * _mixture(x1, s, flow.convention).squeeze(1)
    fx = act * mod_2pi(fx1 + t, flow.convention) + pas * x + frz * x
    return fx, logJ


def plaq_coupling_reverse(flow: Flow, lw: LayerWeights, fx: torch.Tensor):
    """`NCPPlaqCouplingLayer.reverse` (:319-338)."""
    act, frz, pas = plaq_masks(fx.shape[1:], lw.mu, lw.off)
    fx2 = frz * fx
    net_out = cnn(lw, flow.activation, torch.stack((torch.cos(fx2), torch.sin(fx2)), dim=1))
    s, t = net_out[:, :-1], net_out[:, -1]
    y = mod_2pi(act * (fx - t).unsqueeze(1), flow.convention)
    fwd = lambda z: act * _mixture(z, s, flow.convention)
    a, b = (0, TWO_PI) if flow.convention == 0 else (-np.pi, np.pi)   # layers.py:294 for copy B
    x1, iters = _bisect(y, fwd, flow.inv_prec, flow.inv_max_iter, a, b)
    flow.bisect_iters.append(iters)
    local = act * _mixture_logJ(x1, s)
    logJ = -torch.sum(local, dim=(1, 2))
    x1 = x1.squeeze(1)
    x = act * x1 + pas * fx + frz * fx2
    return x, logJ


def layer_forward(flow: Flow, lw: LayerWeights, x: torch.Tensor):
    """`GaugeEquivCouplingLayer.forward` (:160-166).  x: (B,2,L0,L1)."""
    m = link_active_mask(x.shape[2:], lw.mu, lw.off)
    plaq = u1_plaq(x, flow.convention)
    new_plaq, logJ = plaq_coupling_forward(flow, lw, plaq)
    d = new_plaq - plaq
    dl = torch.stack((d, -d), dim=1)
    return m * mod_2pi(dl + x, flow.convention) + (1 - m) * x, logJ


def layer_reverse(flow: Flow, lw: LayerWeights, fx: torch.Tensor):
    """`GaugeEquivCouplingLayer.reverse` (:168-174)."""
    m = link_active_mask(fx.shape[2:], lw.mu, lw.off)
    new_plaq = u1_plaq(fx, flow.convention)
    plaq, logJ = plaq_coupling_reverse(flow, lw, new_plaq)
    d = plaq - new_plaq
    dl = torch.stack((d, -d), dim=1)
    return m * mod_2pi(dl + fx, flow.convention) + (1 - m) * fx, logJ


# ----------------------------------------------------------------------------------------------
# FT-HMC (ipynb/ft_hmc.py:220-249, 394-435)
# ----------------------------------------------------------------------------------------------

def ft_flow(flow: Flow, f: torch.Tensor) -> torch.Tensor:
    """ipynb/ft_hmc.py:220-223"""
    for lw in flow.layers:
        f, _ = layer_forward(flow, lw, f)
    return f.detach()


def ft_flow_inv(flow: Flow, f: torch.Tensor) -> torch.Tensor:
    """ipynb/ft_hmc.py:225-228"""
    for lw in reversed(flow.layers):
        f, _ = layer_reverse(flow, lw, f)
    return f.detach()


def ft_flow_logJ(flow: Flow, f: torch.Tensor):
    """Forward flow that also returns the summed log-Jacobian (the `logJy` of ft_action, :231-235)."""
    logJ = 0.0
    for lw in flow.layers:
        f, lj = layer_forward(flow, lw, f)
        logJ = logJ + lj
    return f, logJ


def ft_action(beta: float, flow: Flow, f: torch.Tensor) -> torch.Tensor:
    """ipynb/ft_hmc.py:230-238: S(F(x)) - sum_layers logJ  -> (B,)."""
    y, logJ = ft_flow_logJ(flow, f)
    return u1_action(beta, y) - logJ


def ft_force(beta: float, flow: Flow, fld: torch.Tensor) -> torch.Tensor:
    """ipynb/ft_hmc.py:240-249 (autograd of sum(ft_action))."""
    leaf = fld.detach().clone().requires_grad_(True)
    s = torch.sum(ft_action(beta, flow, leaf))
    g, = torch.autograd.grad(s, leaf)
    return g.detach()


def ft_leapfrog(beta: float, dt: float, nstep: int, flow: Flow, x: torch.Tensor, p: torch.Tensor):
    """ipynb/ft_hmc.py:394-418 without the per-step diagnostics (norms, an extra ft_action; they do
    not feed back into the integration)."""
    xx = x + 0.5 * dt * p
    pp = p + (-dt) * ft_force(beta, flow, xx)
    for _ in range(nstep - 1):
        xx = xx + dt * pp
        pp = pp + (-dt) * ft_force(beta, flow, xx)
    xx = xx + 0.5 * dt * pp
    return xx, pp


def ft_hmc(beta: float, dt: float, nstep: int, flow: Flow, fld: torch.Tensor,
           p: Optional[torch.Tensor] = None, u: Optional[torch.Tensor] = None, details: bool = False):
    """ipynb/ft_hmc.py:420-435.  Single chain, fld (1,2,L0,L1).  Returns (dH, exp(-dH), acc, newfield)
    as (float, float, 0-d bool tensor, tensor) like the reference; `details=True` appends a dict with
    the latent start x, the proposal xr, H0 and H1."""
    x = ft_flow_inv(flow, fld)
    if p is None:
        p = torch.randn_like(x)
    h0 = ft_action(beta, flow, x).detach() + 0.5 * torch.sum(p * p)
    xx, pp = ft_leapfrog(beta, dt, nstep, flow, x, p)
    xr = regularize(xx)
    h1 = ft_action(beta, flow, xr).detach() + 0.5 * torch.sum(pp * pp)
    if u is None:
        u = torch.rand([], dtype=torch.float64)
    dH = h1 - h0
    exp_mdH = torch.exp(-dH)
    acc = u < exp_mdH
    newx = xr if acc else x
    newfield = ft_flow(flow, newx)
    out = (float(dH), float(exp_mdH), acc, newfield)
    if details:
        out = out + ({"x": x, "xr": xr, "p_end": pp, "h0": h0, "h1": h1},)
    return out


# ----------------------------------------------------------------------------------------------
# hand-derived adjoint of one coupling layer (what the CUDA kernel implements).
# Used by tests to validate the derivation against autograd; not part of the reference.
# ----------------------------------------------------------------------------------------------

def ft_force_adjoint(beta: float, flow: Flow, fld: torch.Tensor) -> torch.Tensor:
    """d/dx [ S(F(x)) - sum logJ ] by an explicit reverse sweep (no autograd)."""
    assert flow.activation in ("silu", "swish", "leaky_relu", "relu")
    K = flow.layers[0].w[-1].shape[0] - 1
    xs = [fld]
    for lw in flow.layers:
        y, _ = layer_forward(flow, lw, xs[-1])
        xs.append(y)
    yN = xs[-1]
    s = torch.sin(u1_plaq(yN))
    g0 = beta * (s - torch.roll(s, 1, 2))
    g1 = beta * (-s + torch.roll(s, 1, 1))
    g = torch.stack((g0, g1), dim=1)
    for lw, x in zip(reversed(flow.layers), reversed(xs[:-1])):
        act, frz, pas = [m.to(torch.float64) for m in plaq_masks(x.shape[2:], lw.mu, lw.off)]
        m = link_active_mask(x.shape[2:], lw.mu, lw.off)
        P = u1_plaq(x, flow.convention)
        # recompute the CNN, keeping pre-activations
        inp = torch.stack((torch.cos(frz * P), torch.sin(frz * P)), dim=1)
        zs, hs = [], [inp]
        for i in range(len(lw.w)):
            z = F.conv2d(F.pad(hs[-1], (1, 1, 1, 1), mode="circular"), lw.w[i], lw.b[i])
            zs.append(z)
            hs.append(_activation(flow.activation, z) if i != len(lw.w) - 1 else z)
        out = zs[-1]
        sk, t = out[:, :K], out[:, K]
        u = (act * P).unsqueeze(1)
        c2, s2 = torch.cos(u / 2) ** 2, torch.sin(u / 2) ** 2
        den = torch.exp(-sk) * c2 + torch.exp(sk) * s2
        el = 1.0 / den                                    # e^{l_k}
        sig = torch.softmax(-torch.log(den), dim=1)
        dbar = (m[0] * g[:, 0] - m[1] * g[:, 1])          # delta-bar on active plaquettes
        w = -1.0
        ubar = act * (dbar * el.mean(dim=1) + w * (sig * (-torch.sinh(sk) * torch.sin(u) * el)).sum(dim=1))
        sbar = act.unsqueeze(0).unsqueeze(0) * (dbar.unsqueeze(1) * torch.sin(u) * el / K
                                                + w * sig * (torch.exp(-sk) * c2 - torch.exp(sk) * s2) * el)
        tbar = act * dbar
        hbar = torch.cat((sbar, tbar.unsqueeze(1)), dim=1)
        for i in reversed(range(len(lw.w))):
            if i != len(lw.w) - 1:
                z = zs[i]
                if flow.activation in ("silu", "swish"):
                    sg = torch.sigmoid(z)
                    hbar = hbar * (sg * (1 + z * (1 - sg)))
                elif flow.activation == "leaky_relu":
                    hbar = hbar * torch.where(z > 0, torch.ones_like(z), 0.01 * torch.ones_like(z))
                else:
                    hbar = hbar * (z > 0).to(z.dtype)
            wt = torch.flip(lw.w[i], dims=(2, 3)).transpose(0, 1)
            hbar = F.conv2d(F.pad(hbar, (1, 1, 1, 1), mode="circular"), wt)
        Pbar = -act * dbar + ubar + frz * (-torch.sin(frz * P) * hbar[:, 0] + torch.cos(frz * P) * hbar[:, 1])
        x0 = g[:, 0] + Pbar - torch.roll(Pbar, 1, 2)
        x1 = g[:, 1] - Pbar + torch.roll(Pbar, 1, 1)
        g = torch.stack((x0, x1), dim=1)
    return g


def ft_action_weight_grad(beta: float, flow: Flow, fld: torch.Tensor):
    """The gradient the reference's reverse-KL train_step back-propagates (ipynb/ft_hmc.py:253-295: loss = mean(logq - logp)
    = mean_b ft_action(xi_b) + const): d/d(weights) of sum_b ft_action(x_b) by torch.autograd, flattened per layer in the
    parameter order of layer.plaq_coupling.net (conv0.weight, conv0.bias, conv1.weight, ...): (n_layers, 955).
    Also returns ft_action (B,)."""
    leaves = []
    layers = []
    for lw in flow.layers:
        w = [t.detach().clone().requires_grad_(True) for t in lw.w]
        b = [t.detach().clone().requires_grad_(True) for t in lw.b]
        leaves.append((w, b))
        layers.append(LayerWeights(w=w, b=b, mu=lw.mu, off=lw.off))
    f2 = Flow(layers=layers, activation=flow.activation, convention=flow.convention)
    act = ft_action(beta, f2, fld.detach())
    flat = [t for w, b in leaves for pair in zip(w, b) for t in pair]
    grads = torch.autograd.grad(torch.sum(act), flat)
    per = len(flat) // len(flow.layers)
    rows = [torch.cat([g.reshape(-1) for g in grads[i * per:(i + 1) * per]]) for i in range(len(flow.layers))]
    return act.detach(), torch.stack(rows)
