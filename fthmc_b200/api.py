"""Drop-in entry points of the FT-HMC trajectory path, same names / argument order / return structure
as the reference scripts (copy A):

    action force plaqphase-free topocharge regularize leapfrog hmc        hmc_2dU1.py:100-155
    topo_charge                                                           ipynb/field_transformation.py:138
    ft_flow ft_flow_inv ft_action ft_force                                ipynb/ft_hmc.py:220-249
    ft_leapfrog ft_hmc                                                    ipynb/ft_hmc.py:394-435

Tensors keep the reference layouts: links (2,L0,L1) for the single-chain plain functions, (B,2,L0,L1)
for the flow functions.  CPU tensors are accepted like in the reference (copied to the GPU, result
copied back); CUDA tensors stay on the device.  Bodies are CUDA kernels reached through the C ABI of
libfthmc_b200.so -- there is no PyTorch/CPU fallback.

Batched extensions (`*_batch`) take explicit momenta `p` and uniforms `u` (parity mode) or a Philox
seed (throughput mode) and return per-chain dH / acc / plaq / Q."""
import math
import os
from functools import reduce

import torch

from . import _lib
from .flow import PackedFlow, pack

F64, F32 = 0, 1


class Param:
    """Same fields as the reference's Param (ipynb/ft_hmc.py:57-72, hmc_2dU1.py:41-70)."""

    def __init__(self, beta=6.0, lat=(64, 64), tau=2.0, nstep=50, ntraj=256, nrun=4, nprint=256, seed=11 * 13,
                 randinit=False, nth=int(os.environ.get("OMP_NUM_THREADS", "2")), nth_interop=2):
        self.beta = beta
        self.lat = tuple(lat)
        self.nd = len(lat)
        self.volume = reduce(lambda x, y: x * y, lat)
        self.tau = tau
        self.nstep = nstep
        self.dt = self.tau / self.nstep
        self.ntraj, self.nrun, self.nprint, self.seed = ntraj, nrun, nprint, seed
        self.randinit, self.nth, self.nth_interop = randinit, nth, nth_interop

    def initializer(self):
        # float64 regardless of torch's default dtype: the reference scripts set the default tensor type to double
        # (hmc_2dU1.py:684, ipynb/ft_hmc.py:522), and the momenta drawn by randn_like follow this dtype's RNG stream
        if self.randinit:
            return torch.empty((self.nd,) + self.lat, dtype=torch.float64).uniform_(-math.pi, math.pi)
        return torch.zeros((self.nd,) + self.lat, dtype=torch.float64)


# ------------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------------
import threading
_tls = threading.local()


def _device(t=None):
    if t is not None and t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("fthmc_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _workspace(flow_handle, B, L0, L1, dev, need=None, kind="chain"):
    """Scratch for one launch.  The kernels keep per-CTA state in it (momenta, the layer blocks of the adjoint, gradient
    accumulators), so two launches in flight must never share one: buffers are cached per (device, STREAM, kind) -- launches
    on one stream are ordered, launches on different streams get different buffers -- and per Python thread.  A buffer is
    allocated while its stream is current, so the caching allocator's stream-ordered reuse makes replacing it safe."""
    if need is None:
        need = _lib.lib().fthmc_workspace_bytes(flow_handle, B, L0, L1)
    cache = getattr(_tls, "ws", None)
    if cache is None:
        cache = _tls.ws = {}
    key = (dev.index, _stream(), kind)
    buf = cache.get(key)
    if buf is None or buf.numel() < need:
        if buf is not None:
            buf.record_stream(torch.cuda.current_stream())
        buf = torch.empty(need, dtype=torch.uint8, device=dev)
        cache[key] = buf
    return buf


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev_in(t, dev, dtype=None):
    """contiguous device copy (no copy if already there)."""
    t = t.detach()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(dev, non_blocking=True).contiguous()


def _back(out, like):
    """result on the device `like` lives on.  Host results land in page-locked memory (torch's caching host
    allocator recycles the blocks), so a D2H copy runs at link speed and a result that is fed back as the next
    input -- the run loops do exactly that with the field -- is also a fast H2D source."""
    if out.is_floating_point() and like.is_floating_point() and out.dtype != like.dtype:
        out = out.to(like.dtype)          # fp32 callers (the package copy's default dtype) get fp32 back; the kernels compute in fp64
    if like.is_cuda:
        return out
    host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return host


def _dt(t):
    if t.dtype == torch.float64:
        return F64
    if t.dtype == torch.float32:
        return F32
    raise _lib.FthmcError(-3, f"unsupported dtype {t.dtype}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _as_batch(f):
    """(2,L0,L1) -> (1,2,L0,L1); (B,2,L0,L1) unchanged."""
    if f.dim() == 3:
        return f.unsqueeze(0), True
    if f.dim() == 4:
        return f, False
    raise _lib.FthmcError(-1, f"links must be (2,L0,L1) or (B,2,L0,L1), got {tuple(f.shape)}")


# ------------------------------------------------------------------------------------------------
# plain Wilson stencils
# ------------------------------------------------------------------------------------------------
def _reduce_call(fn_name, f, order_or_rounded, beta=None):
    fb, single = _as_batch(f)
    dev = _device(fb)
    with torch.cuda.device(dev):
        x = _dev_in(fb, dev)
        B, _, L0, L1 = x.shape
        out = torch.empty(B, dtype=x.dtype, device=dev)
        L = _lib.lib()
        if fn_name == "action":
            _lib.check(L.fthmc_action(x.data_ptr(), B, L0, L1, float(beta), order_or_rounded, out.data_ptr(), _dt(x), _stream()))
        else:
            _lib.check(L.fthmc_topo_charge(x.data_ptr(), B, L0, L1, order_or_rounded, out.data_ptr(), _dt(x), _stream()))
    out = _back(out, f)
    return out[0] if single else out


def action(param, f):
    """hmc_2dU1.py:100 -- -beta * sum cos(plaqphase(f)); (2,L0,L1) -> 0-d (a (B,2,L0,L1) batch gives (B,))."""
    return _reduce_call("action", f, 1, param.beta)


def u1_action(beta, cfgs):
    """U1GaugeAction(beta)(cfgs), ipynb/field_transformation.py:120 -- (B,2,L0,L1) -> (B,)."""
    return _reduce_call("action", cfgs, 0, beta)


def topocharge(f):
    """hmc_2dU1.py:123 -- floor(0.1 + sum regularize(P)/2pi)."""
    return _reduce_call("topo", f, 1)


def topo_charge(x):
    """ipynb/field_transformation.py:138 -- batched, un-rounded."""
    return _reduce_call("topo", x, 0)


def force(param, f, order=1):
    """hmc_2dU1.py:104 -- dS/df (the reference gets it from autograd; closed form here)."""
    fb, single = _as_batch(f)
    dev = _device(fb)
    with torch.cuda.device(dev):
        x = _dev_in(fb, dev)
        B, _, L0, L1 = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.lib().fthmc_force(x.data_ptr(), B, L0, L1, float(param.beta), order, out.data_ptr(), _dt(x), _stream()))
    out = _back(out, f)
    return out[0] if single else out


def regularize(f):
    """hmc_2dU1.py:127"""
    dev = _device(f)
    with torch.cuda.device(dev):
        x = _dev_in(f, dev)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().fthmc_regularize(x.data_ptr(), out.data_ptr(), x.numel(), _dt(x), _stream()))
    return _back(out, f)


# ------------------------------------------------------------------------------------------------
# plain HMC
# ------------------------------------------------------------------------------------------------
def leapfrog(param, x, p):
    """hmc_2dU1.py:132 -- returns (x_, p_)."""
    xb, single = _as_batch(x)
    pb, _ = _as_batch(p)
    dev = _device(xb)
    with torch.cuda.device(dev):
        xd, pd = _dev_in(xb, dev, torch.float64), _dev_in(pb, dev, torch.float64)
        B, _, L0, L1 = xd.shape
        xo, po = torch.empty_like(xd), torch.empty_like(pd)
        ws = _workspace(None, B, L0, L1, dev)
        _lib.check(_lib.lib().fthmc_leapfrog(xd.data_ptr(), pd.data_ptr(), xo.data_ptr(), po.data_ptr(), B, L0, L1,
                                             float(param.beta), float(param.dt), int(param.nstep),
                                             ws.data_ptr(), ws.numel(), _stream()))
    xo, po = _back(xo, x), _back(po, x)
    return (xo[0], po[0]) if single else (xo, po)


_side_streams = {}


def _pipelined_ok(x, p, u, flow_pf):
    """Host batches of several device waves go through the chunked path: the copies of chunk i+1 / i-1 run on side
    streams under the kernel of chunk i."""
    if flow_pf is None or x.is_cuda or x.dtype != torch.float64 or x.dim() != 4 or not x.is_contiguous():
        return False
    for t in (p, u):
        if t is not None and (t.is_cuda or t.dtype != torch.float64 or not t.is_contiguous()):
            return False
    nsm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return x.shape[0] >= 8 * nsm


def _traj_call_pipelined(flow_pf, beta, dt, nstep, x, p, u, seed, traj, chain0, want_h):
    """FT-HMC trajectories of a HOST batch in (up to) four chunks of whole device waves.  The chains are independent and
    the device RNG is keyed by the global chain index, so the chunks reproduce the single launch (fields and decisions
    exactly; per-chain sums to rounding where the CTA width depends on the launch's batch size); what changes is that only
    the first chunk's host-to-device copy and the last chunk's device-to-host copy are exposed."""
    dev = _device(x)
    with torch.cuda.device(dev):
        B, _, L0, L1 = x.shape
        nsm = torch.cuda.get_device_properties(dev).multi_processor_count
        waves = -(-B // nsm)
        per = -(-waves // 4) * nsm
        bounds = [(a, min(B, a + per)) for a in range(0, B, per)]
        main = torch.cuda.current_stream()
        if dev not in _side_streams:
            _side_streams[dev] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        s_in, s_out = _side_streams[dev]
        if u is not None:
            u = u.reshape(-1)
            if u.numel() != B:
                raise _lib.FthmcError(-1, "u must have one entry per chain")
        if p is not None and p.shape != x.shape:
            raise _lib.FthmcError(-1, "p must have the shape of x")
        xd, xo = torch.empty_like(x, device=dev), torch.empty_like(x, device=dev)
        pd = None if p is None else torch.empty_like(p, device=dev)
        ud = None if u is None else torch.empty_like(u, device=dev)
        sc = torch.empty((6, B), dtype=torch.float64, device=dev)      # dH, exp(-dH), plaq, Q, h0, h1
        acc = torch.empty(B, dtype=torch.int32, device=dev)
        h_field = torch.empty(x.shape, dtype=torch.float64, pin_memory=True)
        h_sc = torch.empty((6, B), dtype=torch.float64, pin_memory=True)
        h_acc = torch.empty(B, dtype=torch.int32, pin_memory=True)
        L = _lib.lib()
        ws = _workspace(flow_pf.handle, min(B, per), L0, L1, dev)
        s_in.wait_stream(main)
        s_out.wait_stream(main)

        def copy_in(a, b):                                   # (a pageable source blocks the host here, under the running kernel)
            with torch.cuda.stream(s_in):
                xd[a:b].copy_(x[a:b], non_blocking=True)
                if pd is not None:
                    pd[a:b].copy_(p[a:b], non_blocking=True)
                if ud is not None:
                    ud[a:b].copy_(u[a:b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
            return e

        ev = copy_in(*bounds[0])
        for i, (a, b) in enumerate(bounds):
            main.wait_event(ev)
            _lib.check(L.fthmc_ft_hmc_traj(flow_pf.handle, xd[a:b].data_ptr(), xo[a:b].data_ptr(),
                                           None if pd is None else pd[a:b].data_ptr(), None if ud is None else ud[a:b].data_ptr(),
                                           seed, traj, chain0 + a, b - a, L0, L1, float(beta), float(dt), int(nstep),
                                           sc[0, a:b].data_ptr(), sc[1, a:b].data_ptr(), acc[a:b].data_ptr(),
                                           sc[2, a:b].data_ptr(), sc[3, a:b].data_ptr(),
                                           sc[4, a:b].data_ptr() if want_h else None, sc[5, a:b].data_ptr() if want_h else None,
                                           ws.data_ptr(), ws.numel(), main.cuda_stream))
            k = torch.cuda.Event()
            k.record(main)
            if i + 1 < len(bounds):
                ev = copy_in(*bounds[i + 1])
            with torch.cuda.stream(s_out):
                s_out.wait_event(k)
                h_field[a:b].copy_(xo[a:b], non_blocking=True)
                for r in range(6 if want_h else 4):
                    h_sc[r, a:b].copy_(sc[r, a:b], non_blocking=True)
                h_acc[a:b].copy_(acc[a:b], non_blocking=True)
        main.wait_stream(s_out)
        main.synchronize()
    res = dict(field=h_field, dH=h_sc[0], exp_mdH=h_sc[1], acc=h_acc.bool(), plaq=h_sc[2], topo=h_sc[3])
    if want_h:
        res.update(h0=h_sc[4], h1=h_sc[5])
    return res


def _traj_call(flow_pf, beta, dt, nstep, x, p, u, seed, traj, chain0, want_h=False):
    """shared body of hmc_batch / ft_hmc_batch.  x (B,2,L0,L1) on any device."""
    if _pipelined_ok(x, p, u, flow_pf):
        return _traj_call_pipelined(flow_pf, beta, dt, nstep, x, p, u, seed, traj, chain0, want_h)
    dev = _device(x)
    with torch.cuda.device(dev):
        xd = _dev_in(x, dev, torch.float64)
        B, _, L0, L1 = xd.shape
        pd = None if p is None else _dev_in(p, dev, torch.float64)
        ud = None if u is None else _dev_in(u.reshape(-1), dev, torch.float64)
        if pd is not None and pd.shape != xd.shape:
            raise _lib.FthmcError(-1, "p must have the shape of x")
        if ud is not None and ud.numel() != B:
            raise _lib.FthmcError(-1, "u must have one entry per chain")
        xo = torch.empty_like(xd)
        sc = torch.empty((6, B), dtype=torch.float64, device=dev)      # dH, exp(-dH), plaq, Q, h0, h1
        acc = torch.empty(B, dtype=torch.int32, device=dev)
        L = _lib.lib()
        if flow_pf is None:
            ws = _workspace(None, B, L0, L1, dev)
            _lib.check(L.fthmc_hmc_traj(xd.data_ptr(), xo.data_ptr(), _ptr(pd), _ptr(ud), seed, traj, chain0, B, L0, L1,
                                        float(beta), float(dt), int(nstep), sc[0].data_ptr(), sc[1].data_ptr(),
                                        acc.data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        else:
            ws = _workspace(flow_pf.handle, B, L0, L1, dev)
            _lib.check(L.fthmc_ft_hmc_traj(flow_pf.handle, xd.data_ptr(), xo.data_ptr(), _ptr(pd), _ptr(ud), seed, traj, chain0,
                                           B, L0, L1, float(beta), float(dt), int(nstep), sc[0].data_ptr(), sc[1].data_ptr(),
                                           acc.data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(),
                                           sc[4].data_ptr() if want_h else None, sc[5].data_ptr() if want_h else None,
                                           ws.data_ptr(), ws.numel(), _stream()))
    res = dict(field=_back(xo, x), dH=_back(sc[0], x), exp_mdH=_back(sc[1], x), acc=_back(acc, x).bool(),
               plaq=_back(sc[2], x), topo=_back(sc[3], x))
    if want_h:
        res.update(h0=_back(sc[4], x), h1=_back(sc[5], x))
    return res


def hmc_batch(param, x, p=None, u=None, seed=0, traj=0, chain0=0):
    """B independent plain-HMC trajectories in one persistent launch.  p/u None => device Philox."""
    return _traj_call(None, param.beta, param.dt, param.nstep, x, p, u, seed, traj, chain0)


def hmc(param, x):
    """hmc_2dU1.py:144 -- (dH, exp_mdH, acc, newx).  Momenta and the Metropolis uniform come from the
    torch generator in the reference's order (randn_like(x), then rand([], float64)), so a seeded run
    reproduces the reference chain."""
    p = torch.randn_like(x)
    u = torch.rand([], dtype=torch.float64)
    r = hmc_batch(param, x.unsqueeze(0), p.unsqueeze(0), u.reshape(1))
    return r["dH"][0], r["exp_mdH"][0], r["acc"][0], r["field"][0]


# ------------------------------------------------------------------------------------------------
# field transformation
# ------------------------------------------------------------------------------------------------
def _flow_call(kind, flow, f, beta=0.0, want_logJ=False, convention=0):
    dev = _device(f)
    with torch.cuda.device(dev):
        pf = pack(flow, convention=convention, device=dev)
        xd = _dev_in(f, dev, torch.float64)
        if xd.dim() != 4:
            raise _lib.FthmcError(-1, f"field must be (B,2,L0,L1), got {tuple(xd.shape)}")
        B, _, L0, L1 = xd.shape
        ws = _workspace(pf.handle, B, L0, L1, dev)
        L = _lib.lib()
        logJ = torch.empty(B, dtype=torch.float64, device=dev) if (want_logJ or kind == "action") else None
        out = torch.empty_like(xd) if kind != "action" else None
        if kind == "fwd":
            _lib.check(L.fthmc_flow_fwd(pf.handle, xd.data_ptr(), out.data_ptr(), _ptr(logJ), None, B, L0, L1,
                                        ws.data_ptr(), ws.numel(), _stream()))
        elif kind == "inv":
            _lib.check(L.fthmc_flow_inv(pf.handle, xd.data_ptr(), out.data_ptr(), _ptr(logJ), None, None, B, L0, L1,
                                        ws.data_ptr(), ws.numel(), _stream()))
        elif kind == "action":
            _lib.check(L.fthmc_ft_action(pf.handle, xd.data_ptr(), float(beta), logJ.data_ptr(), None, B, L0, L1,
                                         ws.data_ptr(), ws.numel(), _stream()))
        elif kind == "force":
            _lib.check(L.fthmc_ft_force(pf.handle, xd.data_ptr(), float(beta), out.data_ptr(), B, L0, L1,
                                        ws.data_ptr(), ws.numel(), _stream()))
    if kind == "action":
        return _back(logJ, f)
    if want_logJ:
        return _back(out, f), _back(logJ, f)
    return _back(out, f)


def ft_flow(flow, f, with_logJ=False):
    """ipynb/ft_hmc.py:220 -- F(f) (detached).  with_logJ=True also returns sum_layers logJ (B,)."""
    return _flow_call("fwd", flow, f, want_logJ=with_logJ)


def ft_flow_inv(flow, f, with_logJ=False):
    """ipynb/ft_hmc.py:225 -- F^{-1}(f) by bisection to 1e-6, decision for decision.  The reference's stop test is the maximum
    error over the WHOLE tensor it is given; its trajectory path always passes one chain, and the kernel applies the test
    per chain.  For B > 1 the reference would keep halving every chain until the slowest one converges, so a batched call
    here can stop a chain a few iterations earlier than a batched reference call would (difference <= 1e-6, the bisection
    tolerance); decision-for-decision parity is for B = 1 calls, i.e. `ft_flow_inv(flow, f[b:b+1])` of the reference."""
    return _flow_call("inv", flow, f, want_logJ=with_logJ)


def ft_action(param, flow, f):
    """ipynb/ft_hmc.py:230 -- S(F(f)) - sum logJ, (B,)."""
    return _flow_call("action", flow, f, beta=param.beta)


def ft_force(param, flow, field, create_graph=False):
    """ipynb/ft_hmc.py:240 -- d/dfield sum(ft_action): hand-written adjoint kernel, no autograd.  The reference's
    create_graph=True exists for one purpose, the force-norm training loss (ipynb/ft_hmc.py:266-269: loss = sum(force^2),
    loss.backward()); a kernel cannot hand back an autograd graph, so that use is served by `ft_force_norm_grad`, which
    returns the loss and its gradient with respect to the weights directly."""
    if create_graph:
        raise NotImplementedError("no autograd graph comes out of a CUDA kernel: use ft_force_norm_grad(param, flow, field) for the "
                                  "force-norm loss and its weight gradient (FlowTrainer.train_step(with_force=True) does)")
    return _flow_call("force", flow, field, beta=param.beta)


def ft_leapfrog(param, flow, x, p):
    """ipynb/ft_hmc.py:394 -- returns (x_, p_) (the reference's per-step prints are not reproduced)."""
    dev = _device(x)
    with torch.cuda.device(dev):
        pf = pack(flow, device=dev)
        xd, pd = _dev_in(x, dev, torch.float64), _dev_in(p, dev, torch.float64)
        B, _, L0, L1 = xd.shape
        xo, po = torch.empty_like(xd), torch.empty_like(pd)
        ws = _workspace(pf.handle, B, L0, L1, dev)
        _lib.check(_lib.lib().fthmc_ft_leapfrog(pf.handle, xd.data_ptr(), pd.data_ptr(), xo.data_ptr(), po.data_ptr(),
                                                B, L0, L1, float(param.beta), float(param.dt), int(param.nstep),
                                                ws.data_ptr(), ws.numel(), _stream()))
    return _back(xo, x), _back(po, x)


def ft_hmc_batch(param, flow, field, p=None, u=None, seed=0, traj=0, chain0=0, want_h=False):
    """B independent FT-HMC trajectories, one persistent CTA per chain.  p/u None => device Philox
    keyed by (seed, chain0+b, traj)."""
    dev = _device(field)
    with torch.cuda.device(dev):
        pf = pack(flow, device=dev)
    return _traj_call(pf, param.beta, param.dt, param.nstep, field, p, u, seed, traj, chain0, want_h=want_h)


def ft_hmc(param, flow, field):
    """ipynb/ft_hmc.py:420 -- (float dH, float exp_mdH, 0-d bool acc, newfield), field (1,2,L0,L1)."""
    p = torch.randn_like(field)
    u = torch.rand([], dtype=torch.float64)
    r = ft_hmc_batch(param, flow, field, p, u.reshape(1))
    return float(r["dH"][0]), float(r["exp_mdH"][0]), r["acc"][0], r["field"]


# ------------------------------------------------------------------------------------------------
# run loops: many trajectories per launch, the chains resident in shared memory
# ------------------------------------------------------------------------------------------------
def _run_call(flow_pf, beta, dt, nstep, ntraj, x, p, u, seed, traj0, chain0):
    dev = _device(x)
    with torch.cuda.device(dev):
        xd = _dev_in(x, dev, torch.float64)
        if xd.dim() != 4:
            raise _lib.FthmcError(-1, f"field must be (B,2,L0,L1), got {tuple(xd.shape)}")
        B, _, L0, L1 = xd.shape
        pd = None if p is None else _dev_in(p, dev, torch.float64)
        ud = None if u is None else _dev_in(u, dev, torch.float64)
        if pd is not None and tuple(pd.shape) != (ntraj,) + tuple(xd.shape):
            raise _lib.FthmcError(-1, "p must be (ntraj,) + field.shape")
        if ud is not None and tuple(ud.shape) != (ntraj, B):
            raise _lib.FthmcError(-1, "u must be (ntraj, B)")
        xo = torch.empty_like(xd)
        sc = torch.empty((4, ntraj, B), dtype=torch.float64, device=dev)      # dH, exp(-dH), plaq, Q
        acc = torch.empty((ntraj, B), dtype=torch.int32, device=dev)
        L = _lib.lib()
        handle = None if flow_pf is None else flow_pf.handle
        ws = _workspace(handle, B, L0, L1, dev)
        if flow_pf is None:
            _lib.check(L.fthmc_hmc_run(xd.data_ptr(), xo.data_ptr(), _ptr(pd), _ptr(ud), seed, traj0, chain0, B, L0, L1,
                                       float(beta), float(dt), int(nstep), int(ntraj), sc[0].data_ptr(), sc[1].data_ptr(),
                                       acc.data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
        else:
            _lib.check(L.fthmc_ft_hmc_run(handle, xd.data_ptr(), xo.data_ptr(), _ptr(pd), _ptr(ud), seed, traj0, chain0,
                                          B, L0, L1, float(beta), float(dt), int(nstep), int(ntraj), sc[0].data_ptr(),
                                          sc[1].data_ptr(), acc.data_ptr(), sc[2].data_ptr(), sc[3].data_ptr(),
                                          ws.data_ptr(), ws.numel(), _stream()))
    return dict(field=_back(xo, x), dH=_back(sc[0], x), exp_mdH=_back(sc[1], x), acc=_back(acc, x).bool(),
                plaq=_back(sc[2], x), topo=_back(sc[3], x))


def hmc_run_batch(param, x, ntraj, p=None, u=None, seed=0, traj0=0, chain0=0):
    """`ntraj` consecutive plain-HMC trajectories of B chains in ONE launch (the loop of hmc_2dU1.py:697-707 with the
    field resident on the SM).  Per-trajectory results are (ntraj, B); p (ntraj,B,2,L0,L1) / u (ntraj,B) optional."""
    return _run_call(None, param.beta, param.dt, param.nstep, ntraj, x, p, u, seed, traj0, chain0)


def ft_hmc_run_batch(param, flow, field, ntraj, p=None, u=None, seed=0, traj0=0, chain0=0):
    """`ntraj` consecutive FT-HMC trajectories of B chains in ONE launch (the loop of ipynb/ft_hmc.py:454-467)."""
    dev = _device(field)
    with torch.cuda.device(dev):
        pf = pack(flow, device=dev)
    return _run_call(pf, param.beta, param.dt, param.nstep, ntraj, field, p, u, seed, traj0, chain0)


def _status_lines(first, r, b=0):
    out = []
    for i in range(r["dH"].shape[0]):
        ifacc = "ACCEPT" if bool(r["acc"][i, b]) else "REJECT"
        out.append(f"Traj: {first + i + 1:4}  {ifacc}  dH: {float(r['dH'][i, b]):< 12.8}  exp(-dH): {float(r['exp_mdH'][i, b]):< 12.8}  "
                   f"plaq: {float(r['plaq'][i, b]):< 12.8}  topo: {float(r['topo'][i, b]):< 3.3}\n")
    return out


def _run_loop(param, flow, field, out, topo_history):
    """shared body of run / ft_run: param.nrun blocks of param.ntraj trajectories, one launch per block.  Momenta and
    Metropolis uniforms are drawn from the torch generator in the reference's order (randn_like, then rand([]), per
    trajectory), so a seeded run follows the reference chain."""
    import sys
    from timeit import default_timer as timer
    put = (lambda s: None) if out is None else out.write
    plaq, topo = action(param, field) / (-param.beta * param.volume), topocharge(field)
    put(f"Initial configuration:  plaq: {plaq}  topo: {topo} {field.shape}\n")
    ts = []
    for n in range(param.nrun):
        t = -timer()
        ps, us = [], []
        for _ in range(param.ntraj):
            ps.append(torch.randn_like(field))
            us.append(torch.rand([], dtype=torch.float64))
        p, u = torch.stack(ps).unsqueeze(1), torch.stack(us).reshape(-1, 1)
        if flow is None:
            r = hmc_run_batch(param, field.unsqueeze(0), param.ntraj, p, u)
        else:
            r = ft_hmc_run_batch(param, flow, field.unsqueeze(0), param.ntraj, p, u)
        field = r["field"][0]
        for line in _status_lines(n * param.ntraj, r):
            put(line)
        topo_history.extend(float(v) for v in r["topo"][:, 0])
        t += timer()
        ts.append(t)
    put(f"Run times:  {ts}\n")
    put(f"Per trajectory:  {[t / param.ntraj for t in ts]}\n")
    if out is sys.stdout:
        sys.stdout.flush()
    return field


topo_history = []


def run(param, field=None, out=None):
    """The trajectory loop of run(param, field) (ipynb/ft_hmc.py:180-216, hmc_2dU1.py:686-715): nrun x ntraj plain-HMC
    trajectories of one chain (2,L0,L1), the status line of every trajectory written to `out` (a file-like object;
    None = quiet), the charges appended to `topo_history`.  Returns the final field.  The reference's result-file
    bookkeeping (uniquestr / skip-if-exists) stays with the caller."""
    if field is None:
        field = param.initializer()
    topo_history.clear()
    return _run_loop(param, None, field, out, topo_history)


def ft_run(param, flow, field=None, out=None):
    """The trajectory loop of ft_run(param, flow, field) (ipynb/ft_hmc.py:437-475) for one chain (2,L0,L1)."""
    if field is None:
        field = param.initializer()
    topo_history.clear()
    return _run_loop(param, flow, field, out, topo_history)


def run_hmc(param, x=None, out=None):
    """run_hmc(param, x) (fthmc/hmc.py:57-175) without its directories, plots and dumps: `param.nrun` independent
    experiments of `param.ntraj` plain-HMC trajectories each (every experiment restarts from `param.initializer()`, as the
    reference does, unless x is given), one kernel launch per experiment with the chain resident on the SM.  Returns
    (fields_arr, histories) like the reference: histories[n] has the lists traj / dt / acc / dH / plaq / q / dq; fields_arr[n]
    holds the final field of experiment n (the reference keeps every intermediate field; the resident kernel does not
    write them out).  Momenta and uniforms come from the torch generator in the reference's order."""
    from timeit import default_timer as timer
    fields_arr, histories = [], {}
    for n in range(param.nrun):
        t0 = timer()
        f = param.initializer() if x is None else x
        f = f.to(torch.float64)
        q0 = float(topo_charge(f.unsqueeze(0))[0])
        ps, us = [], []
        for _ in range(param.ntraj):
            ps.append(torch.randn_like(f))
            us.append(torch.rand([], dtype=torch.float64))
        r = hmc_run_batch(param, f.unsqueeze(0), param.ntraj, torch.stack(ps).unsqueeze(1), torch.stack(us).reshape(-1, 1))
        dt = (timer() - t0) / param.ntraj
        qs = [float(v) for v in r["topo"][:, 0]]
        prev = [q0] + qs[:-1]
        histories[n] = {"traj": [n * param.ntraj + i + 1 for i in range(param.ntraj)], "dt": [dt] * param.ntraj,
                        "acc": [float(v) for v in r["acc"][:, 0]], "dH": [float(v) for v in r["dH"][:, 0]],
                        "plaq": [float(v) for v in r["plaq"][:, 0]], "q": [int(v) for v in qs],
                        "dq": [abs(a - b) for a, b in zip(qs, prev)]}
        fields_arr.append([r["field"][0]])
        if out is not None:
            for line in _status_lines(n * param.ntraj, r):
                out.write(line)
    return fields_arr, histories


# ------------------------------------------------------------------------------------------------
# flow training: the gradient of the reverse-KL loss with respect to the CNN weights
# ------------------------------------------------------------------------------------------------
def ft_action_grad(param, flow, x, want_force=False):
    """ft_action(x_b) for every chain and d/d(weights) of sum_b ft_action(x_b), the quantity the reference's reverse-KL
    train_step back-propagates (ipynb/ft_hmc.py:253-295), in ONE launch: (action (B,), grad (n_layers, 955) float64 CPU
    tensor in the parameter order of layer.plaq_coupling.net [, force (B,2,L0,L1)])."""
    import numpy as np
    dev = _device(x)
    with torch.cuda.device(dev):
        pf = pack(flow, device=dev)
        xd = _dev_in(x, dev, torch.float64)
        if xd.dim() != 4:
            raise _lib.FthmcError(-1, f"field must be (B,2,L0,L1), got {tuple(xd.shape)}")
        B, _, L0, L1 = xd.shape
        L = _lib.lib()
        ws = _workspace(pf.handle, B, L0, L1, dev, need=L.fthmc_grad_workspace_bytes(pf.handle, B, L0, L1), kind="grad")
        act = torch.empty(B, dtype=torch.float64, device=dev)
        gd = L.fthmc_grad_doubles()
        gc = torch.empty((pf.n_layers, gd), dtype=torch.float64, device=dev)
        frc = torch.empty_like(xd) if want_force else None
        _lib.check(L.fthmc_ft_action_grad(pf.handle, xd.data_ptr(), float(param.beta), act.data_ptr(), gc.data_ptr(), _ptr(frc),
                                          B, L0, L1, ws.data_ptr(), ws.numel(), _stream()))
        gch = np.ascontiguousarray(gc.cpu().numpy())
    raw = np.zeros((pf.n_layers, 955), dtype=np.float64)
    _lib.check(L.fthmc_grad_unpack(gch.ctypes.data, pf.n_layers, pf.mu.ctypes.data, raw.ctypes.data))
    out = (_back(act, x), torch.from_numpy(raw))
    return out + (_back(frc, x),) if want_force else out


# 6th-order central difference: f'(0) = [3/4 (f1 - f-1) - 3/20 (f2 - f-2) + 1/60 (f3 - f-3)] / h + O(h^6)
_FD6 = ((1, 3.0 / 4.0), (2, -3.0 / 20.0), (3, 1.0 / 60.0))


def ft_force_norm_grad(param, flow, xi, step=1e-3):
    """The second training loss of the reference (ipynb/ft_hmc.py:266-269, used at :367): loss = sum_b |ft_force(xi_b)|^2, and
    its gradient with respect to the CNN weights, which the reference gets from `ft_force(..., create_graph=True)` +
    `loss.backward()`.  With F = d S_FT / d xi and xi held fixed,

        d/dw sum_i F_i^2 = 2 sum_i F_i d^2 S_FT / (d xi_i dw) = d/d eps [ dS_FT/dw (xi + eps v) ] at eps = 0,   v = 2 F,

    i.e. the directional derivative, along the field direction v, of the weight gradient that fthmc_ft_action_grad already
    computes (summed over the batch: the batch direction field (v_b) gives the batch-summed loss gradient in one go).  S_FT
    is analytic in xi (every ingredient is a smooth 2 pi-periodic function of the links), so the derivative is taken by a
    6th-order central difference of six gradient launches at xi +- k h v, max|h v| = `step` radians: truncation ~ step^6,
    rounding ~ 1e-16 / step -- about 1e-12 relative at the default (tests/test_gpu_parity.py holds it to 1e-8 against
    torch.autograd with create_graph=True on the oracle).
    Returns (loss 0-d, grad (n_layers, 955) float64 CPU in the reference's parameter order, force (B,2,L0,L1))."""
    dev = _device(xi)
    with torch.cuda.device(dev):
        pf = pack(flow, device=dev)
        xd = _dev_in(xi, dev, torch.float64)
        F = _flow_call("force", pf, xd, beta=param.beta)
        v = 2.0 * F
        vmax = float(v.abs().max())
        if not vmax > 0.0:
            return _back((F * F).sum(), xi), torch.zeros((pf.n_layers, 955), dtype=torch.float64), _back(F, xi)
        h = step / vmax
        acc = None
        for k, c in _FD6:
            _, gp = ft_action_grad(param, pf, xd + (k * h) * v)
            _, gm = ft_action_grad(param, pf, xd - (k * h) * v)
            term = c * (gp - gm)
            acc = term if acc is None else acc + term
    return _back((F * F).sum(), xi), acc / h, _back(F, xi)


# ------------------------------------------------------------------------------------------------
# the flow as a differentiable torch operation
# ------------------------------------------------------------------------------------------------
def flow_vjp(flow, x, gy, glj, want_grad_x=True):
    """Vector-Jacobian product of the flow map (x, weights) -> (y = F(x), logJ): (grad_weights (n_layers, 955) float64 CPU in
    the reference's parameter order, grad_x (B,2,L0,L1) or None) of  sum_b [<gy_b, y_b> + glj_b logJ_b]  in ONE launch."""
    import numpy as np
    dev = _device(x)
    with torch.cuda.device(dev):
        pf = pack(flow, device=dev)
        xd = _dev_in(x, dev, torch.float64)
        gyd = _dev_in(gy, dev, torch.float64)
        gld = _dev_in(glj.reshape(-1), dev, torch.float64)
        if xd.dim() != 4 or gyd.shape != xd.shape or gld.numel() != xd.shape[0]:
            raise _lib.FthmcError(-1, "x, gy must be (B,2,L0,L1) and glj (B,)")
        B, _, L0, L1 = xd.shape
        L = _lib.lib()
        ws = _workspace(pf.handle, B, L0, L1, dev, need=L.fthmc_grad_workspace_bytes(pf.handle, B, L0, L1), kind="grad")
        gc = torch.empty((pf.n_layers, L.fthmc_grad_doubles()), dtype=torch.float64, device=dev)
        gx = torch.empty_like(xd) if want_grad_x else None
        _lib.check(L.fthmc_flow_vjp(pf.handle, xd.data_ptr(), gyd.data_ptr(), gld.data_ptr(), gc.data_ptr(), _ptr(gx),
                                    B, L0, L1, ws.data_ptr(), ws.numel(), _stream()))
        gch = np.ascontiguousarray(gc.cpu().numpy())
    raw = np.zeros((pf.n_layers, 955), dtype=np.float64)
    _lib.check(L.fthmc_grad_unpack(gch.ctypes.data, pf.n_layers, pf.mu.ctypes.data, raw.ctypes.data))
    return torch.from_numpy(raw), gx


class _FlowFunction(torch.autograd.Function):
    """(raw weights (n_layers, 955), xi (B,2,L0,L1)) -> (F(xi), sum_layers logJ): forward = fthmc_flow_fwd, backward = fthmc_flow_vjp."""

    @staticmethod
    def forward(ctx, raw, xi, activation, convention):
        pf = PackedFlow(raw.detach().double().cpu().numpy(), activation=activation, convention=convention, device=_device(xi))
        y, lj = _flow_call("fwd", pf, xi.detach(), want_logJ=True)
        ctx.pf, ctx.raw_meta = pf, (raw.device, raw.dtype)
        ctx.save_for_backward(xi.detach())
        return y, lj

    @staticmethod
    def backward(ctx, gy, glj):
        (xi,) = ctx.saved_tensors
        gy = torch.zeros_like(xi) if gy is None else gy
        glj = torch.zeros(xi.shape[0], dtype=xi.dtype, device=xi.device) if glj is None else glj
        graw, gx = flow_vjp(ctx.pf, xi, gy.contiguous(), glj.contiguous(), want_grad_x=ctx.needs_input_grad[1])
        dev, dt = ctx.raw_meta
        return (graw.to(device=dev, dtype=dt) if ctx.needs_input_grad[0] else None), (None if gx is None else _back(gx, xi)), None, None


def differentiable_flow(flow, xi, activation="silu", convention=0):
    """x, logJ = differentiable_flow(flow, xi): the forward flow as ONE differentiable torch operation, so that the reference's
    training code -- `apply_flow_to_prior` + any loss + `loss.backward()` (ipynb/ft_hmc.py:253-295,
    ipynb/field_transformation.py:107-115) -- runs on the kernels unchanged: forward is one launch of fthmc_flow_fwd, backward
    one launch of fthmc_flow_vjp.  `flow`: a reference-style ModuleList (gradients reach its Conv2d parameters) or a
    (n_layers, 955) tensor of raw weights in the reference's parameter order.  First-order only (no double backward)."""
    if isinstance(flow, torch.Tensor):
        raw = flow
    else:
        rows = []
        for layer in flow:
            convs = [m for m in layer.plaq_coupling.net if hasattr(m, "weight")]
            rows.append(torch.cat([t.reshape(-1) for c in convs for t in (c.weight, c.bias)]))
        raw = torch.stack(rows)
        from .flow import _activation_of
        activation = _activation_of(flow[0].plaq_coupling.net)
    return _FlowFunction.apply(raw, xi, activation, convention)


class _ActionFunction(torch.autograd.Function):
    """U1GaugeAction(beta)(cfgs) with its gradient from the force stencil."""

    @staticmethod
    def forward(ctx, cfgs, beta):
        ctx.beta = float(beta)
        ctx.save_for_backward(cfgs.detach())
        return _reduce_call("action", cfgs.detach(), 0, beta)

    @staticmethod
    def backward(ctx, gs):
        (cfgs,) = ctx.saved_tensors

        class _P:
            beta = ctx.beta
        f = force(_P, cfgs, order=0)
        return f * gs.to(f.device).reshape(-1, 1, 1, 1), None


def differentiable_u1_action(beta, cfgs):
    """U1GaugeAction(beta)(cfgs) (ipynb/field_transformation.py:120-137) as a differentiable torch operation: forward the action
    stencil, backward the force stencil scaled by the upstream gradient."""
    return _ActionFunction.apply(cfgs, beta)
