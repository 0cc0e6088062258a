"""ctypes access to the serial CPU build of the chain engine (tests/emul/emul.cpp).  Test
infrastructure only: lets the non-GPU test-suite exercise the exact device code (indexing, packing,
adjoint) that the CUDA library compiles."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "emul.cpp")
LIB = os.path.join(HERE, "emul", "libfthmc_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "fthmc_b200", "csrc")

MODES = dict(flow_fwd=0, flow_inv=1, ft_action=2, ft_force=3, ft_leapfrog=4, ft_hmc=5, hmc=6, leapfrog=7)
ACTS = dict(silu=0, swish=0, leaky_relu=1, relu=2)


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("chain_engine.cuh", "chain_programs.cuh", "weight_pack.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-pthread", "-ffp-contract=off", "-mfma", "-shared", "-fPIC",
                               "-o", LIB, SRC])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emul_run.restype = ctypes.c_int
        _lib.emul_run_cluster.restype = ctypes.c_int
        _lib.emul_grad.restype = ctypes.c_int
    return _lib


def _p(a, ct=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ct))


def run(mode, raw_weights, field, *, beta=1.0, dt=0.1, nstep=1, p=None, u=None, act="silu", conv=0,
        tol=1e-6, max_iter=1000, mu=None, off=None, seed=0, traj=0, nranks=0, ntraj=1):
    """field: (B,2,L0,L1) float64.  Returns a dict of outputs.  nranks >= 1: the cluster code path of the
    engine with that many ranks (host threads standing in for the CTAs of a thread-block cluster)."""
    field = np.ascontiguousarray(field, dtype=np.float64)
    B, _, L0, L1 = field.shape
    lib().emul_set_ntraj(int(ntraj))     # run-loop mode: per-trajectory outputs are (ntraj, B), p/u (ntraj, B, ...)
    raw = np.ascontiguousarray(raw_weights, dtype=np.float64) if raw_weights is not None else np.zeros((0, 955))
    n = raw.shape[0]
    mu = np.array([i % 2 for i in range(n)] if mu is None else mu, dtype=np.int32)
    off = np.array([(i // 2) % 4 for i in range(n)] if off is None else off, dtype=np.int32)
    nb = B * int(ntraj)
    out = dict(field=np.zeros_like(field), p=np.zeros_like(field), s=np.zeros(nb), layer_logJ=np.zeros((B, max(n, 1))),
               iters=np.zeros((B, max(n, 1)), dtype=np.int32), expmdH=np.zeros(nb), acc=np.zeros(nb, dtype=np.int32),
               plaq=np.zeros(nb), topo=np.zeros(nb), h0=np.zeros(nb), h1=np.zeros(nb))
    p = None if p is None else np.ascontiguousarray(p, dtype=np.float64)
    u = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    fn = lib().emul_run if nranks == 0 else (lambda *args: lib().emul_run_cluster(nranks, *args))
    rc = fn(MODES[mode], B, L0, L1, n, _p(raw), _p(mu, ctypes.c_int), _p(off, ctypes.c_int),
                        ACTS[act], conv, ctypes.c_double(tol), max_iter, ctypes.c_double(beta), ctypes.c_double(dt),
                        nstep, _p(field), _p(p), _p(u), _p(out["field"]), _p(out["p"]), _p(out["s"]),
                        _p(out["layer_logJ"]), _p(out["iters"], ctypes.c_int), _p(out["expmdH"]),
                        _p(out["acc"], ctypes.c_int), _p(out["plaq"]), _p(out["topo"]), _p(out["h0"]), _p(out["h1"]),
                        ctypes.c_ulonglong(seed), ctypes.c_ulonglong(traj))
    lib().emul_set_ntraj(1)
    assert rc == 0, rc
    return out


def grad(raw_weights, field, *, beta=1.0, act="silu", conv=0, mu=None, off=None, vjp_seed=None, vjp_wlj=None):
    """MODE_FT_GRAD on the CPU build: ft_action (B,), d/d(raw weights) of its sum (n_layers, 955), force (B,2,L0,L1).
    vjp_seed (B,2,L0,L1) + vjp_wlj (B): the vector-Jacobian mode (fthmc_flow_vjp) -- `grad` / `force` are then d/dweights
    and d/dx of sum_b [<seed_b, F(x_b)> + wlj_b logJ_b]."""
    os.environ["FT_EMUL_MMA"] = "1"
    field = np.ascontiguousarray(field, dtype=np.float64)
    B, _, L0, L1 = field.shape
    raw = np.ascontiguousarray(raw_weights, dtype=np.float64)
    n = raw.shape[0]
    mu = np.array([i % 2 for i in range(n)] if mu is None else mu, dtype=np.int32)
    off = np.array([(i // 2) % 4 for i in range(n)] if off is None else off, dtype=np.int32)
    action, g, force = np.zeros(B), np.zeros((n, 955)), np.zeros_like(field)
    if vjp_seed is not None:
        vjp_seed = np.ascontiguousarray(vjp_seed, dtype=np.float64)
        vjp_wlj = np.ascontiguousarray(vjp_wlj, dtype=np.float64)
        lib().emul_set_vjp(_p(vjp_seed), _p(vjp_wlj))
    try:
        rc = lib().emul_grad(B, L0, L1, n, _p(raw), _p(mu, ctypes.c_int), _p(off, ctypes.c_int), ACTS[act], conv,
                             ctypes.c_double(beta), _p(field), _p(action), _p(g), _p(force))
    finally:
        os.environ.pop("FT_EMUL_MMA", None)
    assert rc == 0, rc
    return dict(action=action, grad=g, force=force)
