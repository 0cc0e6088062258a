"""ctypes loader for libfthmc_b200.so (the C ABI declared in include/fthmc_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, an exception is
raised.  Build the library with `python __graft_entry__.py build` (nvcc, sm_100a)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfthmc_b200.so")

c_dp = ctypes.c_void_p
c_int, c_dbl, c_sz, c_ull, c_ll = ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_ulonglong, ctypes.c_longlong

# every symbol include/fthmc_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fthmc_version": (c_int, []),
    "fthmc_last_error_string": (ctypes.c_char_p, []),
    "fthmc_launch_count": (c_ull, []),
    "fthmc_diag_dfma_probe": (c_int, [c_dp, c_int, c_int, c_dp, ctypes.POINTER(c_dbl)]),
    "fthmc_diag_dmma_probe": (c_int, [c_dp, c_int, c_int, c_dp, ctypes.POINTER(c_dbl)]),
    "fthmc_action": (c_int, [c_dp, c_int, c_int, c_int, c_dbl, c_int, c_dp, c_int, c_dp]),
    "fthmc_force": (c_int, [c_dp, c_int, c_int, c_int, c_dbl, c_int, c_dp, c_int, c_dp]),
    "fthmc_topo_charge": (c_int, [c_dp, c_int, c_int, c_int, c_int, c_dp, c_int, c_dp]),
    "fthmc_regularize": (c_int, [c_dp, c_dp, c_ll, c_int, c_dp]),
    "fthmc_chain_ranks": (c_int, [c_int, c_int, c_int]),
    "fthmc_workspace_bytes": (c_sz, [c_dp, c_int, c_int, c_int]),
    "fthmc_leapfrog": (c_int, [c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dbl, c_dbl, c_int, c_dp, c_sz, c_dp]),
    "fthmc_hmc_traj": (c_int, [c_dp, c_dp, c_dp, c_dp, c_ull, c_ull, c_ull, c_int, c_int, c_int, c_dbl, c_dbl, c_int,
                               c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_sz, c_dp]),
    "fthmc_flow_pack": (c_int, [c_dp, c_int, c_dp, c_dp, c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_int,
                                ctypes.POINTER(c_dp)]),
    "fthmc_flow_free": (c_int, [c_dp]),
    "fthmc_flow_update": (c_int, [c_dp, c_dp]),
    "fthmc_flow_n_layers": (c_int, [c_dp]),
    "fthmc_flow_fwd": (c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_sz, c_dp]),
    "fthmc_flow_inv": (c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_sz, c_dp]),
    "fthmc_ft_action": (c_int, [c_dp, c_dp, c_dbl, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_sz, c_dp]),
    "fthmc_ft_force": (c_int, [c_dp, c_dp, c_dbl, c_dp, c_int, c_int, c_int, c_dp, c_sz, c_dp]),
    "fthmc_ft_leapfrog": (c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dbl, c_dbl, c_int, c_dp, c_sz, c_dp]),
    "fthmc_ft_hmc_traj": (c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_ull, c_ull, c_ull, c_int, c_int, c_int, c_dbl, c_dbl,
                                  c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_sz, c_dp]),
    "fthmc_hmc_run": (c_int, [c_dp, c_dp, c_dp, c_dp, c_ull, c_ull, c_ull, c_int, c_int, c_int, c_dbl, c_dbl, c_int, c_int,
                              c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_sz, c_dp]),
    "fthmc_ft_hmc_run": (c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_ull, c_ull, c_ull, c_int, c_int, c_int, c_dbl, c_dbl,
                                 c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_sz, c_dp]),
    "fthmc_grad_workspace_bytes": (c_sz, [c_dp, c_int, c_int, c_int]),
    "fthmc_ft_action_grad": (c_int, [c_dp, c_dp, c_dbl, c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_sz, c_dp]),
    "fthmc_flow_vjp": (c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_sz, c_dp]),
    "fthmc_grad_doubles": (c_int, []),
    "fthmc_grad_unpack": (c_int, [c_dp, c_int, c_dp, c_dp]),
}


class FthmcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfthmc_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: fthmc_b200 has no CPU fallback. Build the CUDA library first "
                "(`python __graft_entry__.py build`, needs nvcc with sm_100a support).")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc):
    if rc != 0:
        raise FthmcError(rc, lib().fthmc_last_error_string().decode())
