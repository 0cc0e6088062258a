"""Packing of a reference flow (`nn.ModuleList` of `GaugeEquivCouplingLayer`,
ipynb/field_transformation.py:339-356) into the device-resident handle the kernels read."""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib

ACTIVATIONS = {"silu": 0, "swish": 0, "leaky_relu": 1, "relu": 2}
RAW_PER_LAYER = 955


def _activation_of(net):
    for m in net:
        name = type(m).__name__.lower()
        if name == "silu":
            return "silu"
        if name == "leakyrelu":
            return "leaky_relu"
        if name == "relu":
            return "relu"
    return "silu"


class PackedFlow:
    """Immutable device copy of a flow's CNN weights + mask parameters (fthmc_flow_pack)."""

    def __init__(self, raw, mu=None, off=None, activation="silu", convention=0, inv_prec=1e-6, inv_max_iter=1000,
                 hidden=(8, 8), n_mix=2, ksize=3, device=None):
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        if raw.ndim != 2 or raw.shape[1] != RAW_PER_LAYER:
            raise _lib.FthmcError(-5, f"expected (n_layers,{RAW_PER_LAYER}) raw weights, got {raw.shape}")
        n = raw.shape[0]
        self.n_layers = n
        self.mu = np.ascontiguousarray([i % 2 for i in range(n)] if mu is None else mu, dtype=np.int32)
        self.off = np.ascontiguousarray([(i // 2) % 4 for i in range(n)] if off is None else off, dtype=np.int32)
        self.activation, self.convention = activation, int(convention)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fthmc_flow_pack(
                raw.ctypes.data, n, self.mu.ctypes.data, self.off.ctypes.data, int(hidden[0]), int(hidden[1]),
                int(n_mix), int(ksize), ACTIVATIONS[activation], self.convention, float(inv_prec), int(inv_max_iter),
                ctypes.byref(h)))
        self.handle = h
        self._fin = weakref.finalize(self, _lib.lib().fthmc_flow_free, h)

    def __len__(self):
        return self.n_layers

    def update(self, raw):
        """Replace the CNN weights in place (same layer count and masks): the per-step re-pack of a training loop."""
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        if raw.shape != (self.n_layers, RAW_PER_LAYER):
            raise _lib.FthmcError(-5, f"expected ({self.n_layers},{RAW_PER_LAYER}) raw weights, got {raw.shape}")
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            _lib.check(_lib.lib().fthmc_flow_update(self.handle, raw.ctypes.data))
        return self


def raw_weights_of(flow_module):
    """(n_layers,955) float64 in the reference's parameter order of layer.plaq_coupling.net."""
    rows = []
    for layer in flow_module:
        convs = [m for m in layer.plaq_coupling.net if hasattr(m, "weight")]
        shapes = [tuple(c.weight.shape) for c in convs]
        if shapes != [(8, 2, 3, 3), (8, 8, 3, 3), (3, 8, 3, 3)]:
            raise _lib.FthmcError(-5, f"unsupported CNN shape {shapes}: built for hidden_sizes=[8,8], "
                                      "n_mixture_comps=2, kernel_size=3")
        rows.append(np.concatenate([np.concatenate([c.weight.detach().double().cpu().numpy().ravel(),
                                                    c.bias.detach().double().cpu().numpy().ravel()]) for c in convs]))
    return np.stack(rows)


_packed = weakref.WeakKeyDictionary()           # module -> {(device, convention): (digest, PackedFlow)}; dies with the module


def pack(flow, convention=0, device=None):
    """PackedFlow for `flow`: a PackedFlow (returned as is) or a reference-style ModuleList.  The packed copy is held in a
    WeakKeyDictionary keyed by the module (so it lives and dies with it) and is validated by a digest of the raw weights on every call: in-place edits
    through `.data` -- which do not bump a tensor's version counter, and which the reference's own set_weights uses -- are
    seen, and a recycled id() can never alias another flow."""
    if isinstance(flow, PackedFlow):
        return flow
    import hashlib
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    raw = raw_weights_of(flow)
    digest = hashlib.blake2b(raw.tobytes(), digest_size=16).digest()
    try:
        slot = _packed.setdefault(flow, {})
    except TypeError:                          # not weak-referenceable / hashable: pack without caching
        slot = {}
    hit = slot.get((dev.index, convention))
    if hit is not None and hit[0] == digest:
        return hit[1]
    layers = list(flow)
    mu, off = [], []
    for i, l in enumerate(layers):
        m, o = i % 2, (i // 2) % 4            # make_u1_equiv_layers, ipynb/field_transformation.py:343-344
        am = getattr(l, "active_mask", None)  # if the layer carries its link mask, read (mu, off) from it
        if am is not None:
            am = am.detach().cpu()
            m = 0 if bool(am[0].any()) else 1
            line = am[0][0, :] if m == 0 else am[1][:, 0]
            o = int(torch.nonzero(line)[0])
        mu.append(m)
        off.append(o)
    pc = layers[0].plaq_coupling
    pf = PackedFlow(raw, mu=mu, off=off, activation=_activation_of(pc.net), convention=convention,
                    inv_prec=getattr(pc, "inv_prec", 1e-6), inv_max_iter=getattr(pc, "inv_max_iter", 1000), device=dev)
    slot[(dev.index, convention)] = (digest, pf)
    return pf


def default_init_raw(n_layers=24, seed=3647):
    """Raw weights of the reference's random-init flow: `torch.manual_seed(seed)` followed by
    `make_u1_equiv_layers(n_layers, n_mixture_comps=2, hidden_sizes=[8,8], kernel_size=3)` in fp64
    (ipynb/ft_hmc.py:519, 310-315).  The reference's `set_weights(layers)` is a no-op on a ModuleList,
    so "random init" is PyTorch's default Conv2d init, drawn in construction order; this re-draws the
    same stream (checked bit-for-bit against the reference in tests/test_host_logic.py)."""
    st = torch.get_rng_state()
    old = torch.get_default_dtype()
    try:
        torch.set_default_dtype(torch.float64)
        torch.manual_seed(seed)
        rows = []
        for _ in range(n_layers):
            convs = [torch.nn.Conv2d(ci, co, 3, padding=1, stride=1, padding_mode="circular")
                     for ci, co in ((2, 8), (8, 8), (8, 3))]
            rows.append(np.concatenate([np.concatenate([c.weight.detach().numpy().ravel(), c.bias.detach().numpy().ravel()])
                                        for c in convs]))
    finally:
        torch.set_default_dtype(old)
        torch.set_rng_state(st)
    return np.stack(rows)


# ------------------------------------------------------------------------------------------------
# on-disk formats and lattice transfer
# ------------------------------------------------------------------------------------------------
def raw_from_state_dict(sd):
    """(n_layers,955) raw weights from the `state_dict()` of a reference flow (`ModuleList` of `GaugeEquivCouplingLayer`):
    keys `"{i}.plaq_coupling.net.{k}.weight|bias"` with k the positions of the three convolutions in the `Sequential`
    (0, 2, 4 for make_conv_net, ipynb/field_transformation.py:84-99).  Also accepts a whole checkpoint dict as written by
    `save_checkpoint` (fthmc/utils/io.py:148-170: the flow is under 'model_state_dict').  Mask tensors and other
    entries are ignored: masks are regenerated from (mu, off) for whatever lattice the flow is applied to."""
    import re
    if "model_state_dict" in sd:
        sd = sd["model_state_dict"]
    pat = re.compile(r"^(?:layers\.)?(\d+)\.plaq_coupling\.net\.(\d+)\.(weight|bias)$")
    layers = {}
    for key, val in sd.items():
        m = pat.match(key)
        if m:
            layers.setdefault(int(m.group(1)), {}).setdefault(int(m.group(2)), {})[m.group(3)] = val
    if not layers:
        raise _lib.FthmcError(-5, "no '<i>.plaq_coupling.net.<k>.weight' entries: not a flow state_dict")
    rows = []
    for i in range(len(layers)):
        if i not in layers:
            raise _lib.FthmcError(-5, f"layer {i} missing from the state_dict")
        convs = [layers[i][k] for k in sorted(layers[i])]
        shapes = [tuple(c["weight"].shape) for c in convs]
        if shapes != [(8, 2, 3, 3), (8, 8, 3, 3), (3, 8, 3, 3)]:
            raise _lib.FthmcError(-5, f"unsupported CNN shape {shapes} in layer {i}")
        rows.append(np.concatenate([np.concatenate([torch.as_tensor(c["weight"]).detach().double().cpu().numpy().ravel(),
                                                    torch.as_tensor(c["bias"]).detach().double().cpu().numpy().ravel()]) for c in convs]))
    return np.stack(rows)


def pack_state_dict(sd, activation="silu", convention=0, inv_prec=1e-6, inv_max_iter=1000, device=None):
    """PackedFlow from a flow `state_dict` / checkpoint dict (see raw_from_state_dict)."""
    return PackedFlow(raw_from_state_dict(sd), activation=activation, convention=convention, inv_prec=inv_prec,
                      inv_max_iter=inv_max_iter, device=device)


def load_flow(path, activation="silu", convention=0, device=None):
    """Flow weights from disk: a `ckpt-era*-epoch*.tar` checkpoint or a bare `state_dict` file (fthmc/utils/io.py:114-197),
    or a pickled `ModuleList` as written by `torch.save(flow, 'flow_b{beta}_l{L}x{L}.dat')` (ipynb/ft_hmc.py:356-373;
    unpickling that one needs the reference's classes importable)."""
    obj = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(obj, dict):
        return pack_state_dict(obj, activation=activation, convention=convention, device=device)
    return pack(obj, convention=convention, device=device)


def flow_resize(flow, lat_new=None):
    """flow_resize(flow, lat_new) (ipynb/ft_hmc.py:511-513) / transfer_to_new_lattice (fthmc/train.py:434-455): the
    reference rebuilds the layers around the same CNNs with masks for the new lattice.  The packed flow holds only
    the translation-equivariant CNN weights and the (mu, off) of each mask, so the SAME handle already applies to every
    lattice size (multiples of 4): this returns the packed flow unchanged."""
    return pack(flow)
