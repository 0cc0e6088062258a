"""`FieldTransformation` -- the class form of the path used by the package's CLI (fthmc/ft_hmc.py:108-257), on the CUDA
entry points.  It runs HMC directly in the latent space of the flow (the reference's `flow_backward` / `flow_forward`
calls around the trajectory are commented out, fthmc/ft_hmc.py:203, 236), with the package conventions: `torch_mod` in
[-pi, pi) (fthmc/utils/layers.py:41-43) and `wrap` = remainder(x + pi, 2 pi) - pi after the leapfrog.

Two methods of the reference class are buggy and are NOT reproduced: `leapfrog` returns
`x + dt/2 v, v` instead of the integrated pair, and `calc_energy` adds `(v*v).sum()` without the factor 1/2 and over the
whole batch.  Here `leapfrog` is the integrator its body computes, and the energy is `action + 1/2 sum v^2` per chain, as in
`hmc` of the same class.
"""
import math
import time

import torch

from . import api
from .flow import pack


class FieldTransformation:
    def __init__(self, flow, config, lfconfig, convention=1):
        """flow: the `layers` ModuleList of a FlowModel (or a PackedFlow); config: needs `.beta .volume .lat .nd`;
        lfconfig: needs `.dt .tau .nstep` (fthmc/config.py: TrainConfig / lfConfig)."""
        self.flow, self.config, self.lfconfig = flow, config, lfconfig
        self.dt, self.tau, self.nstep = lfconfig.dt, lfconfig.tau, lfconfig.nstep
        self.convention = convention
        self._denom = config.beta * config.volume
        self._param = api.Param(beta=config.beta, lat=tuple(config.lat), tau=self.tau, nstep=self.nstep)
        self._param.dt = self.dt

    def _pf(self):
        return pack(self.flow, convention=self.convention)

    # ---- fthmc/ft_hmc.py:135-171 ----
    def action(self, x):
        return api.ft_action(self._param, self._pf(), x)

    def flow_forward(self, x):
        return api.ft_flow(self._pf(), x, with_logJ=True)

    def flow_backward(self, x):
        return api.ft_flow_inv(self._pf(), x, with_logJ=True)

    def force(self, x):
        return api.ft_force(self._param, self._pf(), x)

    @staticmethod
    def wrap(x):
        return torch.remainder(x + math.pi, 2 * math.pi) - math.pi

    def calc_energy(self, x, v):
        return self.action(x) + 0.5 * (v * v).flatten(start_dim=1).sum(-1)

    def leapfrog(self, x, v):
        return api.ft_leapfrog(self._param, self._pf(), x, v)

    # ---- fthmc/ft_hmc.py:190-257 ----
    def hmc(self, x, step=None):
        """One trajectory for every chain of x (B,2,L0,L1) in the latent space; per-chain accept/reject.  Momenta from
        `torch.randn_like`, uniforms from `torch.rand` (float64), like the reference."""
        if torch.cuda.is_available():
            x = x.cuda()
        t0 = time.time()
        metrics = {} if step is None else {"traj": step}
        v = torch.randn_like(x)
        h0 = self.calc_energy(x, v)
        x_, v_ = self.leapfrog(x, v)
        x_ = self.wrap(x_)
        h1 = self.calc_energy(x_, v_)
        dh = h1 - h0
        exp_mdh = torch.exp(-dh)
        acc = torch.rand(dh.shape, dtype=torch.float64, device=dh.device) < exp_mdh
        xnew = torch.where(acc[:, None, None, None], x_, x)
        metrics.update(dt=time.time() - t0, acc=acc, dh=dh, exp_mdh=exp_mdh)
        return xnew, metrics

    _batch_hmc = hmc

    def run(self, x=None, nprint=25, num_trajs=1024, out=None, **unused):
        """fthmc/ft_hmc.py:272-346 without its plots / TensorBoard: `num_trajs` latent-space trajectories of the batch x,
        after each one the physical-space metrics of the flowed field; returns the reference's `history` dict (lists with
        one entry per trajectory: traj, dt, acc, dh, exp_mdh, plaq, q, dq).  `out`: file-like for the status lines."""
        x = self.initializer() if x is None else x
        if torch.cuda.is_available():
            x = x.cuda()
        history = {}
        q = api.topo_charge(x)
        for i in range(num_trajs):
            x, metrics = self.hmc(x, step=i)
            qold = history["q"][i - 1] if "q" in history else q
            x_phys, _ = self.flow_forward(x)
            metrics.update(self.lattice_metrics(x_phys, qold))
            for key, val in metrics.items():
                history.setdefault(key, []).append(val)
            if out is not None and i % nprint == 0:
                out.write(", ".join(f"{k}: {float(torch.as_tensor(v).double().mean()):.5g}" for k, v in metrics.items()) + "\n")
        return history

    def initializer(self, rand=True):
        x = torch.zeros([self.config.nd] + list(self.config.lat), dtype=torch.float64)
        if rand:
            x = x.uniform_(0, 2 * math.pi)
        return x[None, :]

    def lattice_metrics(self, x, qold):
        """plaquette, (un-rounded, batched) topological charge and its change (fthmc/ft_hmc.py:265-270)."""
        q = api.topo_charge(x)
        p = -api.u1_action(self.config.beta, x) / self._denom
        return {"plaq": p, "q": q, "dq": torch.sqrt((q - qold) ** 2)}
