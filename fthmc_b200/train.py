"""Flow training on top of the weight-gradient kernel (ipynb/ft_hmc.py:253-346, fthmc/train.py:162-228): the reverse-KL step
and the force-norm step of the reference's second training stage.

The reference's train_step draws a batch from the uniform prior, flows it, forms loss = mean(logq - logp) and calls
loss.backward() / optimizer.step().  With logq = log prior - sum logJ and logp = -S the loss is mean_b ft_action(xi_b) +
const, so its gradient is (1/B) d/dweights sum_b ft_action(xi_b): one launch of fthmc_ft_action_grad.  The weights live in
ONE (n_layers, 955) tensor in the reference's parameter order; the optimizer is torch's (host side, 23k parameters).
With several GPUs every rank draws its own prior batch, the gradient (and the loss terms) are all-reduced over NCCL
(shard.allreduce_gradient) and every rank takes the same optimizer step."""
import math

import numpy as np
import torch

from . import shard
from .api import ft_action_grad
from .flow import PackedFlow


class FlowTrainer:
    def __init__(self, raw_weights, lattice, beta, lr=1e-4, activation="silu", convention=0, seed=None):
        self.raw = torch.nn.Parameter(torch.as_tensor(np.asarray(raw_weights), dtype=torch.float64).clone())
        self.lattice, self.beta = tuple(lattice), float(beta)
        self.activation, self.convention = activation, convention
        self.opt = torch.optim.Adam([self.raw], lr=lr)           # base_lr = 1e-4, ipynb/ft_hmc.py:316-317
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.gen = torch.Generator(device=self.dev)
        self.gen.manual_seed(int(seed) if seed is not None else torch.seed())
        self.history = {"loss": [], "force": [], "dkl": [], "ess": []}
        self._pf, self._stale = None, True

    # ---- model pieces -------------------------------------------------------------------------------
    def packed(self):
        """device copy of the current weights (re-uploaded into the same handle after every optimizer step)"""
        if self._pf is None:
            self._pf = PackedFlow(self.raw.detach().numpy(), activation=self.activation, convention=self.convention, device=self.dev)
        elif self._stale:
            self._pf.update(self.raw.detach().numpy())
        self._stale = False
        return self._pf

    def sample_prior(self, batch_size):
        """MultivariateUniform(0, 2pi).sample_n (ipynb/ft_hmc.py:304), drawn on the device."""
        shape = (batch_size, 2) + self.lattice
        return torch.rand(shape, dtype=torch.float64, generator=self.gen, device=self.dev) * (2 * math.pi)

    def log_prior(self):
        return -2 * self.lattice[0] * self.lattice[1] * math.log(2 * math.pi)

    # ---- one optimizer step -------------------------------------------------------------------------
    def train_step(self, batch_size, xi=None, group=None, with_force=False, pre_model=None):
        """train_step(model, action, optimizer, metrics, batch_size, param, with_force, pre_model) (ipynb/ft_hmc.py:253-295).

        with_force=False: loss = dkl = mean(logq - logp) with logq = log prior - sum logJ, logp = -S, i.e. mean ft_action +
        log prior.  pre_model (a FlowTrainer or a packed / reference flow): the latent batch is not drawn from the prior but
        pulled back from the pre-trained flow's samples, xi = F^-1(F_pre(xi_pre)) (:258-262), held fixed for the step.
        with_force=True (needs pre_model, as in the reference): loss = sum_b |ft_force(xi_b)|^2 (:266-269), its weight
        gradient from ft_force_norm_grad.  Returns the metrics of this step (also appended to self.history)."""
        from .api import ft_flow, ft_flow_inv, ft_force_norm_grad

        class _P:                                   # the entry points read only beta
            beta = self.beta
        if pre_model is not None and xi is None:
            pre = pre_model.packed() if isinstance(pre_model, FlowTrainer) else pre_model
            xi = ft_flow_inv(self.packed(), ft_flow(pre, self.sample_prior(batch_size)))
        if with_force:
            assert pre_model is not None or xi is not None, "the force-norm step samples through a pre-trained flow (ipynb/ft_hmc.py:267)"
        if xi is None:
            xi = self.sample_prior(batch_size)
        xi = xi.to(self.dev)
        if with_force:
            from .api import ft_action
            fsize, grad, _ = ft_force_norm_grad(_P, self.packed(), xi)
            act = ft_action(_P, self.packed(), xi).cpu()
            fs = fsize.detach().cpu().reshape(1).double()
        else:
            act, grad = ft_action_grad(_P, self.packed(), xi)
            act = act.cpu()
            fs = torch.zeros(1, dtype=torch.float64)
        sums = torch.cat([act.sum().reshape(1), torch.tensor([float(act.numel())], dtype=torch.float64), fs])
        grad, sums = shard.allreduce_gradient(grad, sums, group=group)      # all ranks: same gradient, same step
        nb = float(sums[1])
        dkl = float(sums[0]) / nb + self.log_prior()
        force_size = float(sums[2])
        self.opt.zero_grad()
        self.raw.grad = (grad if with_force else grad / nb).to(torch.float64)      # sum over the batch / mean over the batch
        self.opt.step()
        self._stale = True                           # weights changed: re-upload lazily
        logw = -(act + self.log_prior())             # logp - logq of this rank's batch
        ess = float(torch.exp(2 * torch.logsumexp(logw, 0) - torch.logsumexp(2 * logw, 0)) / act.numel())   # compute_ess
        m = {"loss": force_size if with_force else dkl, "force": force_size, "dkl": dkl, "ess": ess}
        for k, v in m.items():
            self.history.setdefault(k, []).append(v)
        return m


# ------------------------------------------------------------------------------------------------
# the reference's training / evaluation drivers around train_step
# ------------------------------------------------------------------------------------------------
def flow_train(lattice, beta, n_layers=24, n_era=10, n_epoch=30, batch_size=64, base_lr=1e-4, with_force=False, pre_model=None,
               raw_weights=None, seed=None, out=None):
    """flow_train(param, with_force, pre_model) (ipynb/ft_hmc.py:297-346): a 24-layer flow from PyTorch's default Conv2d init
    (the reference's set_weights is a no-op on a ModuleList), Adam at base_lr for the reverse-KL step and, with
    `with_force`, a second Adam at base_lr / 100 for a force-norm step after every reverse-KL step, n_era x n_epoch steps of
    each.  Returns the FlowTrainer (its `.packed()` flow drives every entry point; `.history` has loss / force / dkl / ess)."""
    from .flow import default_init_raw
    raw = default_init_raw(n_layers, 3647 if seed is None else seed) if raw_weights is None else raw_weights
    tr = FlowTrainer(raw, lattice, beta, lr=base_lr, seed=seed)
    opt_kl, opt_wf = tr.opt, torch.optim.Adam([tr.raw], lr=base_lr / 100.0)
    for era in range(n_era):
        for epoch in range(n_epoch):
            tr.opt = opt_kl
            tr.train_step(batch_size)
            if with_force:
                assert pre_model is not None, "the force-norm stage samples through a pre-trained flow (ipynb/ft_hmc.py:267)"
                tr.opt = opt_wf
                tr.train_step(batch_size, with_force=True, pre_model=pre_model)
        if out is not None:
            n = n_epoch * (2 if with_force else 1)
            out.write(f"== Era {era} ==  " + "  ".join(f"{k} {np.mean(v[-n:]):g}" for k, v in tr.history.items()) + "\n")
    tr.opt = opt_kl
    return tr


def blocked_bootstrap(x, n_boot=100, binsize=16, rng=None):
    """(mean, error) of the mean of x by resampling whole bins of `binsize` consecutive samples with replacement (what the
    reference's bootstrap(x, Nboot=, binsize=) estimates, ipynb/field_transformation.py:28-33)."""
    x = np.asarray(x, dtype=np.float64)
    nb = x.shape[0] // binsize
    bins = x[:nb * binsize].reshape(nb, binsize).mean(axis=1)
    rng = np.random.default_rng() if rng is None else rng
    means = bins[rng.integers(nb, size=(n_boot, nb))].mean(axis=1)
    return float(means.mean()), float(means.std())


def flow_eval(flow, beta, lattice, ensemble_size=1024, batch_size=64, generator=None, rng=None):
    """flow_eval(model, action) (ipynb/ft_hmc.py:348-354): an independence-Metropolis ensemble from the flow, its accept rate
    and the topological susceptibility <Q^2> with a blocked bootstrap error."""
    from . import sampler
    from .api import topo_charge
    flow = flow.packed() if isinstance(flow, FlowTrainer) else flow
    ens = sampler.make_mcmc_ensemble(flow, beta, tuple(lattice), batch_size, ensemble_size, generator=generator)
    q = topo_charge(torch.stack(ens["x"], dim=0)).cpu().numpy()
    chi, err = blocked_bootstrap(q ** 2, n_boot=100, binsize=16, rng=rng)
    return {"accept_rate": float(np.mean(ens["accepted"])), "Q2": chi, "Q2_err": err, "ensemble": ens}
