"""Flow-based independence-Metropolis sampler (ipynb/field_transformation.py:23-83, 107-115; fthmc/utils/samplers.py):
proposals x = F(xi), xi ~ Uniform[0, 2pi), with logq = log prior - sum logJ from ONE launch of the forward-flow kernel per
batch (fthmc_flow_fwd) and logp = -S from the action stencil; the accept/reject chain itself is the reference's
sequential host loop (it touches two scalars per proposal)."""
import math

import torch

from .api import ft_flow, u1_action


def apply_flow_to_prior(flow, lattice, batch_size, xi=None, generator=None, device=None):
    """apply_flow_to_prior(prior, coupling_layers, batch_size=..., xi=None) for the reference's uniform prior
    MultivariateUniform(0, 2pi) (ipynb/ft_hmc.py:304): returns (xi, x, logq)."""
    if xi is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        xi = torch.rand((batch_size, 2) + tuple(lattice), dtype=torch.float64, generator=generator,
                        device=dev if generator is None or generator.device.type == "cuda" else "cpu") * (2 * math.pi)
    x, logJ = ft_flow(flow, xi, with_logJ=True)
    logq = -2 * lattice[0] * lattice[1] * math.log(2 * math.pi) - logJ
    return xi, x, logq


def compute_ess(logp, logq):
    """ipynb/field_transformation.py:23-27"""
    logw = logp - logq
    log_ess = 2 * torch.logsumexp(logw, dim=0) - torch.logsumexp(2 * logw, dim=0)
    return torch.exp(log_ess) / len(logw)


def serial_sample_generator(flow, beta, lattice, batch_size, n_samples, generator=None):
    """ipynb/field_transformation.py:37-47: proposals one at a time, a fresh flowed batch whenever one runs out."""
    x = logq = logp = None
    for i in range(n_samples):
        bi = i % batch_size
        if bi == 0:
            _, x, logq = apply_flow_to_prior(flow, lattice, batch_size, generator=generator)
            logp = -u1_action(beta, x)
            x, logq, logp = x.cpu(), logq.cpu(), logp.cpu()
        yield x[bi], logq[bi], logp[bi]


def make_mcmc_ensemble(flow, beta, lattice, batch_size, n_samples, generator=None):
    """make_mcmc_ensemble(model, action, batch_size, N_samples) (ipynb/field_transformation.py:48-83): the independence
    Metropolis chain over flow proposals.  Uniform draws come from torch.rand(1) like the reference's."""
    history = {"x": [], "logq": [], "logp": [], "accepted": []}
    for new_x, new_logq, new_logp in serial_sample_generator(flow, beta, lattice, batch_size, n_samples, generator):
        if len(history["logp"]) == 0:
            accepted = True
        else:
            last_logp, last_logq = history["logp"][-1], history["logq"][-1]
            p_accept = min(1, torch.exp((new_logp - new_logq) - (last_logp - last_logq)))
            if torch.rand(1) < p_accept:
                accepted = True
            else:
                accepted = False
                new_x, new_logp, new_logq = history["x"][-1], last_logp, last_logq
        history["logp"].append(new_logp)
        history["logq"].append(new_logq)
        history["x"].append(new_x)
        history["accepted"].append(accepted)
    return history
