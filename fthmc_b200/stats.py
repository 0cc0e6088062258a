"""Blocked statistics of the topological-charge history of a run (what the reference prints after `run` / `ft_run`,
ipynb/ft_hmc.py:14-56, 168-176): the mean squared change of Q over a separation of `dt` trajectories, with an error from
`n_block` block means.  Written for the (ntraj,) / (ntraj, B) arrays the run kernels return: every quantity is one numpy
reduction over a strided view, nothing loops over trajectories.

Definitions kept from the reference so that its printed numbers are reproduced (tests/test_host_logic.py holds this to the
reference's own output): the blocks are the LAST n_block * (n // n_block) samples; the quoted "error" is the block
variance divided by sqrt(n_block - 1) -- the reference's `sigma`, which is not a standard error; `standard_error=True`
gives sqrt(variance / (n_block - 1)) instead."""
import numpy as np

n_block = 16   # ipynb/ft_hmc.py:33


def block_means(v, nb=None):
    """Means of the last nb equal blocks of v along axis 0 (the reference's block layout: a history shorter than nb gives
    one block per sample).  v: (n,) or (n, ...) -> (nb', ...)."""
    v = np.asarray(v, dtype=np.float64)
    nb = n_block if nb is None else nb
    n = v.shape[0]
    size = n // nb
    if size < 1:
        size, nb = 1, n
    if nb == 0:
        return v[:0]
    tail = v[n - nb * size:]
    return tail.reshape((nb, size) + v.shape[1:]).mean(axis=1)


def block_list(l, nb=None):
    """The blocks themselves (a list of slices of `l`), for callers that want the reference's list form."""
    nb = n_block if nb is None else nb
    size = len(l) // nb
    if size < 1:
        size, nb = 1, len(l)
    first = len(l) - nb * size
    return [l[first + k * size: first + (k + 1) * size] for k in range(nb)]


def mean_and_error(blocks, standard_error=False):
    """(mean, error) of block means along axis 0; error as documented in the module header."""
    b = np.asarray(blocks, dtype=np.float64)
    var = np.mean(np.square(b - b.mean(axis=0)), axis=0)
    scale = np.sqrt(b.shape[0] - 1) if b.shape[0] > 1 else np.nan
    return b.mean(axis=0), (np.sqrt(var) / scale if standard_error else var / scale)


def change_sqr(l, lp, standard_error=False):
    """[mean, error] of (l[i] - lp[i])^2 over the common length, blocked."""
    a, b = np.asarray(l, dtype=np.float64), np.asarray(lp, dtype=np.float64)
    n = min(a.shape[0], b.shape[0])
    if n == 0:
        return []
    m, e = mean_and_error(block_means(np.square(a[:n] - b[:n])), standard_error)
    return [m, e] if np.ndim(m) else [float(m), float(e)]


def change_sqr_vs_dt(l, dt_range=10, standard_error=False):
    """[[dt, mean, error] for dt = 1..dt_range] of the squared change over dt steps."""
    q = np.asarray(l, dtype=np.float64)
    return [[dt] + change_sqr(q, q[dt:], standard_error) for dt in range(1, dt_range + 1)]


def topo_change_sqr(topo_history, dt_range=10, standard_error=False):
    """The table save_topo_change_sqr writes (ipynb/ft_hmc.py:168-176), without the file: the first third of the history is
    dropped as thermalisation."""
    q = np.asarray(topo_history, dtype=np.float64)
    return change_sqr_vs_dt(q[q.shape[0] // 3:], dt_range, standard_error)


def batched_topo_change_sqr(topo, dt=1, drop_frac=1.0 / 3.0):
    """Many-chain form for the (ntraj, B) charge array of hmc_run_batch / ft_hmc_run_batch: mean over chains and
    trajectories of (Q(t+dt) - Q(t))^2 after dropping the first third, with the standard error from the spread over chains."""
    q = np.asarray(topo, dtype=np.float64)
    q = q[int(q.shape[0] * drop_frac):]
    if q.shape[0] <= dt:
        return [float("nan"), float("nan")]
    d2 = np.square(q[dt:] - q[:-dt]).mean(axis=0)
    return [float(d2.mean()), float(d2.std(ddof=1) / np.sqrt(d2.size)) if d2.size > 1 else float("nan")]
