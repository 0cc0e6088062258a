"""Block statistics of the run loops (ipynb/ft_hmc.py:14-53): the change-squared of the topological charge
versus trajectory separation, with blocked errors.  Host-side numpy on the per-trajectory arrays the run kernels
return; same definitions and block layout as the reference (n_block = 16)."""
import numpy as np

n_block = 16   # ipynb/ft_hmc.py:33


def average(l):
    return sum(l) / len(l)


def sigma(l):
    """ipynb/ft_hmc.py:22-25 (note: the reference divides the variance, not its root, by sqrt(n-1))."""
    avg = average(l)
    sq_avg = average([np.square(v - avg) for v in l])
    return sq_avg / np.sqrt(len(l) - 1)


def sub_avg(l):
    avg = average(l)
    return np.array([x - avg for x in l])


def block_list(l, nb=None):
    """ipynb/ft_hmc.py:35-44: the last n_block * size_block entries, in n_block consecutive blocks."""
    n_block_local = n_block if nb is None else nb
    size_block = len(l) // n_block_local
    if size_block < 1:
        size_block = 1
        n_block_local = len(l)
    if n_block_local == 0:
        return []
    start = len(l) - n_block_local * size_block
    return [l[start + i * size_block: start + (i + 1) * size_block] for i in range(n_block_local)]


def change_sqr(l, lp):
    """ipynb/ft_hmc.py:46-53: blocked mean and error of (l[i] - lp[i])^2."""
    size = min(len(l), len(lp))
    if size == 0:
        return []
    vs = [np.square(l[i] - lp[i]) for i in range(size)]
    vs = list(map(average, block_list(vs)))
    return [average(vs), sigma(vs)]


def change_sqr_vs_dt(l, dt_range=10):
    """ipynb/ft_hmc.py:55-56."""
    return [[i] + change_sqr(l, l[i:]) for i in range(1, dt_range + 1)]


def topo_change_sqr(topo_history, dt_range=10):
    """save_topo_change_sqr (ipynb/ft_hmc.py:168-176) without the file: drops the first third of the history."""
    drop_len = len(topo_history) // 3
    return change_sqr_vs_dt(list(topo_history[drop_len:]), dt_range)


def batched_topo_change_sqr(topo, dt=1, drop_frac=1.0 / 3.0):
    """Many-chain form for the (ntraj, B) charge array of hmc_run_batch / ft_hmc_run_batch: mean over chains and
    trajectories of (Q(t+dt) - Q(t))^2 after dropping the first third, with the error from the spread over chains."""
    q = np.asarray(topo, dtype=np.float64)
    q = q[int(q.shape[0] * drop_frac):]
    if q.shape[0] <= dt:
        return [float("nan"), float("nan")]
    d2 = np.square(q[dt:] - q[:-dt]).mean(axis=0)
    return [float(d2.mean()), float(d2.std(ddof=1) / np.sqrt(d2.size)) if d2.size > 1 else float("nan")]
