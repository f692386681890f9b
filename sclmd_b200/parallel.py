"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The path shards without any data-path collective:
  * MD       -- independent trajectories: rank r owns a contiguous block of global trajectory
                indices (the Philox counter is the GLOBAL index, so results do not depend on the
                number of ranks); ONE all-reduce of [per-bath sum of cur, count] at the end.
  * NEGF/sig -- disjoint contiguous frequency blocks; all-gather of T(w) (or all-reduce of the
                trapezoid partial sums).
"""
import numpy as np


def shard_range(n, rank, world):
    """contiguous block [lo, hi) of n items for `rank`; sizes differ by at most one"""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device_for_backend():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def allreduce_sum(values):
    """sum a small float64 vector over all ranks (identity when not distributed)"""
    import torch
    v = np.asarray(values, dtype=np.float64)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return v.copy()
    t = torch.from_numpy(v.copy()).to(_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def ensemble_mean_currents(current_sums, count, curcof=243414.0):
    """[nbaths] heat currents averaged over every trajectory of every rank and over time, in the units of
    the kappa.* files (np.mean(cur)*curcof, md.py:663).  current_sums: per-bath sum over this rank's
    trajectories and time slots; count: ntraj_local*nmd."""
    tot = allreduce_sum(list(current_sums) + [float(count)])
    return tot[:-1] / tot[-1] * curcof


def thermal_conductance(mean_currents, T, delta):
    """tools.calTC for two baths (tools.py:193): (J0 - J1)/2/(delta*T)"""
    return (mean_currents[0] - mean_currents[1]) / 2.0 / (delta * T)


def gather_blocks(local_block, n_total):
    """all-gather contiguous blocks produced with shard_range back into one [n_total] vector on every rank"""
    import torch
    dist = _dist()
    v = np.asarray(local_block, dtype=np.float64)
    if dist is None or dist.get_world_size() == 1:
        return v.copy()
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    dev = _device_for_backend()
    pad = torch.zeros(mx, dtype=torch.float64, device=dev)
    pad[:len(v)] = torch.from_numpy(v).to(dev)
    bufs = [torch.zeros(mx, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return np.concatenate([bufs[r][:hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)])


def trapezoid_partial(values, lo, n_total, h):
    """this rank's share of the trapezoid rule of negf.py:261-267 over points [lo, lo+len(values)):
    h/2 * (2*sum - first - last) with the end-point weights applied only by the ranks owning them"""
    v = np.asarray(values, dtype=np.float64)
    s = 2.0 * v.sum()
    if lo == 0 and len(v):
        s -= v[0]
    if lo + len(v) == n_total and len(v):
        s -= v[-1]
    return h / 2.0 * s
