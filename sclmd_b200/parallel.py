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
    """where collective payloads live: the GPU of this rank for NCCL (LOCAL_RANK under torchrun -- NOT the CUDA runtime's current
    device, which the C library switches when a handle on another device is used), the host for gloo"""
    import os
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        if "LOCAL_RANK" in os.environ:
            return torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def rank_world():
    """(rank, world size) of the torch.distributed job this process belongs to; (0, 1) outside one"""
    dist = _dist()
    if dist is None:
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def local_device(default=0):
    """the GPU of this rank: LOCAL_RANK under torchrun, else `default`"""
    import os
    return int(os.environ.get("LOCAL_RANK", default))


def broadcast_array(a, src=0):
    """every rank gets rank `src`'s array (same shape and dtype on all ranks); identity when not distributed"""
    import torch
    dist = _dist()
    a = np.ascontiguousarray(a)
    if dist is None or dist.get_world_size() == 1:
        return a.copy()
    if np.iscomplexobj(a):
        return broadcast_array(a.view(np.float64), src).view(np.complex128)
    t = torch.from_numpy(a.copy()).to(_device_for_backend())
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def broadcast_int(v, src=0):
    """rank `src`'s integer on every rank (random seeds: one stream for the whole ensemble)"""
    return int(broadcast_array(np.array([int(v)], dtype=np.int64), src)[0])


def gather_rows(local_rows, n_total):
    """all-gather of contiguous row blocks produced with shard_range: local [n_local, ...] -> [n_total, ...] on every rank
    (real or complex; the trailing shape is the same everywhere)"""
    v = np.ascontiguousarray(local_rows)
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return v.copy()
    if np.iscomplexobj(v):
        return gather_rows(v.view(np.float64), n_total).view(np.complex128)
    trail = v.shape[1:]
    width = int(np.prod(trail)) if trail else 1
    flat = gather_blocks_2d(v.reshape(len(v), width).astype(np.float64), n_total, width)
    return flat.reshape((n_total,) + trail)


def gather_blocks_2d(v, n_total, width):
    import torch
    dist = _dist()
    world = dist.get_world_size()
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    dev = _device_for_backend()
    pad = torch.zeros((mx, width), dtype=torch.float64, device=dev)
    if len(v):
        pad[:len(v)] = torch.from_numpy(v).to(dev)
    bufs = [torch.zeros((mx, width), dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return np.concatenate([bufs[r][:hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)], axis=0)


def sharded_sweep(fn, omegas):
    """frequency sweep split into contiguous blocks over the ranks (negf.py:114-115 / selfenergy.py:156-160: every frequency is
    independent), all-gathered: every rank returns the full result.  fn(omegas_block) -> [n_block, ...]"""
    om = np.asarray(omegas, dtype=float)
    rank, world = rank_world()
    if world == 1:
        return np.asarray(fn(om))
    lo, hi = shard_range(len(om), rank, world)
    part = np.asarray(fn(om[lo:hi])) if hi > lo else None
    if part is None:                        # more ranks than frequencies: an empty block of the right trailing shape
        trail = np.asarray(fn(om[:1])).shape[1:]
        part = np.zeros((0,) + trail, dtype=np.asarray(fn(om[:1])).dtype)
    return gather_rows(part, len(om))


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
