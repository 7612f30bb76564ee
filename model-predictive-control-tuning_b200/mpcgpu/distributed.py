"""Population sharding across the GPUs of one box.

Candidates are independent, so the path has no data-path collective: every rank evaluates its shard and
ONE all-gather of fitness per generation puts the full cost array on every rank (what the serial optimiser
above needs).  The shard assignment is work-balanced: candidates are sorted by an a-priori work estimate
and dealt round-robin, with the inverse permutation applied after the gather.
"""
from __future__ import annotations

import numpy as np


def work_estimate(N, Nu, delta, lam, dead_max=None, soft=False):
    """Relative cost of a candidate: moves, how hard it pushes on the MV limits, and -- when the plant's longest dead
    time `dead_max` is given -- the bonus for prediction horizons that barely clear it (those tunings limit-cycle: a
    QP at every sample).  soft=True: plants with soft output bands (moves per prediction row, smallest move weight).
    The same key the library uses to order launches and to deal shards (mpc_work_key in csrc/mpcgpu.cu, mpcgpu_work_estimate)."""
    delta = np.abs(np.asarray(delta, float)); lam = np.abs(np.asarray(lam, float))
    if soft:
        return 0.15 * np.asarray(Nu, float) + 1.15 * np.asarray(Nu, float) / np.asarray(N, float) - 0.09 * np.log10(lam.min(axis=1) + 1e-300)
    w = np.log10(delta.max(axis=1) / (lam.min(axis=1) + 1e-300) + 1e-300) + 0.15 * np.asarray(Nu, float)
    if dead_max is not None:
        w = w + 2.0 * (np.asarray(N) <= int(dead_max) + 2)
    return w


def shard_indices(n: int, world: int, rank: int, work=None) -> np.ndarray:
    """Indices of the candidates rank `rank` evaluates.  With a work estimate: sorted round-robin."""
    order = np.arange(n) if work is None else np.argsort(-np.asarray(work), kind="stable")
    return order[rank::world]


def evaluate_sharded(evaluate, N, Nu, delta, lam, mode="gam", group=None, device=None, dead_max=None, soft=False):
    """evaluate(N, Nu, delta, lam, mode) -> cost array for a shard (an `Evaluator.eval_batch` wrapper).
    Returns the full-population cost on every rank (n x ny for 'gam', n for 'vns').
    `group`: torch.distributed process group (None: default group; not initialised: single rank)."""
    import torch
    import torch.distributed as dist
    N = np.asarray(N); Nu = np.asarray(Nu); delta = np.asarray(delta, float); lam = np.asarray(lam, float)
    n = len(N)
    if not (dist.is_available() and dist.is_initialized()):
        return np.asarray(evaluate(N, Nu, delta, lam, mode))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    work = work_estimate(N, Nu, delta, lam, dead_max, soft)
    if device is None and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())
    mine = shard_indices(n, world, rank, work)
    local = np.asarray(evaluate(N[mine], Nu[mine], delta[mine], lam[mine], mode), dtype=np.float64)
    width = delta.shape[1] if mode == "gam" else 1   # GAM: one cost per output (ny may be 1); VNS: a scalar
    per = (n + world - 1) // world                     # equal-size slabs for all_gather_into_tensor
    slab = torch.full((per, width), float("nan"), dtype=torch.float64, device=device)
    slab[: len(mine)] = torch.as_tensor(local.reshape(len(mine), width), device=device)
    gathered = torch.empty((world * per, width), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(gathered, slab, group=group)
    g = gathered.cpu().numpy().reshape(world, per, width)
    out = np.empty((n, width))
    for r in range(world):
        idx = shard_indices(n, world, r, work)
        out[idx] = g[r, : len(idx)]
    return out if mode == "gam" else out[:, 0]
