"""Row sharding of the density-estimation path over the ranks of one node.

Query points and interpolation-matrix rows are independent, so rank g owns the contiguous row block
[g n / G, (g + 1) n / G); centres, factors and weights are replicated.  The only exchanges are
  * the normal equations of the NNLS: partial A^T A, A^T f, A^T r and |r|^2 are summed over ranks
    (inside the C ABI with NCCL, ncm_sd_gpu_comm_init; `allreduce_sum` below is the same operation on host
    arrays for the gloo tests), and
  * the concatenation of the per-rank m2lnp blocks (`allgather_rows`).
Nothing here touches the device: the functions take whatever `torch.distributed` backend is initialised.
"""
from __future__ import annotations

import numpy as np


def row_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [r0, r1) of rank `rank`: contiguous, disjoint, covering 0..n, sizes differing by at most one."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"row_range: bad rank {rank} of {world}")
    return (n * rank) // world, (n * (rank + 1)) // world


def allgather_rows(local: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """Concatenate the row blocks of all ranks (block sizes as given by row_range) on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    r0, r1 = row_range(n_total, rank, world)
    local = np.ascontiguousarray(local, dtype=np.float64)
    if local.shape[0] != r1 - r0:
        raise ValueError(f"allgather_rows: rank {rank} holds {local.shape[0]} rows, expected {r1 - r0}")
    tail = local.shape[1:]
    width = int(np.prod(tail)) if tail else 1
    cap = max(row_range(n_total, g, world)[1] - row_range(n_total, g, world)[0] for g in range(world))
    buf = torch.zeros(cap * width, dtype=torch.float64)
    buf[: local.size] = torch.from_numpy(local.reshape(-1))
    if dist.get_backend(group) == "nccl":
        buf = buf.cuda()
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    parts = []
    for g in range(world):
        g0, g1 = row_range(n_total, g, world)
        parts.append(outs[g].cpu().numpy()[: (g1 - g0) * width].reshape((g1 - g0,) + tail))
    return np.concatenate(parts, axis=0)


def allreduce_sum(arr: np.ndarray, group=None) -> np.ndarray:
    """Sum a host array over ranks (the host-side statement of the ncclAllReduce inside ncm_sd_gpu_nnls_solve)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


def max_over_ranks(value: float, group=None) -> float:
    """Device-time of a step = the slowest rank."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class ShardedEval:
    """Batched eval_m2lnp with query rows sharded over ranks: every rank evaluates its block with `eval_fn`
    (the C-ABI call on its own device) and all ranks receive the full vector."""

    def __init__(self, eval_fn, group=None):
        self.eval_fn = eval_fn
        self.group = group

    def __call__(self, X: np.ndarray) -> np.ndarray:
        import torch.distributed as dist

        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        r0, r1 = row_range(X.shape[0], rank, world)
        local = self.eval_fn(np.ascontiguousarray(X[r0:r1])) if r1 > r0 else np.empty(0)
        return allgather_rows(np.asarray(local, dtype=np.float64), X.shape[0], self.group)
