"""Two ranks, two GPUs, NCCL inside the C ABI: the row-sharded prepare_interp (IM row blocks per rank,
all-reduced normal equations, replicated passive-set Cholesky) returns on every rank the weights of the unsharded solve, and the
row-sharded batched evaluation concatenates to the unsharded one.  Skipped on a single-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    import torch
    import torch.distributed as dist

    from helpers import make_sd, mvnd_problem, upload_from_oracle
    from numcosmo_b200 import capi, shard
    from oracle import ncm_oracle as O

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        d, n = 6, 1203   # odd: ragged row blocks
        mu, cov, X, m2lnL = mvnd_problem(O, d, n, seed=31)
        sd = make_sd(O, O.SD_VKDE, O.KERNEL_GAUSS, 3.0, X)
        rowscale = 1.0 / np.exp(-0.5 * (m2lnL - m2lnL.min()))
        # unsharded solve on this rank's own device
        ref = capi.Context(rank)
        upload_from_oracle(ref, capi, O, sd, O.SD_VKDE, O.KERNEL_GAUSS, 3.0, X)
        ref.compute_IM(rowscale)
        x_ref, rn_ref, st_ref = ref.nnls_solve()
        # sharded
        c = capi.Context(rank)
        href = upload_from_oracle(c, capi, O, sd, O.SD_VKDE, O.KERNEL_GAUSS, 3.0, X)
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(world, rank, uid[0])
        r0, r1 = shard.row_range(n, rank, world)
        c.set_row_shard(r0, r1 - r0)
        c.compute_IM(rowscale)
        x, rn, st = c.nnls_solve()
        same_set = np.array_equal(x > 0, x_ref > 0)
        err_w = float(np.max(np.abs(x - x_ref)) / np.max(np.abs(x_ref)))
        # weights identical on all ranks (replicated decisions on all-reduced data)
        xs = shard.allgather_rows(x[None, :], world)
        identical = bool(np.array_equal(xs[0], xs[-1]))
        # sharded evaluation
        w = (1.0 - 0.01) * x / x.sum() + 0.01 / n
        c.set_weights(w, href)
        ref.set_weights(w, href)
        Q = np.vstack([X[:301] + 0.01, mu + 2.0 * (X[301:500] - mu)])
        full = ref.eval_m2lnp(Q)
        got = shard.ShardedEval(lambda q: c.eval_m2lnp(q))(Q)
        q.put((rank, same_set, err_w, identical, bool(np.array_equal(got, full)), abs(rn - rn_ref) / rn_ref, st["n_chol"], st_ref["n_chol"]))
        c.close()
        ref.close()
    finally:
        dist.destroy_process_group()


def test_sharded_prepare_interp_and_eval_two_ranks():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_set, err_w, identical, eval_equal, err_rn, n1, n2 in res:
        assert same_set, (rank, "passive sets differ")
        assert err_w < 1e-6, (rank, err_w)          # conditioning-limited; the summation order of M differs
        assert err_rn < 1e-9, (rank, err_rn)
        assert identical and eval_equal, (rank, identical, eval_equal)
