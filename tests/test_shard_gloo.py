"""The N > 1 host logic on CPU: world_size-2 gloo processes exercise the row sharding, the gather of per-rank
m2lnp blocks and the rank-summed normal equations (the exchange the NCCL path does inside the C ABI), with the
CPU oracle standing in for the per-rank device evaluation."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    from helpers import make_sd, mvnd_problem
    from numcosmo_b200 import shard
    from oracle import ncm_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d, n = 4, 203   # odd on purpose: ragged shards
        mu, cov, X, m2lnL = mvnd_problem(O, d, n, seed=11)
        sd = make_sd(O, O.SD_VKDE, O.KERNEL_GAUSS, 3.0, X, m2lnp=m2lnL)
        Q = np.vstack([X[:77] + 0.01, mu + 2.0 * (X[77:150] - mu)])
        full = sd.eval_m2lnp_batch(Q, 1)
        # 1. sharded evaluation + gather == unsharded
        got = shard.ShardedEval(lambda x: sd.eval_m2lnp_batch(x, 1))(Q)
        ok_eval = np.array_equal(got, full)
        # 2. row-sharded normal equations summed over ranks == A^T A, A^T f of the whole IM
        IM = sd.compute_IM()
        r0, r1 = shard.row_range(n, rank, world)
        M = shard.allreduce_sum(IM[r0:r1].T @ IM[r0:r1])
        b = shard.allreduce_sum(IM[r0:r1].T @ np.ones(r1 - r0))
        ok_M = np.allclose(M, IM.T @ IM, rtol=1e-13, atol=0) and np.allclose(b, IM.T @ np.ones(n), rtol=1e-13, atol=0)
        # 3. the NNLS on the summed system is the NNLS of the unsharded one (same passive set, same weights)
        x_full, _, st_full = O.nnls_solve(IM, np.ones(n))
        ok_rows = shard.row_range(n, 0, world)[0] == 0 and shard.row_range(n, world - 1, world)[1] == n
        # 4. timing reduction
        t = shard.max_over_ranks(float(rank + 1))
        # 5. empty shard (more ranks than rows is legal for the partition function)
        e0, e1 = shard.row_range(1, rank, world)
        ok_empty = (e1 - e0) in (0, 1)
        q.put((rank, ok_eval, ok_M, ok_rows, t == float(world), ok_empty, float(np.abs(x_full).sum()) > 0))
    finally:
        dist.destroy_process_group()


def test_row_range_partition():
    sys.path.insert(0, ROOT)
    from numcosmo_b200 import shard

    for n in (0, 1, 7, 2048, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            edges = [shard.row_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.row_range(10, 2, 2)


def test_spmd_partitions_of_the_c_abi():
    """The two partitions of the multi-rank (auto-shard) mode of the C ABI, restated: capi.cu eval_host / compute_IM take the row block
    [n r / G, n (r + 1) / G) and all-gather blocks padded to cap = ceil (n / G); ncm_sd_gpu_vkde_prepare takes the centre block
    [min (n, r cap), min (n, (r + 1) cap)) and all-gathers IN PLACE at offset r cap.  Both must tile 0..n exactly, fit their padded
    slots, and -- for the in-place gather -- put every non-empty block at its global position."""
    sys.path.insert(0, ROOT)
    from numcosmo_b200 import shard

    for n in (1, 7, 8, 63, 64, 65, 2048, 16384, 65537):
        for G in (1, 2, 3, 4, 8):
            cap = (n + G - 1) // G
            # query / IM rows
            gath = np.full(G * cap, np.nan)
            for r in range(G):
                a, b = (n * r) // G, (n * (r + 1)) // G
                assert (a, b) == shard.row_range(n, r, G) and 0 <= b - a <= cap
                gath[r * cap:r * cap + (b - a)] = np.arange(a, b)                # what rank r contributes to the padded all-gather
            out = np.concatenate([gath[r * cap:r * cap + shard.row_range(n, r, G)[1] - shard.row_range(n, r, G)[0]] for r in range(G)])
            assert np.array_equal(out, np.arange(n))
            # VKDE prepare_kernel centres: contiguous cap-blocks, gathered in place
            buf = np.full(G * cap, np.nan)
            edges = []
            for r in range(G):
                cbeg = min(n, r * cap)
                ncl = max(0, min(n, cbeg + cap) - cbeg)
                edges.append((cbeg, cbeg + ncl))
                assert ncl == 0 or cbeg == r * cap                                 # a non-empty block already sits at its gather offset
                buf[r * cap:r * cap + ncl] = np.arange(cbeg, cbeg + ncl)
            assert edges[0][0] == 0 and max(e[1] for e in edges) == n
            assert all(edges[i][1] == edges[i + 1][0] or edges[i + 1][0] == edges[i + 1][1] == n for i in range(G - 1))
            assert np.array_equal(buf[:n], np.arange(n))


def test_sharded_path_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r
