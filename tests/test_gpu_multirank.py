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


def _worker_host_api(rank, world, port, q):
    """SPMD mode of the host mirror: every rank makes the same ncm_stats_dist_* / APES calls; the work behind them is sharded."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    import torch
    import torch.distributed as dist

    from helpers import mvnd_problem
    from numcosmo_b200 import stats_dist as S
    from oracle import ncm_oracle as O

    torch.cuda.set_device(rank)
    S.lib().ncm_b200_set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        # --- prepare_interp + batched eval, sharded vs unsharded, ragged sizes ---
        d, n = 7, 1501
        mu, cov, X, m2lnL = mvnd_problem(O, d, n, seed=41)
        Q = np.vstack([X[:333] + 0.01, mu + 2.0 * (X[333:700] - mu)])
        res = {}
        for mode in ("single", "sharded"):
            sd = S.StatsDistVKDE(S.StatsDistKernelST(d, 3.0), S.StatsDistCV.NONE)
            if mode == "sharded":
                sd.comm_init_from_torch()
            for x in X:
                sd.add_obs(x)
            sd.prepare_interp(m2lnL)
            res[mode] = (sd.peek_weights().copy(), sd.eval_m2lnp_array(Q), sd.get_rnorm(), sd.nnls_stats(), sd.get_timers()[0]["comm"])
        w1, e1, r1, s1, _ = res["single"]
        w2, e2, r2, s2, comm_ms = res["sharded"]
        same_set = bool(np.array_equal(w1 > 0.011 / n, w2 > 0.011 / n))
        err_w = float(np.max(np.abs(w1 - w2)) / w1.max())
        err_e = float(np.max(np.abs(e1 - e2) / np.abs(e1)))
        # --- APES: the sharded run accepts the very sequence of the single-GPU run, on every rank ---
        W, da, iters = 1000, 5, 4
        mu, cov, Xa, ml = mvnd_problem(O, da, W, seed=43)
        U = np.ascontiguousarray(np.linalg.cholesky(cov).T)
        lb, ub = np.full(da, -50.0), np.full(da, 50.0)
        acc = {}
        for mode in ("single", "sharded"):
            ap = S.FitESMCMCWalkerAPES(W, da, S.FitESMCMCWalkerAPESMethod.VKDE, S.FitESMCMCWalkerAPESKType.GAUSS, 1.0, True)
            ap.set_use_threads(True)
            if mode == "sharded":
                ap.comm_init_from_torch()
            th, m = Xa.copy(), ml.copy()
            a, _ = ap.run("mvnd", lb, ub, th, m, iters, S.RNG(77), target_args=(mu, U))
            acc[mode] = (a, th)
        same_acc = bool(np.array_equal(acc["single"][0], acc["sharded"][0]))
        err_th = float(np.max(np.abs(acc["single"][1] - acc["sharded"][1])))
        t = torch.from_numpy(acc["sharded"][1].copy()).cuda()
        t0 = t.clone()
        dist.broadcast(t0, src=0)
        ranks_identical = bool(torch.equal(t, t0))
        q.put((rank, same_set, err_w, err_e, abs(r1 - r2) / r1, comm_ms, same_acc, err_th, ranks_identical, float(acc["sharded"][0].mean())))
    finally:
        dist.destroy_process_group()


def test_host_api_multirank_mode_two_ranks():
    """multi-rank mode in the host API (host/stats_dist.cc, host/apes.cc): IM rows and query rows sharded, NCCL
    all-reduce / all-gather on the device, accepted sequence identical to the one-GPU run and identical on both ranks."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_host_api, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_set, err_w, err_e, err_rn, comm_ms, same_acc, err_th, ranks_identical, acc_rate in res:
        assert same_set and err_w < 1e-9 and err_e < 1e-9 and err_rn < 1e-9, (rank, same_set, err_w, err_e, err_rn)
        assert same_acc and err_th < 1e-9 and ranks_identical and acc_rate > 0.1, (rank, same_acc, err_th, ranks_identical, acc_rate)


def _worker_dist_chol(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "NCM_SD_GPU_DIST_CHOL_MIN_N": "1024"})
    import torch
    import torch.distributed as dist

    from numcosmo_b200 import capi

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        c = capi.Context(rank)
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(world, rank, uid[0])
        out = []
        for n in (1024, 1500, 2561, 5000):
            rs = np.random.default_rng(n)           # identical data on both ranks
            B = rs.standard_normal((n + 10, n))
            S = B.T @ B + 0.1 * np.eye(n)
            b = rs.standard_normal(n)
            ld = (n + 7) // 8 * 8
            dM = torch.full((n, ld), float("nan"), dtype=torch.float64, device="cuda")
            Sh = np.full((n, ld), np.nan)
            iu = np.triu_indices(n)
            Sh[iu] = S[iu]
            dM.copy_(torch.from_numpy(Sh))
            dB = torch.from_numpy(b.copy()).cuda()
            torch.cuda.synchronize()
            info = c.dposv_upper_dev(n, dM.data_ptr(), ld, dB.data_ptr())
            x = dB.cpu().numpy()
            U = np.triu(dM.cpu().numpy()[:, :n])
            xr = np.linalg.solve(S, b)
            Ur = np.linalg.cholesky(S).T
            # both ranks hold the same factor and solution, bit for bit
            t = dB.clone()
            dist.broadcast(t, src=0)
            same = bool(torch.equal(t, dB))
            out.append((n, info, float(np.max(np.abs(x - xr)) / np.max(np.abs(xr))), float(np.max(np.abs(U - Ur)) / np.max(np.abs(Ur))), same))
        # not positive definite: the 1-based pivot index reaches every rank
        n = 2048
        rs = np.random.default_rng(5)
        B = rs.standard_normal((n + 10, n))
        S = B.T @ B + 0.1 * np.eye(n)
        S[1300, 1300] = -1.0
        dM = torch.from_numpy(np.ascontiguousarray(np.triu(S))).cuda()
        dB = torch.zeros(n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        bad = c.dposv_upper_dev(n, dM.data_ptr(), n, dB.data_ptr())
        q.put((rank, out, bad))
        c.close()
    finally:
        dist.destroy_process_group()


def test_distributed_cholesky_two_ranks():
    """the passive-set Cholesky with its trailing updates distributed over the ranks (csrc/dist_chol.cu; block columns
    of 512 dealt cyclically, panel rows exchanged with ncclBroadcast / ncclAllGather): dposv against numpy, ragged orders, identical
    results on both ranks, failing pivot reported everywhere.  The threshold (8192 in production) is lowered through the environment."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_dist_chol, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, bad in res:
        for n, info, err_x, err_u, same in out:
            assert info == 0 and err_x < 1e-9 and err_u < 1e-10 and same, (rank, n, info, err_x, err_u, same)
        assert bad == 1301, (rank, bad)
