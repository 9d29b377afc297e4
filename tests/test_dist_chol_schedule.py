"""The N > 1 schedule of csrc/dist_chol.cu on CPU: the distributed Cholesky solve of the NNLS passive-set systems (what replaces
ncm_matrix_cholesky_solve, ncm_matrix.c:1199-1210, when |P| >= 8192 and the context carries a communicator), restated in numpy with
the SAME index arithmetic -- 1-D block-cyclic block columns (owner k mod G), the staging layout of the panel all-gather
(cap / first / nown, slot (g, m) <-> block column first (g) + m G, padding slots past the last block), the two tile lists of a
trailing update (block row k + 1 first, the rest after; only tiles whose block column this rank owns), the slot of a diagonal
inverse in the gathered array (wall (k) = (k mod G) capk + k / G), the first-failing-pivot rule -- and run

  * over real exchanges: world_size 2, gloo (broadcast of the diagonal block, all_gather of the panel rows / of the inverses);
  * in one process for G = 1, 3, 4, 8 with ragged last blocks and more ranks than block columns (a loop over simulated ranks).

Every rank must end with the complete factor, bit-identical across ranks, equal to LAPACK's upper factor to rounding; the replicated
block triangular solves with the gathered inverses must solve the system; a matrix that is not positive definite must report the
1-based index of the first failing pivot on every rank.  The block / tile sizes are parameters here (the device uses DB = 512 with
128-tiles): the arithmetic under test does not depend on them."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chol_upper_info(A):
    """dpotrf 'U' on a copy: (U, info), info = 1-based failing pivot or 0 (column-by-column, the order of the failing index is LAPACK's)."""
    n = A.shape[0]
    U = np.triu(A).copy()
    for j in range(n):
        s = U[j, j] - U[:j, j] @ U[:j, j]
        if not s > 0.0:
            return U, j + 1
        U[j, j] = np.sqrt(s)
        if j + 1 < n:
            U[j, j + 1:] = (U[j, j + 1:] - U[:j, j] @ U[:j, j + 1:]) / U[j, j]
    return U, 0


def _tile_lists(n, nb, DB, T, G, me):
    """dist_chol.cu:266-285: per step k the (ti, tj) T-tiles with ti >= (DB/T)(k+1), ti <= tj, block column of tj owned by `me`; list A =
    those of block row k + 1, list B = the rest."""
    TPB, nt = DB // T, (n + T - 1) // T
    A, B = [], []
    for k in range(nb):
        a, b = [], []
        for j in range(k + 1, nb):
            if j % G != me:
                continue
            for tj in range(j * TPB, min((j + 1) * TPB, nt)):
                for ti in range((k + 1) * TPB, tj + 1):
                    (a if ti < (k + 2) * TPB else b).append((ti, tj))
        A.append(a)
        B.append(b)
    return A, B


def _first(k, g, G):
    return k + 1 + ((g - (k + 1)) % G + G) % G          # smallest j > k with j mod G == g  (dist_chol.cu:53, 313)


def _step_local(M, n, DB, T, k, me, G, nb, Ukk, cap):
    """Panel solve of this rank's block columns into its staging area [cap, DB, DB] (panel_trsm_kernel: blockIdx.z = m <-> j = first + m G)."""
    k0, bs = k * DB, min(DB, n - k * DB)
    stage = np.zeros((cap, DB, DB))
    first = _first(k, me, G)
    nown = (nb - 1 - first) // G + 1 if first < nb else 0
    assert nown <= cap
    for m in range(nown):
        j = first + m * G
        j0, wj = j * DB, min(DB, n - j * DB)
        # U_kj = U_kk^-T A_kj: forward substitution, row by row
        X = M[k0:k0 + bs, j0:j0 + wj].copy()
        for r in range(bs):
            X[r] = (X[r] - Ukk[:r, r] @ X[:r]) / Ukk[r, r]
        stage[m, :bs, :wj] = X
    return stage


def _unpack(M, n, DB, k, G, nb, cap, gathered):
    """unpack_panel_kernel: gathered[g, m] -> rows of block k, block column first (g) + m G; slots past the last block column are padding."""
    k0, bs = k * DB, min(DB, n - k * DB)
    seen = set()
    for g in range(G):
        for m in range(cap):
            j = _first(k, g, G) + m * G
            if j >= nb:
                continue
            assert j not in seen
            seen.add(j)
            j0, wj = j * DB, min(DB, n - j * DB)
            M[k0:k0 + bs, j0:j0 + wj] = gathered[g, m, :bs, :wj]
    assert seen == set(range(k + 1, nb))                 # every trailing block column arrives exactly once


def _update(M, n, DB, T, k, tiles):
    k0, bs = k * DB, min(DB, n - k * DB)
    P = M[k0:k0 + bs]
    for ti, tj in tiles:
        i0, i1, j0, j1 = ti * T, min((ti + 1) * T, n), tj * T, min((tj + 1) * T, n)
        M[i0:i1, j0:j1] -= P[:, i0:i1].T @ P[:, j0:j1]


def _solve_replicated(M, n, DB, G, Wall, capk, rhs):
    """dist_chol.cu:365-392: y_k = W_kk^T (b_k - U[0:k0, k]^T y), x_k = W_kk (y_k - U[k, k1:] x), W_kk from slot (k mod G) capk + k / G."""
    nb = (n + DB - 1) // DB
    x = rhs.copy()
    wall = lambda k: Wall[(k % G) * capk + k // G]
    for k in range(nb):
        k0, bs = k * DB, min(DB, n - k * DB)
        r = x[k0:k0 + bs] - M[:k0, k0:k0 + bs].T @ x[:k0]
        x[k0:k0 + bs] = wall(k)[:bs, :bs].T @ r
    for k in range(nb - 1, -1, -1):
        k0, bs = k * DB, min(DB, n - k * DB)
        k1 = k0 + bs
        r = x[k0:k1] - M[k0:k1, k1:] @ x[k1:]
        x[k0:k1] = wall(k)[:bs, :bs] @ r
    return x


def dist_chol_rank(M, rhs, DB, T, me, G, bcast, allgather):
    """One rank's view of dpotrf_upper_solve_dist.  `bcast (obj, owner)` and `allgather (arr) -> [G, ...]` are the two collectives."""
    n = M.shape[0]
    nb = (n + DB - 1) // DB
    capk = (nb + G - 1) // G
    listA, listB = _tile_lists(n, nb, DB, T, G, me)
    Wmine = np.zeros((capk, DB, DB))
    info_acc = 0
    for k in range(nb):
        k0, bs, owner = k * DB, min(DB, n - k * DB), k % G
        pk = None
        if me == owner:
            Ukk, info = _chol_upper_info(M[k0:k0 + bs, k0:k0 + bs])
            pk = (Ukk, 0 if info == 0 else k0 + info)                    # dc_set_info_kernel
            if info == 0:
                Wmine[k // G, :bs, :bs] = np.linalg.inv(Ukk)             # off the chain, gathered once at the end
        Ukk, info = bcast(pk, owner)
        if info_acc == 0 and info != 0:                                  # dc_acc_info_kernel: the first failing pivot is kept
            info_acc = info
        M[k0:k0 + bs, k0:k0 + bs] = np.triu(Ukk) + np.tril(M[k0:k0 + bs, k0:k0 + bs], -1)
        if k + 1 >= nb:
            break
        cap = (nb - k - 1 + G - 1) // G
        if info_acc != 0:
            Ukk = np.eye(bs)                                             # keep the collectives matched; the result is discarded
        gathered = allgather(_step_local(M, n, DB, T, k, me, G, nb, Ukk, cap))
        _unpack(M, n, DB, k, G, nb, cap, gathered)
        _update(M, n, DB, T, k, listA[k])
        _update(M, n, DB, T, k, listB[k])
    if info_acc != 0 or rhs is None:
        return info_acc, None
    Wall = allgather(Wmine).reshape(G * capk, DB, DB)
    return 0, _solve_replicated(M, n, DB, G, Wall, capk, rhs)


def _spd(n, seed):
    rs = np.random.default_rng(seed)
    B = rs.standard_normal((n + 10, n))
    return B.T @ B + 0.1 * np.eye(n), rs.standard_normal(n)


def _run_simulated(S, b, DB, T, G):
    """All ranks in one process, in lock step: every rank's matrix lives in a list, the collectives are list operations."""
    n = S.shape[0]
    nb = (n + DB - 1) // DB
    capk = (nb + G - 1) // G
    Ms = [np.triu(S).copy() for _ in range(G)]
    lists = [_tile_lists(n, nb, DB, T, G, g) for g in range(G)]
    Wm = [np.zeros((capk, DB, DB)) for _ in range(G)]
    info_acc = 0
    for k in range(nb):
        k0, bs, owner = k * DB, min(DB, n - k * DB), k % G
        Ukk, info = _chol_upper_info(Ms[owner][k0:k0 + bs, k0:k0 + bs])
        info = 0 if info == 0 else k0 + info
        if info == 0:
            Wm[owner][k // G, :bs, :bs] = np.linalg.inv(Ukk)
        if info_acc == 0 and info != 0:
            info_acc = info
        for g in range(G):
            Ms[g][k0:k0 + bs, k0:k0 + bs] = np.triu(Ukk)
        if k + 1 >= nb:
            break
        cap = (nb - k - 1 + G - 1) // G
        U_use = Ukk if info_acc == 0 else np.eye(bs)
        gathered = np.stack([_step_local(Ms[g], n, DB, T, k, g, G, nb, U_use, cap) for g in range(G)])
        for g in range(G):
            _unpack(Ms[g], n, DB, k, G, nb, cap, gathered)
            _update(Ms[g], n, DB, T, k, lists[g][0][k])
            _update(Ms[g], n, DB, T, k, lists[g][1][k])
    if info_acc != 0:
        return info_acc, Ms, None
    Wall = np.stack(Wm).reshape(G * capk, DB, DB)
    return 0, Ms, [_solve_replicated(Ms[g], n, DB, G, Wall, capk, b) for g in range(G)]


@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("n,DB,T", [(64, 16, 4), (61, 16, 4), (33, 16, 8), (100, 8, 2), (16, 16, 4), (7, 16, 4), (130, 32, 8)])
def test_block_cyclic_schedule_simulated(n, DB, T, G):
    S, b = _spd(n, seed=n + G)
    info, Ms, xs = _run_simulated(S, b, DB, T, G)
    assert info == 0
    Uref = np.linalg.cholesky(S).T
    for g in range(G):
        # the trailing updates of a block column are done by its owner only: a non-owner's copy of rows it never needed is stale BELOW
        # the panel rows it received, but everything the factor consists of -- diagonal blocks (broadcast) and panel rows (gathered) --
        # is complete on every rank
        U = np.triu(Ms[g])
        assert np.max(np.abs(U - Uref)) <= 1e-12 * np.abs(Uref).max(), (g, np.max(np.abs(U - Uref)))
        assert np.array_equal(np.triu(Ms[g]), np.triu(Ms[0]))            # bit-identical: each tile is computed once and copied
        assert np.max(np.abs(S @ xs[g] - b)) <= 1e-9 * np.abs(b).max()
        assert np.array_equal(xs[g], xs[0])


def test_every_trailing_tile_is_updated_exactly_once_by_its_owner():
    """Union over ranks of the A and B lists of step k = all upper tiles below block row k, each once; A = block row k + 1 exactly."""
    for n, DB, T, G in [(64, 16, 4, 2), (61, 16, 4, 3), (200, 32, 8, 8), (130, 32, 8, 4), (100, 8, 2, 5)]:
        nb, TPB, nt = (n + DB - 1) // DB, DB // T, (n + T - 1) // T
        per_rank = [_tile_lists(n, nb, DB, T, G, g) for g in range(G)]
        for k in range(nb):
            allA = [t for g in range(G) for t in per_rank[g][0][k]]
            allB = [t for g in range(G) for t in per_rank[g][1][k]]
            want = {(ti, tj) for tj in range((k + 1) * TPB, nt) for ti in range((k + 1) * TPB, tj + 1)}
            assert len(allA) + len(allB) == len(set(allA) | set(allB)) == len(want) and set(allA) | set(allB) == want
            assert all((k + 1) * TPB <= ti < (k + 2) * TPB for ti, _ in allA) and all(ti >= (k + 2) * TPB for ti, _ in allB)
            for g in range(G):
                assert all((tj // TPB) % G == g for _, tj in per_rank[g][0][k] + per_rank[g][1][k])


@pytest.mark.parametrize("G", [2, 3])
@pytest.mark.parametrize("bad", [0, 5, 17, 40, 60])
def test_first_failing_pivot_is_reported(G, bad):
    n, DB, T = 61, 16, 4
    S, b = _spd(n, seed=3)
    S[bad, bad] = -1.0                                                     # pivot `bad` fails (and possibly later ones: the first is kept)
    info, _, xs = _run_simulated(S, b, DB, T, G)
    assert info == _chol_upper_info(S)[1] == bad + 1 and xs is None


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import test_dist_chol_schedule as me_mod

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def bcast(obj, owner):
            box = [obj]
            dist.broadcast_object_list(box, src=owner)
            return box[0]

        def allgather(arr):
            t = torch.from_numpy(np.ascontiguousarray(arr))
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return np.stack([o.numpy() for o in outs])

        out = []
        for n, DB, T in [(61, 16, 4), (130, 32, 8), (16, 16, 4)]:
            S, b = me_mod._spd(n, seed=n)
            M = np.triu(S).copy()
            info, x = me_mod.dist_chol_rank(M, b, DB, T, rank, world, bcast, allgather)
            Uref = np.linalg.cholesky(S).T
            ok = info == 0 and np.max(np.abs(np.triu(M) - Uref)) <= 1e-12 * np.abs(Uref).max() and np.max(np.abs(S @ x - b)) <= 1e-9 * np.abs(b).max()
            # identical bits on both ranks
            peers = allgather(np.concatenate([np.triu(M).ravel(), x]))
            out.append(bool(ok) and np.array_equal(peers[0], peers[1]))
        S, b = me_mod._spd(61, seed=9)
        S[37, 37] = -2.0
        info, x = me_mod.dist_chol_rank(np.triu(S).copy(), b, 16, 4, rank, world, bcast, allgather)
        out.append(info == 38 and x is None)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_dist_chol_world2_gloo():
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
    for rank, out in res:
        assert all(out), (rank, out)
