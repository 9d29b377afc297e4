#!/usr/bin/env python
"""bench.py -- APES density-estimation hot path on B200 (see DESIGN.md, section Measurement).

Workload at N = 1 (BASELINE.json configs[1]): APES + VKDE, Gaussian kernel, 10-D multivariate normal
target (NcmDataGaussCovMVND recipe), 4096 walkers.  One "step" = one whole-ensemble APES iteration
= 2 half-steps, each: interpolation matrix (N x N pairs, N = W/2 = 2048) + NNLS weight solve +
batched eval_m2lnp of the 2N transition points against the N centres.  Pairs per step = 6 N^2.

  value   device path with the block's inputs (centres, factors, query points) resident in HBM:
          IM -> NNLS -> weights -> eval, timed with CUDA events on the context streams.
  e2e     the same iteration through the reference-facing host API (ncm_b200_esmcmc_run over
          ncm_stats_dist_* -> C ABI) with HOST buffers: prepare_kernel on the host, uploads, IM, NNLS,
          host-ordered proposal sampling, batched eval, likelihood + accept; wall clock.
  --impl reference   the CPU oracle port of the reference's algorithm (OpenMP + OpenBLAS, all host cores).

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "kernel_pair_evals_per_s"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="apes_vkde_gauss_mvnd10_w4096",
                    choices=["apes_vkde_gauss_mvnd10_w4096", "eval_sweep", "prepare_interp", "apes_e2e", "cv"])
    ap.add_argument("--target", default="funnel", choices=["funnel", "rosenbrock", "mvnd"], help="--workload apes_e2e: configs[3] (funnel) / configs[0] (rosenbrock)")
    ap.add_argument("--over-smooth", type=float, default=None)
    ap.add_argument("--walkers", type=int, default=None, help="default: 4096 at N = 1 (configs[1]); 32768 for the sharded N > 1 workload (configs[2]/[3] class)")
    ap.add_argument("--dim", type=int, default=None, help="default: 10 at N = 1; 20 for the sharded N > 1 workload")
    ap.add_argument("--sweep-q", type=int, default=65536)
    ap.add_argument("--sweep-n", type=int, default=65536)
    ap.add_argument("--sd", default="vkde", choices=["kde", "vkde"])
    ap.add_argument("--kernel", default="gauss", choices=["gauss", "st3", "cauchy"])
    ap.add_argument("--cv", default="split", choices=["split", "split_nofit", "loo"], help="--workload cv: the cross-validation mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--apes-multi", default="sharded", choices=["replicas", "sharded"],
                    help="APES workload at N > 1.  sharded (default): ONE ensemble, interpolation-matrix rows and query rows sharded over the ranks, "
                         "centres replicated, NCCL all-reduce of the normal equations and all-gather of the densities on the data path (the north-star "
                         "partitioning, strong scaling), on a configs[2]/[3]-class size (32768 walkers = 16384 centres, d = 20) where sharding can pay.  "
                         "replicas: one independent ensemble per GPU, no data-path collective (weak scaling; how several chains are run).")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":   # the reference arm describes the workload of the --gpus N arm it sits next to, torchrun or not
        world = max(world, a.gpus)
    big = world > 1 and a.apes_multi == "sharded" and a.workload == "apes_vkde_gauss_mvnd10_w4096"
    if a.walkers is None:
        a.walkers = 32768 if big else 4096
    if a.dim is None:
        a.dim = 20 if big else 10
    return a


# ---------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = {"hbm_gbs": 6650.0, "source": "fallback"}
    if os.path.exists(p):
        try:
            peaks.update(json.load(open(p)))
            peaks["source"] = "measured"
        except Exception:
            pass
    # FP64 peaks are not in MEASURED_PEAKS.json (bf16 + HBM only): measured on this pool with
    # tools/microbench (profiles/r01_fp64_peaks.jsonl): cuBLAS DGEMM 8192^3 and a DMMA issue loop.
    peaks["fp64_dgemm_tflops"] = 35.46
    peaks["fp64_dmma_tflops"] = 37.1
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device, self.samples, self._stop, self._t = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        # rank 0 only: nvidia-smi costs ~0.1 s of a host core per call, and the ranks share the host cores with the NNLS decisions
        if int(os.environ.get("RANK", "0")) == 0:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_problem(W, d, seed=1):
    """configs[1] inputs: ncm_data_gauss_cov_mvnd_new_full (d, sigma in [2e-2, 5e-2], cor_level 30, mu in [1, 2]) and W walkers drawn
    from the target."""
    from oracle import ncm_oracle as O
    from helpers import mvnd_problem

    mu, cov, X, m2lnL = mvnd_problem(O, d, W, seed)
    tgt = O.Target(O.TARGET_MVND, d, np.full(d, -50.0), np.full(d, 50.0), mu=mu, cov=cov)
    return mu, cov, tgt, X, m2lnL


def make_problem_b200(S, W, d, seed=1, sigma=(2e-2, 5e-2), cor_level=30.0, mu_range=(1.0, 2.0)):
    """Same recipe and the same MT19937 stream as make_problem, drawn with the product's own NcmRNG mirror
    (bit-identical to the oracle's, tests/test_oracle_known_answers.py) so that the GPU arm never touches oracle/."""
    rng = S.RNG(seed)
    P = np.zeros((d, d))
    cm = np.eye(d)
    for k in range(d - 1):          # ncm_matrix_fill_rand_cor, ncm_matrix.c:1687-1735
        for i in range(k + 1, d):
            p = (rng.beta_gen(cor_level, cor_level) - 0.5) * 2.0
            P[k, i] = p
            for l in range(k - 1, -1, -1):
                p = p * np.sqrt((1.0 - P[l, i] ** 2) * (1.0 - P[l, k] ** 2)) + P[l, i] * P[l, k]
            cm[k, i] = cm[i, k] = p
    for k in range(d):              # ncm_matrix_fill_rand_cov, :1752-1770
        s = rng.uniform_gen(sigma[0], sigma[1])
        cm[:, k] *= s
        cm[k, :] *= s
    mu = np.array([rng.uniform_gen(*mu_range) for _ in range(d)])
    L = np.linalg.cholesky(cm)
    z = np.array([[rng.gaussian_gen(0.0, 1.0) for _ in range(d)] for _ in range(W)])
    X = np.ascontiguousarray(mu + z @ L.T)
    return mu, cm, np.ascontiguousarray(L.T), X, np.einsum("ij,ij->i", z, z)


KT = {"gauss": ("GAUSS", 0, 1.0), "st3": ("ST3", 1, 3.0), "cauchy": ("CAUCHY", 1, 1.0)}


# ---------------------------------------------------------------------------------------------------
def cpu_apes(args, W, d, iters, nthreads, warm=True):
    """The reference algorithm on the host cores (oracle port): returns seconds per iteration + stage timers."""
    from oracle import ncm_oracle as O

    mu, cov, tgt, X, m2lnL = make_problem(W, d)
    _, okind, nu = KT[args.kernel]
    O.lib().orc_set_blas_threads(nthreads)
    os.environ["OMP_NUM_THREADS"] = str(nthreads)
    ap = O.APES(W, d, O.SD_VKDE, okind, nu, over_smooth=1.0, use_interp=True, use_threads=True)
    theta, ml = X.copy(), m2lnL.copy()
    rng = O.RNG(1234)
    accs = []
    if warm:
        accs.append(ap.run(tgt, theta, ml, 1, rng, nthreads=nthreads))   # warm-up iteration
    t0 = time.perf_counter()
    accs.append(ap.run(tgt, theta, ml, iters, rng, nthreads=nthreads))
    dt = (time.perf_counter() - t0) / iters
    cpu_apes.last_accepted = np.concatenate(accs, axis=0)   # accepted / rejected of every iteration from the common start, seed 1234
    return dt, ap.timers()


def apes_config(args, W, d, N, world):
    """`config` and `scaling` of the APES workload: ONE definition for both arms, so the driver compares like with like."""
    sharded = world > 1 and args.apes_multi == "sharded"
    replicas = 1 if sharded or world == 1 else world
    part = ("single GPU" if world == 1 else
            f"ONE ensemble over {world} GPUs: interpolation-matrix rows and query rows sharded, centres and factors replicated, ncclAllReduce of the "
            "normal equations + ncclAllGather of the densities on the data path" if sharded else
            f"{replicas} independent ensembles, one per GPU, no data-path collective")
    cfg = {"workload": f"APES iteration, VKDE {args.kernel} kernel, {d}-D MVND, {W} walkers (6 N^2 pairs/step, N={N})",
           "multi_gpu": "single" if world == 1 else args.apes_multi, "partitioning": part,
           "l2_policy": "each half-step streams a fresh IM and two normal-matrix-sized buffers (3 x %.0f MB at N = %d); the two half-steps alternate "
                        "contexts, so no timed kernel re-reads data left by its previous launch (working set per step > 126 MB L2)" % (N * N * 8 / 1e6, N)}
    return cfg, ("weak" if replicas > 1 else "strong" if world > 1 else "weak")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, d = args.walkers, args.dim
    ncores = os.cpu_count() or 1
    N = W // 2
    pairs = 6.0 * N * N
    # warm-up happens inside cpu_apes (1 iteration); args.warmup - 1 further untimed iterations are folded into it.  The sharded N > 1
    # workload (32768 walkers) costs the CPU about a minute per iteration: its sample is bounded to ONE iteration without warm-up.
    big = W > 8192
    iters = 1 if big else max(1, args.steps)
    dt, timers = cpu_apes(args, W, d, iters, ncores, warm=not big)
    val = pairs / dt
    cfg, scaling = apes_config(args, W, d, N, args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "note": "CPU oracle port of the reference algorithm on the host cores (the reference itself cannot be built here: no GLib/GSL/meson); "
                "one ensemble of the arm's size, whatever --gpus says",
        "walker_steps_per_s": W / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": ncores, "kind": "port",
                         "sample": f"{iters} full APES iteration(s) at W={W}, d={d}" + (" without warm-up (bounded sample: ~1 min of CPU per iteration)" if big else "") +
                                   " (OpenMP walkers-parallel eval, centre-parallel IM, threaded OpenBLAS NNLS)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_s_total": {k: float(v) for k, v in timers.items()},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def _device_half_steps(torch, capi, sds, ctxs, theta, ml, N, d, steps, warmup, shard, clock_device):
    """`value`: device-resident half-steps through the C ABI on the contexts behind sd0 / sd1 (centres, packed factors and lnnorms as the
    last prepare_kernel left them).  shard = (world, rank) when the contexts carry a communicator in auto-shard mode: compute_IM then takes
    this rank's row block, the NNLS all-reduces the normal equations, the 2N query rows are split and the densities all-gathered."""
    world, rank = shard
    gctx, dQ, dOut, dAll, rowscale, hrefs = [], [], [], [], [], []
    for b in range(2):
        mlc = ml[N:] if b == 0 else ml[:N]
        c = ctxs[b]
        # sd_b's centres are the OTHER half of the ensemble (walker_apes.c:751-811).  Block 1 moved after sd0 was last prepared, so
        # re-run prepare_kernel on the current positions: centres, factors and m2lnL in the context are then those of a real half-step.
        sds[b].reset()
        for xrow in (theta[N:] if b == 0 else theta[:N]):
            sds[b].add_obs(xrow)
        sds[b].prepare()
        c.n_kernels = c.n_obs = N
        c.d = d
        f = np.exp(-0.5 * (mlc - mlc.min()))
        rowscale.append(1.0 / f)
        blk = theta[:N] if b == 0 else theta[N:]
        q_all = np.vstack([blk + 1e-3, blk])            # theta*_k and theta_k of the block: 2N query points
        cap = (2 * N + world - 1) // world
        q0, q1 = (2 * N * rank) // world, (2 * N * (rank + 1)) // world
        dQ.append(torch.from_numpy(np.ascontiguousarray(q_all[q0:q1])).cuda())
        dOut.append(torch.zeros(cap, dtype=torch.float64, device="cuda"))
        dAll.append(torch.zeros(cap * world, dtype=torch.float64, device="cuda") if world > 1 else None)
        gctx.append(c)
        hrefs.append(sds[b].get_href())
    streams = [torch.cuda.ExternalStream(c.stream) for c in gctx]

    def half_step(b):
        c = gctx[b]
        c.compute_IM(rowscale[b])
        x, rnorm, st = c.nnls_solve()
        w = (1.0 - 0.01) * x / x.sum() + 0.01 / N
        c.set_weights(w, hrefs[b])
        c.eval_m2lnp_dev(dQ[b].shape[0], dQ[b].data_ptr(), d, dOut[b].data_ptr())
        if world > 1:
            c.allgather_dev(dOut[b].data_ptr(), dAll[b].data_ptr(), dOut[b].shape[0])
        return st

    import torch.distributed as dist

    for _ in range(max(warmup, 3)):
        half_step(0)
        half_step(1)
    for c in gctx:
        c.synchronize()
        c.reset_timers()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(2 * steps)]
    with ClockSampler(clock_device) as clk:
        for it in range(steps):
            for b in range(2):
                e0, e1 = ev[2 * it + b]
                e0.record(streams[b])
                half_step(b)
                e1.record(streams[b])
        for c in gctx:
            c.synchronize()
    torch.cuda.synchronize()
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    if world > 1:
        t_dev = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.barrier()
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
        dev_ms = float(t_dev.item())
    launches = sum(c.get_timers()[1] for c in gctx)
    # per-stage device time of the step (separate pass with the CUDA-event stage timers on)
    for c in gctx:
        c.enable_timers(True)
        c.reset_timers()
    nroof = 3
    agg = {"n_chol": 0, "chol_flops": 0.0, "n_lowrank": 0, "n_trinv": 0, "lowrank_flops": 0.0, "n_lowrank_fallback": 0, "max_lowrank_k": 0}
    for _ in range(nroof):
        for b in range(2):
            st = half_step(b)
            for k in agg:
                agg[k] = max(agg[k], st[k]) if k == "max_lowrank_k" else agg[k] + st[k]
    tm = {k: 0.0 for k in capi.T_NAMES}
    for c in gctx:
        t, _ = c.get_timers()
        for k in tm:
            tm[k] += t[k] / nroof
        c.enable_timers(False)
    for k in agg:
        if k != "max_lowrank_k":
            agg[k] = agg[k] / nroof
    return dev_ms / steps, launches, tm, agg, clk.summary()


def run_b200(args):
    import torch
    import torch.distributed as dist

    from numcosmo_b200 import capi
    from numcosmo_b200 import stats_dist as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:   # the ranks share the host cores: no OpenMP oversubscription in the host-side prepare_kernel
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; numcosmo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = S.lib()
    lib.ncm_b200_set_device(local_rank)
    if world > 1:
        lib.ncm_b200_set_num_threads(max(1, (os.cpu_count() or 1) // world))

    W, d = args.walkers, args.dim
    N = W // 2
    pairs_step = 6.0 * N * N
    sharded = world > 1 and args.apes_multi == "sharded"
    nshard, shard_rank = (world, rank) if sharded else (1, 0)     # ranks one ensemble is split over
    replicas = 1 if sharded or world == 1 else world              # independent ensembles (one per GPU)
    # sharded: every rank holds the SAME ensemble and draws the SAME generator stream (SPMD); replicas: one ensemble per rank
    mu, cov, U_tgt, X, m2lnL = make_problem_b200(S, W, d, seed=1 + (rank if replicas > 1 else 0))
    ktn, okind, nu = KT[args.kernel]
    lb, ub = np.full(d, -50.0), np.full(d, 50.0)
    peaks = load_peaks()

    def new_apes():
        ap = S.FitESMCMCWalkerAPES(W, d, S.FitESMCMCWalkerAPESMethod.VKDE, getattr(S.FitESMCMCWalkerAPESKType, ktn), 1.0, True)
        ap.set_use_threads(True)
        return ap

    def e2e_run(apes, theta, ml, rng):
        """end to end through the host API with HOST buffers; returns (seconds per iteration, accept rate, h2d, d2h, launches, stage split)"""
        acc_w, _ = apes.run("mvnd", lb, ub, theta, ml, max(args.warmup, 1), rng, target_args=(mu, U_tgt), record_accept=True)
        e2e_run.acc_warm = acc_w
        sds = apes.peek_sds()
        ctxs = [capi.Context.borrowed(lib.ncm_stats_dist_b200_peek_ctx(sd._h)) for sd in sds]
        for c in ctxs:
            c.reset_timers()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        acc, _ = apes.run("mvnd", lb, ub, theta, ml, args.steps, rng, target_args=(mu, U_tgt), record_accept=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        traffic = [c.get_traffic() for c in ctxs]
        launches = sum(c.get_timers()[1] for c in ctxs)
        apes.enable_timers(True)
        for c in ctxs:
            c.reset_timers()
        _, stage = apes.run("mvnd", lb, ub, theta, ml, 2, rng, target_args=(mu, U_tgt), record_accept=False)
        apes.enable_timers(False)
        if world > 1:   # the job's rate is set by the slowest rank
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, acc, sum(t[0] for t in traffic) / args.steps, sum(t[1] for t in traffic) / args.steps, launches / args.steps, \
            {k: v / 2 for k, v in stage.items()}, sds, ctxs

    # ---------------- single-GPU arm of the SAME workload (sharded mode only): the denominator of the strong-scaling speed-up ----------
    single = None
    if sharded:
        ap1 = new_apes()
        th1, ml1 = X.copy(), m2lnL.copy()
        dt1, acc1, _, _, _, _, sds1, ctxs1 = e2e_run(ap1, th1, ml1, S.RNG(1234))
        ms1, _, tm1, agg1, _ = _device_half_steps(torch, capi, sds1, ctxs1, th1, ml1, N, d, args.steps, args.warmup, (1, 0), local_rank)
        single = {"value": pairs_step / (ms1 * 1e-3), "ms_per_step": ms1, "e2e_value": pairs_step / dt1, "e2e_ms_per_step": dt1 * 1e3,
                  "step_share_ms": {k: round(v, 4) for k, v in tm1.items()}, "accept_hash": int(np.packbits(acc1).astype(np.uint64).sum()),
                  "note": "this rank's GPU alone, unsharded, same ensemble and generator seed (all ranks measure it at the same time; rank 0's is reported)"}
        del ap1, sds1, ctxs1
        torch.cuda.empty_cache()

    # ---------------- e2e: host API, host buffers, whole iteration ----------------
    apes = new_apes()
    if sharded:
        apes.comm_init_from_torch()      # one NCCL communicator per NcmStatsDist context, auto-shard on
    theta, ml = X.copy(), m2lnL.copy()
    rng = S.RNG(1234)
    e2e_dt, acc, h2d_step, d2h_step, e2e_launches, stage, sds, ctxs = e2e_run(apes, theta, ml, rng)
    accept_rate = float(acc.mean())
    accept_hash = int(np.packbits(acc).astype(np.uint64).sum())

    # ---------------- value: device-resident half-steps through the C ABI ----------------
    ms_per_step, launches, tm, agg, clocks = _device_half_steps(torch, capi, sds, ctxs, theta, ml, N, d, args.steps, args.warmup, (nshard, shard_rank),
                                                                local_rank)
    value = replicas * pairs_step / (ms_per_step * 1e-3)

    # ---------------- roofline of the dominant kernel ----------------
    rows_local = N // nshard
    P64 = peaks["fp64_dgemm_tflops"]
    nchol = max(agg["n_chol"], 1e-9)
    chol_ach = agg["chol_flops"] / (tm["chol"] * 1e-3) / 1e12 if tm["chol"] > 0 else 0.0
    syrk_flops = float(rows_local) * N * N                 # n_obs . n_kernels^2 per half-step on this rank, 2 flops each
    syrk_ach = 2.0 * syrk_flops / (tm["syrk"] * 1e-3) / 1e12
    lr_ach = agg["lowrank_flops"] / (tm["lowrank"] * 1e-3) / 1e12 if tm["lowrank"] > 0 else 0.0
    eval_pairs = pairs_step / nshard
    eval_ach = eval_pairs * (d * d + 2.0 * d + 1.0) / ((tm["eval"] + tm["IM"]) * 1e-3) / 1e12
    kernels = {
        "chol": {"kernel": ("chol_fused_kernel: single-launch Cholesky solve (dposv) of the first passive-set system of each NNLS, one persistent cooperative "
                            "launch (spine CTA + 147 tile workers, DMMA.8x8x4 updates)") if N <= 4096 else
                           "chol_diag / chol_panel / ata_kernel<AtaBig> / chol_backsolve: blocked right-looking Cholesky solve with look-ahead (chol.cu)",
                 "achieved": chol_ach, "frac": chol_ach / P64, "ms_per_step": tm["chol"], "launches_per_step": agg["n_chol"],
                 "flops_per_launch": agg["chol_flops"] / nchol, "ms_per_launch": tm["chol"] / nchol,
                 # one `ncu --set full` capture at n = 2048 (profiles/r01f_ncu_full_chol_fused.txt): dram read + write per launch
                 "traffic": 17544192 + 4864 if N <= 4096 else None},
        "lowrank": {"kernel": "trinv_step1/2 + gemm_tn_splitk + syrk_splitk (DMMA.8x8x4 GEMMs) + lr_small_kernel (k x k L J L^T in shared memory) + row products: "
                              "passive-set systems after the first solved by low-rank modification of its factor (lowrank.cu)",
                    "achieved": lr_ach, "frac": lr_ach / P64, "ms_per_step": tm["lowrank"], "solves_per_step": agg["n_lowrank"],
                    "trinv_per_step": agg["n_trinv"], "fallbacks_per_step": agg["n_lowrank_fallback"], "max_k": agg["max_lowrank_k"],
                    "flops_model": "|B|^3/3 per triangular inverse + 2 |B|^2 (k + 1) per solve"},
        "syrk": {"kernel": "ata_kernel<AtaBig> (M = IM^T IM over this rank's %d rows, n = %d, DMMA.8x8x4)" % (rows_local, N), "achieved": syrk_ach,
                 "frac": syrk_ach / P64, "ms_per_step": tm["syrk"], "ms_per_launch": tm["syrk"] / 2.0},
        "vkde_eval+IM": {"kernel": "vkde_kernel<%d,0/1>" % d, "pairs_per_s": eval_pairs / ((tm["eval"] + tm["IM"]) * 1e-3), "achieved": eval_ach,
                         "frac": eval_ach / P64, "ms_per_step": tm["eval"] + tm["IM"]},
    }
    dom = max(("chol", "lowrank", "syrk", "vkde_eval+IM"), key=lambda k: kernels[k]["ms_per_step"])
    roofline = {"bound": "tensor", "kernel": kernels[dom]["kernel"], "achieved": kernels[dom]["achieved"], "peak": P64, "unit": "TFLOP/s",
                "frac": kernels[dom]["frac"], "traffic": kernels[dom].get("traffic"),
                "dominant": dom, "ms_per_step": kernels[dom]["ms_per_step"],
                "note": "dominant stage of the timed step by CUDA-event stage time.  chol: latency-bound by construction (n sequential pivots, 16 - 17 us per "
                        "64-column phase, tools/chol_trace.py), now ONE factorisation per NNLS instead of 6 (the other passive sets go through `lowrank`); "
                        "lowrank: GEMM-shaped work at small sizes (|B| ~ 2000, k <= 200) plus one single-CTA k x k factorisation per solve.  "
                        "Every stage is listed under `others` with its own fraction of the FP64 DGEMM peak.",
                "peak_source": "FP64 is not in MEASURED_PEAKS.json (bf16 + HBM only); cuBLAS DGEMM 8192^3 measured on this pool, "
                               "profiles/r01_fp64_peaks.jsonl",
                "others": {k: v for k, v in kernels.items() if k != dom},
                "step_share_ms": {k: round(v, 4) for k, v in tm.items()}}

    # ---------------- CPU baseline (rank 0, bounded sample) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # reported at N = 1 only (the other ranks would sit idle behind it)
        ncores = os.cpu_count() or 1
        cdt, ctm = cpu_apes(args, W, d, 3, ncores)
        # both arms start from the same ensemble with the same generator seed: their accepted / rejected sequences must be the same
        acc_cpu = cpu_apes.last_accepted
        acc_gpu = np.concatenate([e2e_run.acc_warm, acc], axis=0)
        ncmp = min(len(acc_cpu), len(acc_gpu))
        cpu = {"value": pairs_step / cdt, "unit": UNIT, "cores": ncores, "kind": "port",
               "sample": f"3 full APES iterations at W={W}, d={d} on the host cores (oracle port; 1 warm-up iteration)",
               "ms_per_step": cdt * 1e3, "walker_steps_per_s": W / cdt,
               "accepted_sequence_identical_to_gpu_arm": bool(np.array_equal(np.asarray(acc_cpu[:ncmp]).astype(bool), np.asarray(acc_gpu[:ncmp]).astype(bool))),
               "iterations_compared": int(ncmp)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": apes_config(args, W, d, N, world)[1], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": apes_config(args, W, d, N, world)[0],
            "walker_steps_per_s": replicas * W / (ms_per_step * 1e-3),
            "e2e": {"value": replicas * pairs_step / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": replicas * h2d_step, "d2h_bytes_per_step": replicas * d2h_step,
                    "ms_per_step": e2e_dt * 1e3, "walker_steps_per_s": replicas * W / e2e_dt, "accept_rate": accept_rate, "accept_hash": accept_hash,
                    "stage_ms_per_step": {k: round(v, 3) for k, v in stage.items()}, "gpu_launches_per_step": e2e_launches,
                    "api": "ncm_b200_esmcmc_run -> ncm_stats_dist_prepare_interp / ncm_stats_dist_eval_m2lnp_array -> C ABI" +
                           (" in multi-rank (SPMD) mode: every rank runs the chain with the same generator seed, the density work behind the calls is "
                            "sharded (ncm_stats_dist_b200_comm_init); h2d / d2h are this rank's bytes" if sharded else " (one GPU per ensemble)")},
            "gpu_launches": int(launches) * replicas,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if sharded:
            line["comm"] = {"ms_per_step": tm["comm"], "allreduce_M_bytes_per_half_step": N * ((N + 7) // 8 * 8) * 8,
                            "what": "ncclAllReduce of M = IM^T IM (once per half-step), of b, A^T r and |r|^2 (per outer NNLS iteration), ncclAllGather of the 2N densities"}
            line["single_gpu_same_workload"] = single
            line["strong_scaling_speedup"] = {"value": value / single["value"], "e2e": (pairs_step / e2e_dt) / single["e2e_value"],
                                              "accepted_sequence_identical_to_single_gpu": accept_hash == single["accept_hash"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def sweep_problem(d, N, Q, sd, seed=5):
    """configs[4] inputs: centres and queries i.i.d. from a d-dim MVND, Dirichlet-like weights
    with 10 % exact zeros, href = 1.  For VKDE the per-centre factors are synthetic: the global factor times a random
    well-conditioned upper-triangular perturbation (running the kNN prepare_kernel on 65536 centres is not what is timed here)."""
    rs = np.random.default_rng(seed)
    sig = rs.uniform(2e-2, 5e-2, size=d)
    R = rs.normal(size=(d, d)) / np.sqrt(d)
    cor = 0.7 * np.eye(d) + 0.3 * (R @ R.T)
    cov = cor * np.outer(sig, sig)
    Ug = np.ascontiguousarray(np.linalg.cholesky(cov).T)          # upper factor, cov = Ug^T Ug
    mu = rs.uniform(1.0, 2.0, size=d)
    C = np.ascontiguousarray(mu + rs.normal(size=(N, d)) @ Ug)
    X = np.ascontiguousarray(mu + rs.normal(size=(Q, d)) @ Ug)
    w = rs.uniform(size=N)
    w[rs.uniform(size=N) < 0.1] = 0.0
    w /= w.sum()
    lndet_g = 2.0 * np.log(np.diag(Ug)).sum()
    if sd == "kde":
        return mu, Ug, C, X, w, None, lndet_g
    U_all = np.empty((N, d, d))
    blk = 8192
    for i0 in range(0, N, blk):
        n = min(blk, N - i0)
        T = np.triu(rs.normal(size=(n, d, d)) * (0.15 / np.sqrt(d)), 1)
        T[:, np.arange(d), np.arange(d)] = rs.uniform(0.3, 0.6, size=(n, d))
        U_all[i0:i0 + n] = np.triu(T @ Ug)
    return mu, Ug, C, X, w, U_all, lndet_g


def run_sweep(args):
    """--workload eval_sweep: batched eval_m2lnp, Q proposals x N centres (BASELINE.json configs[4])."""
    import torch
    import torch.distributed as dist

    from numcosmo_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; numcosmo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    d, N, Q = args.dim, args.sweep_n, args.sweep_q
    ktn, okind, nu = KT[args.kernel]
    peaks = load_peaks()
    mu, Ug, Cn, X, w, U_all, lndet_g = sweep_problem(d, N, Q, args.sd)
    ctx = capi.Context(local_rank)
    ctx.set_kernel(okind, nu, d)
    lg = lambda z: float(__import__("math").lgamma(z))
    def lnnorm_of(lndet):   # _kernel_gauss.c:201-207 / _kernel_st.c:240-252
        if okind == 0:
            return 0.5 * (d * np.log(2.0 * np.pi) + lndet)
        return lg(nu / 2.0) - lg((nu + d) / 2.0) + 0.5 * d * (np.log(np.pi) + np.log(nu)) + 0.5 * lndet
    if args.sd == "kde":
        Zc = np.ascontiguousarray(np.linalg.solve(Ug.T, Cn.T).T)     # invUsample = sample . U^-1
        ctx.upload_kde(Zc, N, Ug, lnnorm_of(lndet_g))
    else:
        lnn = lnnorm_of(2.0 * np.log(np.abs(U_all[:, np.arange(d), np.arange(d)])).sum(axis=1))
        ctx.upload_vkde(Cn, N, U_all, lnn)
        del U_all
    ctx.set_weights(w, 1.0)
    q0, q1 = (Q * rank) // world, (Q * (rank + 1)) // world
    Xl = np.ascontiguousarray(X[q0:q1])
    hX = torch.from_numpy(Xl).pin_memory()
    dX = hX.cuda()
    dOut = torch.empty(q1 - q0, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream)
    # L2 policy: the centre records (N x REC x 8 B) are re-streamed once per query tile; between timed steps a 256 MB buffer is
    # overwritten so that no step starts with the records of the previous one in the 126 MB L2.
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")

    def step():
        ctx.eval_m2lnp_dev(q1 - q0, dX.data_ptr(), d, dOut.data_ptr())

    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()
    ctx.reset_timers()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        for it in range(args.steps):
            with torch.cuda.stream(stream):
                flush.fill_(float(it))
            ev[it][0].record(stream)
            step()
            ev[it][1].record(stream)
        ctx.synchronize()
    torch.cuda.synchronize()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.barrier()
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms = float(t_dev.item()) / args.steps
    pairs = float(Q) * N
    value = pairs / (ms * 1e-3)
    launches = ctx.get_timers()[1]

    # e2e: host buffers in, host results out, through the C ABI call a user makes (ncm_sd_gpu_eval_m2lnp)
    out = np.empty(q1 - q0)
    ctx.eval_m2lnp(Xl, out)
    ctx.reset_timers()
    ne = max(2, min(args.steps, 5))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(ne):
        ctx.eval_m2lnp(Xl, out)
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / ne], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_dt = float(e2e_dt.item())
    h2d, d2h = ctx.get_traffic()

    # roofline of the eval kernel (it is > 99 % of the step): flops per pair from SURVEY.md section 8d
    P64 = peaks["fp64_dgemm_tflops"]
    if args.sd == "kde":
        fl_pair = 2.0 * d + 3.0
        kname = "kde_kernel (DMMA.8x8x4 GEMM + online LSE)"
    else:
        fl_pair = d * d + 2.0 * d + 1.0
        kname = ("vkde_mma_kernel (per-centre W = L^-1 products on DMMA.8x8x4 + online LSE)" if d >= 21 else
                 "vkde_kernel (per-centre forward substitution in registers + online LSE)")
    ach = (pairs / world) * fl_pair / (ms * 1e-3) / 1e12
    # epilogue-inclusive model: per pair the kernel function + online log-sum-exp run on the same FP64 datapath as the
    # contraction: exp_nonpos_fast = 15 FP64 operations + 3 for the LSE update = 36 flop-equivalents (Gauss); Student-t adds
    # log1p_nonneg_fast (14 operations) + the kappa multiply = 30 more
    epi = 36.0 + (30.0 if okind == 1 else 0.0)
    ach_epi = (pairs / world) * (fl_pair + epi) / (ms * 1e-3) / 1e12
    rec_bytes = None
    if args.sd == "vkde":
        rec_bytes = (d * (d + 1) / 2 + d + 2) * 8.0
    # dram read + write per launch from the `ncu --set full` captures under profiles/ (only for the captured configurations)
    captured = {("vkde", "gauss", 30, 65536, 65536): 1449735000 + 11664640, ("vkde", "gauss", 30, 32768, 32768): 387631360 + 7327744,
                ("kde", "gauss", 10, 65536, 65536): 12617216, ("vkde", "gauss", 20, 65536, 65536): 139484928 + 9936384}
    traffic = captured.get((args.sd, args.kernel, d, Q, N)) if world == 1 else None
    roofline = {"bound": "tensor", "kernel": kname, "achieved": ach, "peak": P64, "unit": "TFLOP/s", "frac": ach / P64, "traffic": traffic,
                "pipe_note": "ncu (profiles/r01d_ncu_full_*.txt, 65536^2): sm__pipe_shared_cycles_active (the FP64 datapath DMMA and DFMA share) 89 % of the "
                             "elapsed cycles for vkde_mma d=30 and 87 % for kde d=10; 74 % (LSU wavefronts 62 %) for the substitution kernel at d=20",
                "flops_per_pair": fl_pair, "with_epilogue": {"flop_equiv_per_pair": fl_pair + epi, "achieved": ach_epi, "frac": ach_epi / P64},
                "streamed_GBs": (None if rec_bytes is None else ((q1 - q0) / 128.0) * N * rec_bytes / (ms * 1e-3) / 1e9),
                "peak_source": "FP64 is not in MEASURED_PEAKS.json (bf16 + HBM only); cuBLAS DGEMM 8192^3 measured on this pool, "
                               "profiles/r01_fp64_peaks.jsonl"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sweep(args, d, N, Q)
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"batched eval_m2lnp sweep: {Q} proposals x {N} centres, d={d}, {args.sd.upper()} {args.kernel} kernel; "
                                       "query rows sharded over ranks, centres replicated",
                           "l2_policy": "256 MB flush buffer overwritten before every timed step"},
                "e2e": {"value": pairs / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d / ne, "d2h_bytes_per_step": d2h / ne, "ms_per_step": e2e_dt * 1e3,
                        "api": "ncm_sd_gpu_eval_m2lnp (host X in, host m2lnp out)"},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_sweep(args, d, N, Q, budget_s=15.0):
    """Oracle port of eval_m2lnp on the host cores for a bounded query sub-sample (all centres), extrapolated linearly in Q."""
    from oracle import ncm_oracle as O

    ncores = os.cpu_count() or 1
    _, okind, nu = KT[args.kernel]
    mu, Ug, Cn, X, w, U_all, lndet_g = sweep_problem(d, min(N, 16384), 4096, args.sd)
    Nc = Cn.shape[0]
    sd = O.StatsDist(O.SD_KDE if args.sd == "kde" else O.SD_VKDE, okind, d, nu)
    sd.set_use_threads(True)
    sd.set_local_frac(0.01)
    sd.add_obs_matrix(Cn)
    assert sd.prepare() == 0
    qs = 256
    t0 = time.perf_counter()
    sd.eval_m2lnp_batch(X[:qs], ncores)
    dt = time.perf_counter() - t0
    qn = int(max(qs, min(4096, qs * budget_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    sd.eval_m2lnp_batch(X[:qn], ncores)
    dt = time.perf_counter() - t0
    return {"value": qn * float(Nc) / dt, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": f"{qn} queries x {Nc} centres (oracle port, OpenMP over queries as ncm_fit_esmcmc.c:2158; factors from the oracle's own "
                      f"prepare_kernel with local_frac 0.01); pairs/s is size-independent, so it extrapolates linearly to {Q} x {N}"}


# ---------------------------------------------------------------------------------------------------
def run_prepare_interp(args):
    """--workload prepare_interp (BASELINE.json configs[2]): interpolation matrix + NNLS weights for N centres in d dimensions,
    IM row blocks sharded over ranks, normal equations all-reduced by NCCL inside the C ABI, passive-set Cholesky replicated."""
    import torch
    import torch.distributed as dist

    from numcosmo_b200 import capi, shard
    from numcosmo_b200 import stats_dist as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; numcosmo_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = S.lib()
    lib.ncm_b200_set_device(local_rank)
    d, N = args.dim, args.sweep_n
    ktn, okind, nu = KT[args.kernel]
    peaks = load_peaks()
    rs = np.random.default_rng(2)
    sig = rs.uniform(2e-2, 5e-2, size=d)
    R = rs.normal(size=(d, d)) / np.sqrt(d)
    cov = (0.7 * np.eye(d) + 0.3 * (R @ R.T)) * np.outer(sig, sig)
    Ug = np.linalg.cholesky(cov).T
    z = rs.normal(size=(N, d))
    X = np.ascontiguousarray(rs.uniform(1.0, 2.0, size=d) + z @ Ug)
    m2lnp = np.einsum("ij,ij->i", z, z)                     # exact MVND -2 ln L at the centres
    kern = S.StatsDistKernelGauss(d) if okind == 0 else S.StatsDistKernelST(d, nu)
    sd = (S.StatsDistVKDE if args.sd == "vkde" else S.StatsDistKDE)(kern, S.StatsDistCV.NONE)
    sd.set_use_threads(True)
    for x in X:
        sd.add_obs(x)
    t0 = time.perf_counter()
    sd.prepare()                                             # prepare_kernel (kNN + local covariances + factors) + upload
    torch.cuda.synchronize()
    t_prepare_kernel = time.perf_counter() - t0
    c = capi.Context.borrowed(lib.ncm_stats_dist_b200_peek_ctx(sd._h))
    c.n_kernels = c.n_obs = N
    c.d = d
    if world > 1:
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        c.comm_init(world, rank, uid[0])
    r0, r1 = shard.row_range(N, rank, world)
    c.set_row_shard(r0, r1 - r0)
    rowscale = 1.0 / np.exp(-0.5 * (m2lnp - m2lnp.min()))
    c.set_href(sd.get_href())
    stream = torch.cuda.ExternalStream(c.stream)

    def step():
        c.compute_IM(rowscale)
        return c.nnls_solve()

    for _ in range(max(1, min(args.warmup, 2))):
        x, rnorm, st = step()
    c.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = max(1, min(args.steps, 5))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with ClockSampler(local_rank) as clk:
        for it in range(steps):
            ev[it][0].record(stream)
            x, rnorm, st = step()
            ev[it][1].record(stream)
        c.synchronize()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    if world > 1:
        dist.barrier()
        ms = shard.max_over_ranks(ms)
    # per-stage device time (one more step with the stage timers on)
    c.enable_timers(True)
    c.reset_timers()
    step()
    tm, launches = c.get_timers()
    c.enable_timers(False)
    pairs = float(N) * N
    P64 = peaks["fp64_dgemm_tflops"]
    syrk_ach = float(r1 - r0) * N * N / (tm["syrk"] * 1e-3) / 1e12
    chol_ach = st["chol_flops"] / (tm["chol"] * 1e-3) / 1e12
    dom = "syrk" if tm["syrk"] >= tm["chol"] else "chol"
    roofline = {"bound": "tensor",
                "kernel": ("ata_kernel<AtaBig> (normal equations M = IM^T IM, DMMA.8x8x4)" if dom == "syrk" else
                           "blocked Cholesky solve (chol_diag/panel kernels + ata_kernel trailing updates on DMMA.8x8x4)"),
                "achieved": syrk_ach if dom == "syrk" else chol_ach, "peak": P64, "unit": "TFLOP/s",
                "frac": (syrk_ach if dom == "syrk" else chol_ach) / P64, "traffic": None,
                "syrk_tflops": syrk_ach, "chol_tflops": chol_ach, "n_chol": st["n_chol"], "n_passive": st["n_passive"],
                "stage_ms": {k: round(v, 3) for k, v in tm.items()},
                "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool, profiles/r01_fp64_peaks.jsonl"}
    if rank == 0:
        line = {"metric": METRIC, "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"prepare_interp: interpolation matrix + NNLS weights, N={N} centres, d={d}, {args.sd.upper()} {args.kernel} kernel; "
                                       "IM row blocks sharded over ranks, normal equations all-reduced (NCCL), passive-set Cholesky replicated",
                           "l2_policy": "IM, M and the gathered passive-set copy are 3 x N^2 x 8 B, far beyond the 126 MB L2 at N = 16384"},
                "prepare_kernel_s": t_prepare_kernel, "rnorm": rnorm, "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline,
                "e2e": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def run_apes_e2e(args):
    """--workload apes_e2e: whole APES iterations through the host API for the other BASELINE.json chains, on one GPU or, under torchrun, in
    multi-rank (SPMD) mode with the density work sharded over the ranks (e.g. --target funnel --walkers 32768 --dim 30 --gpus 8):
    configs[3] (VKDE Student-t on the 30-D Neal funnel, 32768 walkers: --target funnel --walkers 32768 --dim 30 --kernel cauchy) and
    configs[0] (example_apes.py: 2-D Rosenbrock, 400 walkers, ST kernel: --target rosenbrock --walkers 400 --dim 2 --kernel st3)."""
    import torch

    from numcosmo_b200 import capi
    from numcosmo_b200 import stats_dist as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; numcosmo_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    lib = S.lib()
    lib.ncm_b200_set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
        lib.ncm_b200_set_num_threads(max(1, (os.cpu_count() or 1) // world))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    W, d = args.walkers, args.dim
    N = W // 2
    ktn, okind, nu = KT[args.kernel]
    rs = np.random.default_rng(4)
    if args.target == "funnel":
        # ncm_data_funnel.c:113-132; init nu ~ N(0, 3^2), x_i ~ N(0, e^nu); numcosmo_py/experiments/funnel.py:52 over_smooth 0.2
        nuv = rs.normal(0.0, 3.0, size=W)
        X = np.empty((W, d))
        X[:, 0] = nuv
        X[:, 1:] = rs.normal(size=(W, d - 1)) * np.exp(0.5 * nuv)[:, None]
        lb = np.concatenate([[-1000.0], np.full(d - 1, -1.0e6)])
        ub = np.concatenate([[1000.0], np.full(d - 1, 1.0e6)])
        m2lnL = (d - 1) * nuv + (nuv / 3.0) ** 2 + np.sum(X[:, 1:] ** 2, axis=1) * np.exp(-nuv)
        os_ = 0.2 if args.over_smooth is None else args.over_smooth
        target = S.TARGET_FUNNEL
    elif args.target == "mvnd":
        # the configs[1] recipe at any size (e.g. --walkers 65536: the largest ensemble the north star names)
        mu_t, cov_t, U_t, X, m2lnL = make_problem_b200(S, W, d)
        lb, ub = np.full(d, -50.0), np.full(d, 50.0)
        os_ = 1.0 if args.over_smooth is None else args.over_smooth
        target = S.TARGET_MVND
    else:
        # ncm_model_rosenbrock.c:139-142 bounds; numcosmo_py/experiments/rosenbrock.py:46-49 init
        X = np.array([1.0, 1.0])[None, :] + 1.0e2 * rs.normal(size=(W, 2)) * 0.01
        lb, ub = np.array([-200.0, -400.0]), np.array([200.0, 800.0])
        m2lnL = 0.1 * (100.0 * (X[:, 1] - X[:, 0] ** 2) ** 2 + (1.0 - X[:, 0]) ** 2)
        os_ = 1.1 if args.over_smooth is None else args.over_smooth
        target = S.TARGET_ROSENBROCK
    X = np.ascontiguousarray(X)
    apes = S.FitESMCMCWalkerAPES(W, d, S.FitESMCMCWalkerAPESMethod.VKDE, getattr(S.FitESMCMCWalkerAPESKType, ktn), os_, True)
    apes.set_use_threads(True)
    if world > 1:
        apes.comm_init_from_torch()
    theta, ml = X.copy(), np.ascontiguousarray(m2lnL)
    rng = S.RNG(4)
    warm = max(1, min(args.warmup, 2))
    targs = (mu_t, U_t) if args.target == "mvnd" else None
    apes.run(target, lb, ub, theta, ml, warm, rng, target_args=targs, record_accept=False)
    sds = apes.peek_sds()
    ctxs = [capi.Context.borrowed(lib.ncm_stats_dist_b200_peek_ctx(sd._h)) for sd in sds]
    for c in ctxs:
        c.reset_timers()
    torch.cuda.synchronize()
    steps = max(1, args.steps)
    if world > 1:
        dist.barrier()
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        acc, _ = apes.run(target, lb, ub, theta, ml, steps, rng, target_args=targs, record_accept=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    traffic = [c.get_traffic() for c in ctxs]
    launches = sum(c.get_timers()[1] for c in ctxs)
    apes.enable_timers(True)
    _, stage = apes.run(target, lb, ub, theta, ml, 1, rng, target_args=targs, record_accept=False)
    apes.enable_timers(False)
    nn = [sd.nnls_stats() for sd in sds]
    uses = [c.vkde_path() for c in ctxs]
    pairs_step = 6.0 * N * N
    line = {"metric": METRIC, "value": pairs_step / dt, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"APES iteration end to end (host API, host buffers), VKDE {args.kernel} kernel, {d}-D {args.target}, {W} walkers, "
                                   f"over_smooth {os_} (6 N^2 pairs/step, N={N})" +
                                   (f"; ONE ensemble over {world} GPUs in multi-rank mode (IM rows / query rows sharded, NCCL all-reduce + all-gather)" if world > 1 else ""),
                       "multi_gpu": "sharded" if world > 1 else "single"},
            "accept_hash": int(np.packbits(acc).astype(np.uint64).sum()),
            "walker_steps_per_s": W / dt, "accept_rate": float(np.mean(acc)),
            "e2e": {"value": pairs_step / dt, "unit": UNIT, "h2d_bytes_per_step": sum(t[0] for t in traffic) / steps,
                    "d2h_bytes_per_step": sum(t[1] for t in traffic) / steps, "stage_ms_per_step": {k: round(v, 3) for k, v in stage.items()}},
            "vkde_tensor_core_path": [bool(u[0]) for u in uses], "vkde_max_cond": [u[1] for u in uses], "nnls": nn,
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": None, "cpu_baseline": None}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
def run_cv(args):
    """--workload cv: one prepare_interp under a cross-validation mode through the host API on ONE GPU, host
    buffers, wall clock.  CV_SPLIT is what the reference's own callers use (tests/c/ncm/fit/test_ncm_fit_esmcmc.c:711, tools/mcat_analyze.c:422):
    11 IM + NNLS passes, then a one-parameter levmar fit whose every residual evaluation is IM + NNLS + a batched eval of all observations."""
    import torch

    from numcosmo_b200 import stats_dist as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; numcosmo_b200 has no CPU fallback")
    lib = S.lib()
    lib.ncm_b200_set_device(int(os.environ.get("LOCAL_RANK", "0")))
    d, n_obs = args.dim, args.sweep_n
    ktn, okind, nu = KT[args.kernel]
    cv_name = args.cv.upper()

    def problem(n, seed=2):
        rs = np.random.default_rng(seed)
        sig = rs.uniform(2e-2, 5e-2, size=d)
        R = rs.normal(size=(d, d)) / np.sqrt(d)
        cov = (0.7 * np.eye(d) + 0.3 * (R @ R.T)) * np.outer(sig, sig)
        z = rs.normal(size=(n, d))
        return np.ascontiguousarray(rs.uniform(1.0, 2.0, size=d) + z @ np.linalg.cholesky(cov).T), np.einsum("ij,ij->i", z, z)

    X, m2lnp = problem(n_obs)
    kern = S.StatsDistKernelGauss(d) if okind == 0 else S.StatsDistKernelST(d, nu)
    sd = (S.StatsDistVKDE if args.sd == "vkde" else S.StatsDistKDE)(kern, getattr(S.StatsDistCV, cv_name))
    sd.set_use_threads(True)
    for x in X:
        sd.add_obs(x)
    steps, times, launches = max(1, min(args.steps, 3)), [], 0
    with ClockSampler(0) as clk:
        for it in range(1 + steps):                          # one warm-up call (allocations, first-launch costs)
            sd.set_over_smooth(1.0)
            sd.enable_timers(True) if it == steps else None
            t0 = time.perf_counter()
            sd.prepare_interp(m2lnp)
            times.append(time.perf_counter() - t0)
    tm, launches = sd.get_timers()
    ms = 1e3 * float(np.mean(times[1:]))
    lnos, val = sd.cv_trace()
    nk = sd.get_n_kernels()
    n_eval = len(lnos)
    per_eval = {"split": float(n_obs) * nk, "split_nofit": float(n_obs - nk) * nk, "loo": float(nk) * nk}[args.cv]
    if args.cv == "split":                                   # the levmar evaluations add a batched eval of all observations to IM + NNLS
        pairs = 11 * per_eval + (n_eval - 11 + 1) * 2 * per_eval
    elif args.cv == "loo" and not (args.sd == "kde" and okind == 0):
        pairs = n_eval * per_eval
    elif args.cv == "loo":
        pairs = 2 * n_eval * per_eval
    else:
        pairs = n_eval * per_eval + float(n_obs) * nk
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import ncm_oracle as O

        n_cpu = min(n_obs, 2048)
        Xc, mc = problem(n_cpu)
        o = O.StatsDist(O.SD_VKDE if args.sd == "vkde" else O.SD_KDE, okind, d, nu, getattr(O, "CV_" + cv_name))
        o.set_use_threads(True)
        o.add_obs_matrix(Xc)
        O.lib().orc_set_blas_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        assert o.prepare_interp(mc) == 0
        t_cpu = time.perf_counter() - t0
        cpu = {"value": 1.0 / t_cpu, "unit": "prepare_interp/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"the same CV_{cv_name} prepare_interp on {n_cpu} observations (oracle port, all host cores), {len(o.cv_trace()[0])} objective "
                         f"evaluations in {t_cpu:.2f} s -- NOT the same size as the GPU line unless n_obs <= 2048", "n_obs": n_cpu, "s": t_cpu}
    line = {"metric": METRIC, "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": 1, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"prepare_interp with NCM_STATS_DIST_CV_{cv_name}: {args.sd.upper()} {args.kernel} kernel, {n_obs} observations, d={d}, "
                                   f"{nk} kernels; end to end through the host API (host buffers, wall clock)",
                       "l2_policy": "every objective evaluation rebuilds the IM (n_obs x n_kernels x 8 B) and the normal equations at a new bandwidth"},
            "objective_evaluations": n_eval, "over_smooth": sd.get_over_smooth(), "rnorm": sd.get_rnorm(),
            "stage_ms_last_call": {k: round(v, 3) for k, v in tm.items()}, "gpu_launches": int(launches), "clocks": clk.summary(),
            "e2e": {"value": pairs / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "ms_per_step": ms},
            "roofline": None, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    # torchrun exports OMP_NUM_THREADS=1; the host-side OpenMP loops (oracle port for the reference arm, prepare_kernel pieces and the
    # per-walker transition terms for the B200 arm) read it when their library is first loaded, so it is set here, before any of them is
    ncores = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(ncores if args.impl == "reference" else max(1, ncores // world))
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "eval_sweep":
        run_sweep(args)
    elif args.workload == "apes_e2e":
        run_apes_e2e(args)
    elif args.workload == "cv":
        if args.cv == "loo" and args.sweep_n == 65536 and args.dim == 10:
            # outside KDE-Gauss the CV_LOO objective is a Monte-Carlo integral whose draw count grows like var(p) / mean(p)^2: minutes at d = 10
            # on either side (ncm_stats_dist.c:606-640); the default LOO line is therefore a small low-dimensional fit
            args.sweep_n, args.dim = 2048, 3
        if args.sweep_n == 65536:
            args.sweep_n = 8192
        run_cv(args)
    elif args.workload == "prepare_interp":
        if args.sweep_n == 65536 and args.dim == 10:   # the defaults of the other workloads: use configs[2] sizes
            args.sweep_n, args.dim = 16384, 20
        run_prepare_interp(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
