"""bench.py contract checks that need no GPU: both arms describe the same workload (one `config` for the b200 arm and the
`--impl reference` arm at every N), defaults finish within minutes, the N > 1 default is ONE sharded ensemble (strong scaling)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _parse(monkeypatch, argv, world=None):
    import bench

    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    if world is None:
        monkeypatch.delenv("WORLD_SIZE", raising=False)
    else:
        monkeypatch.setenv("WORLD_SIZE", str(world))
    return bench, bench.parse()


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_both_arms_print_the_same_config(monkeypatch, n):
    bench, a = _parse(monkeypatch, ["--gpus", str(n)], world=n)
    bench_r, r = _parse(monkeypatch, ["--gpus", str(n), "--impl", "reference"], world=n)
    assert (a.walkers, a.dim) == (r.walkers, r.dim)
    cfg_a, sc_a = bench.apes_config(a, a.walkers, a.dim, a.walkers // 2, n)
    cfg_r, sc_r = bench.apes_config(r, r.walkers, r.dim, r.walkers // 2, r.gpus)
    assert cfg_a == cfg_r and sc_a == sc_r
    if n == 1:
        assert (a.walkers, a.dim) == (4096, 10) and sc_a == "weak" and cfg_a["multi_gpu"] == "single"      # BASELINE configs[1]
    else:
        assert (a.walkers, a.dim) == (32768, 20) and sc_a == "strong" and cfg_a["multi_gpu"] == "sharded"    # one ensemble over N GPUs


def test_reference_arm_resolves_the_multi_gpu_workload_without_torchrun(monkeypatch):
    """`python bench.py --impl reference --gpus 8` launched plainly (no WORLD_SIZE) still describes the 8-GPU arm's workload."""
    _, r = _parse(monkeypatch, ["--gpus", "8", "--impl", "reference"], world=None)
    assert (r.walkers, r.dim) == (32768, 20)


def test_defaults(monkeypatch):
    _, a = _parse(monkeypatch, [], world=None)
    assert a.gpus == 1 and a.steps == 10 and a.warmup >= 3 and a.impl == "b200"


def test_replicas_mode_is_weak_scaling(monkeypatch):
    bench, a = _parse(monkeypatch, ["--gpus", "4", "--apes-multi", "replicas"], world=4)
    cfg, sc = bench.apes_config(a, a.walkers, a.dim, a.walkers // 2, 4)
    assert sc == "weak" and cfg["multi_gpu"] == "replicas" and (a.walkers, a.dim) == (4096, 10)
