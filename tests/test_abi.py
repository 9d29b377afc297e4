"""The drop-in boundary without a GPU: both shared libraries load, export every function their public headers declare
(include/ncm_sd_gpu.h, include/ncm_stats_dist_b200.h), and the product path refuses to run without an sm_100 device
(no CPU fallback) instead of computing anything on the host."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, txt)))


def test_c_abi_exports_every_declared_symbol():
    from numcosmo_b200 import capi

    lib = C.CDLL(capi.LIB_PATH)
    names = _declared("ncm_sd_gpu.h", "ncm_sd_gpu_")
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # the Python binding table names the same entry points
    assert set(capi.SYMBOLS) <= set(names) | {"ncm_sd_gpu_stream", "ncm_sd_gpu_synchronize"}, sorted(set(capi.SYMBOLS) - set(names))


def test_host_mirror_exports_every_declared_symbol():
    from numcosmo_b200 import stats_dist as S

    lib = S.lib()
    names = [n for p in ("ncm_stats_dist_", "ncm_fit_esmcmc_walker_apes_", "ncm_b200_", "ncm_rng_", "ncm_vector_", "ncm_matrix_")
             for n in _declared("ncm_stats_dist_b200.h", p)]
    assert len(names) >= 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    """Here (no GPU) ctx_new must answer ENODEV and the host mirror must raise; on a GPU box this test is skipped."""
    from numcosmo_b200 import capi
    from numcosmo_b200 import stats_dist as S

    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.GpuError) as e:
        capi.Context(0)
    assert e.value.args[0] == capi.ENODEV or "ENODEV" in str(e.value) or "no usable" in str(e.value)
    sd = S.StatsDistKDE(S.StatsDistKernelGauss(2), S.StatsDistCV.NONE)
    rs = np.random.default_rng(0)
    for x in rs.normal(size=(20, 2)):
        sd.add_obs(x)
    with pytest.raises(Exception) as e2:
        sd.prepare()
    assert "no CPU fallback" in str(e2.value) or "CUDA device" in str(e2.value)


def test_product_code_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "numcosmo_b200")):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt) and re.search(r"import\s+oracle|from\s+oracle|ncm_oracle|libncm_oracle|orc_", txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
