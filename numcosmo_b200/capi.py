"""ctypes binding of libncm_sd_gpu.so (include/ncm_sd_gpu.h).

This is the only way the Python side of numcosmo_b200 reaches the CUDA kernels.  The library is
built in-tree (numcosmo_b200/lib/) by ``make -C numcosmo_b200/csrc`` / ``__graft_entry__.build()``;
there is no fallback: a missing library or a missing sm_100 device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libncm_sd_gpu.so")

KERNEL_GAUSS, KERNEL_ST = 0, 1
KDE, VKDE = 0, 1
T_NAMES = ("eval", "IM", "syrk", "chol", "nnls_misc", "h2d", "d2h", "prep", "lowrank", "comm")

OK, EINVAL, ENODEV, ECUDA, ENOTPD, ENCCL, ENOMEM = range(7)

# every symbol include/ncm_sd_gpu.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = (
    "ncm_sd_gpu_ctx_new", "ncm_sd_gpu_ctx_free", "ncm_sd_gpu_last_error", "ncm_sd_gpu_device_count", "ncm_sd_gpu_stream",
    "ncm_sd_gpu_synchronize", "ncm_sd_gpu_set_kernel", "ncm_sd_gpu_upload_kde", "ncm_sd_gpu_upload_vkde", "ncm_sd_gpu_set_weights",
    "ncm_sd_gpu_set_href", "ncm_sd_gpu_get_weights", "ncm_sd_gpu_eval_m2lnp", "ncm_sd_gpu_eval", "ncm_sd_gpu_eval_m2lnp_dev",
    "ncm_sd_gpu_compute_IM", "ncm_sd_gpu_nnls_solve", "ncm_sd_gpu_nnls_solve_host", "ncm_sd_gpu_sample_apply", "ncm_sd_gpu_sample_philox",
    "ncm_sd_gpu_comm_unique_id", "ncm_sd_gpu_comm_init", "ncm_sd_gpu_set_row_shard", "ncm_sd_gpu_set_auto_shard", "ncm_sd_gpu_allgather_dev", "ncm_sd_gpu_get_timers", "ncm_sd_gpu_reset_timers",
    "ncm_sd_gpu_enable_timers", "ncm_sd_gpu_get_traffic", "ncm_sd_gpu_dsyrk_ata_dev", "ncm_sd_gpu_dpotrf_upper_dev",
    "ncm_sd_gpu_vkde_prepare", "ncm_sd_gpu_vkde_finish", "ncm_sd_gpu_dposv_upper_dev", "ncm_sd_gpu_dtrtri_upper_dev", "ncm_sd_gpu_dsysv_upper_dev", "ncm_sd_gpu_dgels_cols_dev", "ncm_sd_gpu_vkde_path", "ncm_sd_gpu_host_alloc", "ncm_sd_gpu_host_free",
)


class GpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ncm_sd_gpu error {code}: {msg}")
        self.code = code


class NNLSStats(C.Structure):
    _fields_ = [("n_chol", C.c_int), ("n_lu", C.c_int), ("n_outer", C.c_int), ("n_passive", C.c_int), ("chol_flops", C.c_double),
                ("syrk_flops", C.c_double), ("n_lowrank", C.c_int), ("n_lowrank_fallback", C.c_int), ("n_trinv", C.c_int), ("max_lowrank_k", C.c_int),
                ("lowrank_flops", C.c_double), ("n_dist_chol", C.c_int), ("n_qr", C.c_int), ("n_lowrank_nested", C.c_int), ("reserved_", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def load():
    """Load the CUDA library; raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C numcosmo_b200/csrc` (or __graft_entry__.build()). "
                "numcosmo_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
        ull = C.c_ulonglong
        sig = {
            "ncm_sd_gpu_ctx_new": (i, [C.POINTER(vp), i]),
            "ncm_sd_gpu_ctx_free": (i, [vp]),
            "ncm_sd_gpu_last_error": (C.c_char_p, [vp]),
            "ncm_sd_gpu_device_count": (i, []),
            "ncm_sd_gpu_stream": (vp, [vp]),
            "ncm_sd_gpu_synchronize": (i, [vp]),
            "ncm_sd_gpu_set_kernel": (i, [vp, i, d, i]),
            "ncm_sd_gpu_upload_kde": (i, [vp, i, i, _dp, i, _dp, i, d]),
            "ncm_sd_gpu_upload_vkde": (i, [vp, i, i, _dp, i, _dp, _dp]),
            "ncm_sd_gpu_vkde_prepare": (i, [vp, i, i, _dp, i, _dp, i, i, _dp, _ip]),
            "ncm_sd_gpu_vkde_finish": (i, [vp, _dp, i, _ip, _dp]),
            "ncm_sd_gpu_set_weights": (i, [vp, i, _dp, d]),
            "ncm_sd_gpu_set_href": (i, [vp, d]),
            "ncm_sd_gpu_get_weights": (i, [vp, i, _dp]),
            "ncm_sd_gpu_eval_m2lnp": (i, [vp, i, _dp, i, _dp]),
            "ncm_sd_gpu_eval": (i, [vp, i, _dp, i, _dp]),
            "ncm_sd_gpu_eval_m2lnp_dev": (i, [vp, i, vp, i, vp]),
            "ncm_sd_gpu_compute_IM": (i, [vp, _dp, _dp]),
            "ncm_sd_gpu_nnls_solve": (i, [vp, d, _dp, _dp, C.POINTER(NNLSStats)]),
            "ncm_sd_gpu_nnls_solve_host": (i, [vp, i, i, _dp, i, _dp, d, _dp, _dp, C.POINTER(NNLSStats)]),
            "ncm_sd_gpu_sample_apply": (i, [vp, i, _ip, _dp, i, _dp, _dp, i]),
            "ncm_sd_gpu_sample_philox": (i, [vp, i, ull, ull, _dp, i, _ip]),
            "ncm_sd_gpu_comm_unique_id": (i, [C.c_char_p]),
            "ncm_sd_gpu_comm_init": (i, [vp, i, i, C.c_char_p]),
            "ncm_sd_gpu_set_row_shard": (i, [vp, i, i]),
            "ncm_sd_gpu_set_auto_shard": (i, [vp, i]),
            "ncm_sd_gpu_allgather_dev": (i, [vp, vp, vp, i]),
            "ncm_sd_gpu_get_timers": (i, [vp, _dp, C.POINTER(ll)]),
            "ncm_sd_gpu_reset_timers": (i, [vp]),
            "ncm_sd_gpu_enable_timers": (i, [vp, i]),
            "ncm_sd_gpu_get_traffic": (i, [vp, C.POINTER(ll), C.POINTER(ll)]),
            "ncm_sd_gpu_dsyrk_ata_dev": (i, [vp, i, i, vp, i, vp, i]),
            "ncm_sd_gpu_dpotrf_upper_dev": (i, [vp, i, vp, i, _ip]),
            "ncm_sd_gpu_dposv_upper_dev": (i, [vp, i, vp, i, vp, _ip]),
            "ncm_sd_gpu_vkde_path": (i, [vp, _ip, _dp]),
            "ncm_sd_gpu_dtrtri_upper_dev": (i, [vp, i, vp, i, vp, vp]),
            "ncm_sd_gpu_dsysv_upper_dev": (i, [vp, i, vp, i, vp, _ip]),
            "ncm_sd_gpu_dgels_cols_dev": (i, [vp, i, i, vp, i, vp, vp, vp, _ip]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "float64 C-contiguous array required"
    return a.ctypes.data_as(_dp)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One ncm_sd_gpu_ctx (one process, one device)."""

    def __init__(self, device: int = 0):
        L = load()
        h = C.c_void_p()
        rc = L.ncm_sd_gpu_ctx_new(C.byref(h), device)
        if rc != OK:
            raise GpuError(rc, "no usable sm_100 CUDA device" if rc == ENODEV else "context creation failed")
        self._h = h
        self.device = device
        self.d = 0
        self.n_kernels = 0
        self.n_obs = 0

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_own", True):
                load().ncm_sd_gpu_ctx_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != OK:
            raise GpuError(rc, load().ncm_sd_gpu_last_error(self._h).decode())

    @property
    def stream(self) -> int:
        return load().ncm_sd_gpu_stream(self._h)

    def synchronize(self):
        self._ck(load().ncm_sd_gpu_synchronize(self._h))

    def set_kernel(self, kind: int, nu: float, d: int):
        self._ck(load().ncm_sd_gpu_set_kernel(self._h, kind, float(nu), d))
        self.d = d

    def upload_kde(self, invUsample, n_kernels: int, U, lnnorm: float):
        invUsample, U = _f64(invUsample), _f64(U)
        self._ck(load().ncm_sd_gpu_upload_kde(self._h, invUsample.shape[0], n_kernels, _p(invUsample), invUsample.shape[1], _p(U), U.shape[1], float(lnnorm)))
        self.n_obs, self.n_kernels = invUsample.shape[0], n_kernels

    def upload_vkde(self, sample, n_kernels: int, U_all, lnnorms):
        sample, U_all, lnnorms = _f64(sample), _f64(U_all), _f64(lnnorms)
        assert U_all.shape == (n_kernels, self.d, self.d) and lnnorms.shape == (n_kernels,)
        self._ck(load().ncm_sd_gpu_upload_vkde(self._h, sample.shape[0], n_kernels, _p(sample), sample.shape[1], _p(U_all), _p(lnnorms)))
        self.n_obs, self.n_kernels = sample.shape[0], n_kernels

    def vkde_prepare(self, sample, invUsample, n_kernels: int, k: int):
        """kNN + local covariance + Cholesky on the device; returns (U_all [n x d x d], fail [n])."""
        sample, invUsample = _f64(sample), _f64(invUsample)
        U = np.empty((n_kernels, self.d, self.d))
        fail = np.empty(n_kernels, dtype=np.int32)
        self._ck(load().ncm_sd_gpu_vkde_prepare(self._h, sample.shape[0], n_kernels, _p(sample), sample.shape[1], _p(invUsample), invUsample.shape[1],
                                                 k, _p(U), fail.ctypes.data_as(_ip)))
        self.n_obs, self.n_kernels = sample.shape[0], n_kernels
        return U, fail

    def vkde_finish(self, lnnorms, fixed_idx=None, fixed_U=None):
        lnnorms = _f64(lnnorms)
        nf = 0 if fixed_idx is None else len(fixed_idx)
        fi = np.ascontiguousarray(fixed_idx, dtype=np.int32) if nf else None
        fu = _f64(fixed_U) if nf else None
        self._ck(load().ncm_sd_gpu_vkde_finish(self._h, _p(lnnorms), nf, fi.ctypes.data_as(_ip) if nf else None, _p(fu) if nf else None))

    def set_weights(self, weights, href: float):
        weights = _f64(weights)
        self._ck(load().ncm_sd_gpu_set_weights(self._h, weights.size, _p(weights), float(href)))

    def set_href(self, href: float):
        self._ck(load().ncm_sd_gpu_set_href(self._h, float(href)))

    def get_weights(self) -> np.ndarray:
        w = np.zeros(self.n_kernels)
        self._ck(load().ncm_sd_gpu_get_weights(self._h, self.n_kernels, _p(w)))
        return w

    def eval_m2lnp(self, X, out=None) -> np.ndarray:
        X = _f64(X)
        if X.ndim == 1:
            X = X[None, :]
        if out is None:
            out = np.empty(X.shape[0])
        self._ck(load().ncm_sd_gpu_eval_m2lnp(self._h, X.shape[0], _p(X), X.shape[1], _p(out)))
        return out

    def eval(self, X) -> np.ndarray:
        X = _f64(X)
        if X.ndim == 1:
            X = X[None, :]
        out = np.empty(X.shape[0])
        self._ck(load().ncm_sd_gpu_eval(self._h, X.shape[0], _p(X), X.shape[1], _p(out)))
        return out

    def eval_m2lnp_dev(self, q: int, dX_ptr: int, ldx: int, dOut_ptr: int):
        self._ck(load().ncm_sd_gpu_eval_m2lnp_dev(self._h, q, dX_ptr, ldx, dOut_ptr))

    def set_auto_shard(self, on: bool = True):
        self._ck(load().ncm_sd_gpu_set_auto_shard(self._h, int(on)))

    def allgather_dev(self, dsend_ptr: int, drecv_ptr: int, count: int):
        self._ck(load().ncm_sd_gpu_allgather_dev(self._h, dsend_ptr, drecv_ptr, count))

    def set_row_shard(self, row0: int, nrows: int):
        self._ck(load().ncm_sd_gpu_set_row_shard(self._h, row0, nrows))
        self._nrows = nrows

    def compute_IM(self, row_scale=None, fetch: bool = False, nrows=None):
        rs = _f64(row_scale) if row_scale is not None else None
        nrows = nrows if nrows is not None else getattr(self, "_nrows", self.n_obs)
        IM = np.empty((nrows, self.n_kernels)) if fetch else None
        self._ck(load().ncm_sd_gpu_compute_IM(self._h, _p(rs) if rs is not None else None, _p(IM) if fetch else None))
        return IM

    def nnls_solve(self, reltol: float = np.finfo(float).eps):
        x = np.zeros(self.n_kernels)
        rnorm = C.c_double()
        st = NNLSStats()
        self._ck(load().ncm_sd_gpu_nnls_solve(self._h, reltol, _p(x), C.byref(rnorm), C.byref(st)))
        return x, rnorm.value, st.as_dict()

    def nnls_solve_host(self, A, f, reltol: float = np.finfo(float).eps):
        A, f = _f64(A), _f64(f)
        x = np.zeros(A.shape[1])
        rnorm = C.c_double()
        st = NNLSStats()
        self._ck(load().ncm_sd_gpu_nnls_solve_host(self._h, A.shape[0], A.shape[1], _p(A), A.shape[1], _p(f), reltol, _p(x), C.byref(rnorm), C.byref(st)))
        return x, rnorm.value, st.as_dict()

    def sample_apply(self, kidx, Z, scale=None) -> np.ndarray:
        kidx = np.ascontiguousarray(kidx, dtype=np.int32)
        Z = _f64(Z)
        sc = _f64(scale) if scale is not None else None
        X = np.empty_like(Z)
        self._ck(load().ncm_sd_gpu_sample_apply(self._h, Z.shape[0], kidx.ctypes.data_as(_ip), _p(Z), Z.shape[1], _p(sc) if sc is not None else None, _p(X), X.shape[1]))
        return X

    def sample_philox(self, q: int, seed: int, offset: int = 0):
        X = np.empty((q, self.d))
        kidx = np.empty(q, dtype=np.int32)
        self._ck(load().ncm_sd_gpu_sample_philox(self._h, q, seed, offset, _p(X), self.d, kidx.ctypes.data_as(_ip)))
        return X, kidx

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._ck(load().ncm_sd_gpu_comm_init(self._h, nranks, rank, unique_id))

    def enable_timers(self, on: bool = True):
        self._ck(load().ncm_sd_gpu_enable_timers(self._h, int(on)))

    def reset_timers(self):
        self._ck(load().ncm_sd_gpu_reset_timers(self._h))

    def get_timers(self):
        ms = np.zeros(len(T_NAMES))
        n = C.c_longlong()
        self._ck(load().ncm_sd_gpu_get_timers(self._h, _p(ms), C.byref(n)))
        return dict(zip(T_NAMES, ms.tolist())), n.value

    def get_traffic(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(load().ncm_sd_gpu_get_traffic(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @classmethod
    def borrowed(cls, handle):
        """Wrap a context owned by a host-mirror object (ncm_stats_dist_b200_peek_ctx)."""
        self = cls.__new__(cls)
        self._h = C.c_void_p(handle)
        self._own = False
        self.device = -1
        self.d = self.n_kernels = self.n_obs = 0
        return self

    def dsyrk_ata_dev(self, nrows, ncols, dA_ptr, lda, dM_ptr, ldm):
        self._ck(load().ncm_sd_gpu_dsyrk_ata_dev(self._h, nrows, ncols, dA_ptr, lda, dM_ptr, ldm))

    def vkde_path(self):
        """(uses_mma, max cond_1 of the factors) of the last VKDE upload."""
        m, cnd = C.c_int(), C.c_double()
        self._ck(load().ncm_sd_gpu_vkde_path(self._h, C.byref(m), C.byref(cnd)))
        return bool(m.value), cnd.value

    def dposv_upper_dev(self, n, dM_ptr, ldm, dRhs_ptr) -> int:
        info = C.c_int()
        self._ck(load().ncm_sd_gpu_dposv_upper_dev(self._h, n, dM_ptr, ldm, dRhs_ptr, C.byref(info)))
        return info.value

    def dsysv_upper_dev(self, n, dM_ptr, ldm, dRhs_ptr) -> int:
        info = C.c_int()
        self._ck(load().ncm_sd_gpu_dsysv_upper_dev(self._h, n, dM_ptr, ldm, dRhs_ptr, C.byref(info)))
        return info.value

    def dgels_cols_dev(self, m, n, dA_ptr, lda, dIdx_ptr, dF_ptr, dX_ptr) -> int:
        info = C.c_int()
        self._ck(load().ncm_sd_gpu_dgels_cols_dev(self._h, m, n, dA_ptr, lda, dIdx_ptr, dF_ptr, dX_ptr, C.byref(info)))
        return info.value

    def dtrtri_upper_dev(self, n, dU_ptr, ld, dW_ptr, dScratch_ptr):
        self._ck(load().ncm_sd_gpu_dtrtri_upper_dev(self._h, n, dU_ptr, ld, dW_ptr, dScratch_ptr))

    def dpotrf_upper_dev(self, n, dM_ptr, ldm) -> int:
        info = C.c_int()
        self._ck(load().ncm_sd_gpu_dpotrf_upper_dev(self._h, n, dM_ptr, ldm, C.byref(info)))
        return info.value


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().ncm_sd_gpu_comm_unique_id(buf)
    if rc != OK:
        raise GpuError(rc, "ncclGetUniqueId failed")
    return buf.raw


def device_count() -> int:
    return load().ncm_sd_gpu_device_count()
