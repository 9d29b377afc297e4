"""ctypes binding of the CPU oracle (oracle/_build/libncm_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by numcosmo_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libncm_oracle.so")
_REF_LEVMAR_PATH = os.path.join(_HERE, "_ref", "liblevmar_ref.so")   # the reference's own levmar, built by `make ref`

KERNEL_GAUSS, KERNEL_ST = 0, 1
SD_KDE, SD_VKDE = 0, 1
CV_NONE, CV_SPLIT, CV_SPLIT_NOFIT, CV_LOO = 0, 1, 2, 3
COV_SAMPLE, COV_FIXED, COV_ROBUST_DIAG, COV_ROBUST = 0, 1, 2, 3
TARGET_MVND, TARGET_ROSENBROCK, TARGET_FUNNEL = 0, 1, 2

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc + SciPy's OpenBLAS)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_FMIN_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_int, C.c_void_p)
_LM_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)
_ref_levmar = None


def ref_levmar():
    """The reference's own levmar (oracle/_ref/liblevmar_ref.so), or None when it was never built."""
    global _ref_levmar
    if _ref_levmar is None and os.path.exists(_REF_LEVMAR_PATH):
        L = C.CDLL(_REF_LEVMAR_PATH)
        L.dlevmar_dif.restype = C.c_int
        L.dlevmar_dif.argtypes = [_LM_FN, _dp, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, C.c_void_p]
        _ref_levmar = L
    return _ref_levmar


_REF_KDTREE_PATH = os.path.join(_HERE, "_ref", "libkdtree_ref.so")    # the reference's own kd-tree, built by `make ref`
_ref_kdtree = None


def ref_kdtree_knn(points, queries, k):
    """k nearest neighbours of points[queries] by the REFERENCE'S kd-tree, in its list order; None when it was never built."""
    global _ref_kdtree
    if _ref_kdtree is None:
        if not os.path.exists(_REF_KDTREE_PATH):
            return None
        L = C.CDLL(_REF_KDTREE_PATH)
        L.orc_ref_kdtree_knn.restype = C.c_int
        L.orc_ref_kdtree_knn.argtypes = [_dp, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_long), _dp]
        _ref_kdtree = L
    pts = np.ascontiguousarray(points, dtype=np.float64)
    q = np.ascontiguousarray(queries, dtype=np.int32)
    idx = np.zeros((len(q), k), dtype=np.int64)
    dist = np.zeros((len(q), k))
    rc = _ref_kdtree.orc_ref_kdtree_knn(_p(pts), pts.shape[0], pts.shape[1], q.ctypes.data_as(C.POINTER(C.c_int)), len(q), k,
                                        idx.ctypes.data_as(C.POINTER(C.c_long)), _p(dist))
    if rc != 0:
        raise RuntimeError(f"reference kd-tree returned a list of the wrong length for query {-rc - 1}")
    return idx, dist


_REF_NNLS_PATH = os.path.join(_HERE, "_ref", "libnnls_ref.so")        # the reference's vendored Lawson-Hanson NNLS, built by `make ref`
_ref_nnls = None


def ref_nnls_lh(A, f):
    """min ||A x - f||, x >= 0 by the REFERENCE'S OWN Lawson-Hanson code (numcosmo/external/misc/nnls.c: nnls_c, the solver behind
    ncm_nnls.c:873-935); returns (x, rnorm, mode) or None when the library was never built.  Column-major (f2c) inside."""
    global _ref_nnls
    if _ref_nnls is None:
        if not os.path.exists(_REF_NNLS_PATH):
            return None
        L = C.CDLL(_REF_NNLS_PATH)
        ip = C.POINTER(C.c_int)
        L.nnls_c.restype = C.c_int
        L.nnls_c.argtypes = [_dp, ip, ip, ip, _dp, _dp, _dp, _dp, _dp, ip, ip]
        _ref_nnls = L
    A = np.asarray(A, dtype=np.float64)
    m, n = A.shape
    a = np.asfortranarray(A).copy(order="F")
    b = np.ascontiguousarray(f, dtype=np.float64).copy()
    x, w, zz = np.zeros(n), np.zeros(n), np.zeros(m)
    index = np.zeros(n, dtype=np.int32)
    rnorm, mode = C.c_double(), C.c_int()
    mi, ni = C.c_int(m), C.c_int(n)
    _ref_nnls.nnls_c(a.ctypes.data_as(_dp), C.byref(mi), C.byref(mi), C.byref(ni), _p(b), _p(x), C.byref(rnorm), _p(w), _p(zz),
                     index.ctypes.data_as(C.POINTER(C.c_int)), C.byref(mode))
    return x, rnorm.value, mode.value


def knn_brute(points, query, k):
    """The oracle's (squared distance, index)-ordered exact kNN (restatement used by the VKDE prepare_kernel)."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    idx = np.zeros(k, dtype=np.int64)
    dist = np.zeros(k)
    lib().orc_knn_brute(_p(pts), pts.shape[0], pts.shape[1], int(query), int(k), idx.ctypes.data_as(C.POINTER(C.c_long)), _p(dist))
    return idx, dist


def use_ref_levmar(on: bool = True) -> bool:
    """Route the oracle's CV_SPLIT fit through the reference's dlevmar_dif (True) or the restatement (False)."""
    L = ref_levmar() if on else None
    lib().orc_set_levmar_dif(C.cast(L.dlevmar_dif, C.c_void_p) if L is not None else None)
    return L is not None


class _RNG(C.Structure):
    _fields_ = [("mt", C.c_ulong * 624), ("mti", C.c_int)]


class _NNLSStats(C.Structure):
    _fields_ = [("n_chol", C.c_int), ("n_lu", C.c_int), ("n_qr", C.c_int), ("n_outer", C.c_int), ("n_passive", C.c_int)]


class _Target(C.Structure):
    _fields_ = [("kind", C.c_int), ("d", C.c_int), ("mu", _dp), ("cov_inv_U", _dp), ("lb", _dp), ("ub", _dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        sig = {
            "orc_rng_set": (None, [vp, C.c_ulong]),
            "orc_rng_get": (C.c_ulong, [vp]),
            "orc_rng_uniform": (d, [vp]),
            "orc_rng_uniform_pos": (d, [vp]),
            "orc_ran_flat": (d, [vp, d, d]),
            "orc_ran_gaussian": (d, [vp, d]),
            "orc_ran_ugaussian": (d, [vp]),
            "orc_ran_gaussian_ziggurat": (d, [vp, d]),
            "orc_ran_gamma": (d, [vp, d, d]),
            "orc_ran_chisq": (d, [vp, d]),
            "orc_ran_beta": (d, [vp, d, d]),
            "orc_kernel_get_rot_bandwidth": (d, [vp, d]),
            "orc_cholesky_lndet": (d, [_dp, i, i]),
            "orc_kernel_get_lnnorm": (d, [vp, _dp, i]),
            "orc_kernel_eval_unnorm": (d, [vp, d]),
            "orc_kernel_eval_unnorm_vec": (None, [vp, _dp, i, _dp, i, i]),
            "orc_kernel_eval_sum0_gamma_lambda": (None, [vp, _dp, _dp, _dp, _dp, i, _dp, _dp]),
            "orc_kernel_eval_sum1_gamma_lambda": (None, [vp, _dp, _dp, d, _dp, i, _dp, _dp]),
            "orc_kernel_sample": (None, [vp, _dp, i, d, _dp, _dp, vp]),
            "orc_nnls_solve": (d, [_dp, i, i, i, _dp, _dp, d, vp]),
            "orc_sort_smallest_index": (None, [C.POINTER(C.c_int), i, _dp, i, i]),
            "orc_sort_largest_index": (None, [C.POINTER(C.c_int), i, _dp, i, i]),
            "orc_nmsimplex2_new": (vp, [i]),
            "orc_nmsimplex2_free": (None, [vp]),
            "orc_nmsimplex2_set": (i, [vp, _FMIN_FN, vp, _dp, _dp]),
            "orc_nmsimplex2_iterate": (i, [vp]),
            "orc_nmsimplex2_x": (_dp, [vp]),
            "orc_nmsimplex2_fval": (d, [vp]),
            "orc_nmsimplex2_size": (d, [vp]),
            "orc_nmsimplex2_minimize": (i, [vp, _FMIN_FN, vp, _dp, _dp, d, i]),
            "orc_lm_dif": (i, [_LM_FN, _dp, _dp, i, i, i, _dp, _dp, vp]),
            "orc_set_levmar_dif": (None, [vp]),
            "orc_sd_get_over_smooth": (d, [vp]),
            "orc_sd_get_cv_trace": (i, [vp, _dp, _dp, i]),
            "orc_stats_Qn_from_sorted_data": (d, [_dp, i]),
            "orc_knn_brute": (None, [_dp, i, i, i, i, C.POINTER(C.c_long), _dp]),
            "orc_sd_new": (vp, [i, i, d, i, i]),
            "orc_sd_free": (None, [vp]),
            "orc_sd_set_over_smooth": (None, [vp, d]),
            "orc_sd_set_shrink": (None, [vp, d]),
            "orc_sd_set_split_frac": (None, [vp, d]),
            "orc_sd_set_use_threads": (None, [vp, i]),
            "orc_sd_set_cov_type": (None, [vp, i]),
            "orc_sd_set_cov_fixed": (None, [vp, _dp, i]),
            "orc_sd_set_nearPD_maxiter": (None, [vp, i]),
            "orc_sd_set_local_frac": (None, [vp, d]),
            "orc_sd_set_use_rot_href": (None, [vp, i]),
            "orc_sd_reset": (None, [vp]),
            "orc_sd_add_obs": (None, [vp, _dp]),
            "orc_sd_prepare": (i, [vp]),
            "orc_sd_prepare_interp": (i, [vp, _dp, i]),
            "orc_sd_eval": (d, [vp, _dp]),
            "orc_sd_eval_m2lnp": (d, [vp, _dp]),
            "orc_sd_eval_m2lnp_batch": (None, [vp, _dp, i, i, _dp, i]),
            "orc_sd_eval_batch": (None, [vp, _dp, i, i, _dp, i]),
            "orc_sd_kernel_choose": (i, [vp, vp]),
            "orc_sd_sample": (None, [vp, _dp, vp]),
            "orc_sd_set_weights": (None, [vp, _dp]),
            "orc_sd_get_dim": (i, [vp]),
            "orc_sd_get_sample_size": (i, [vp]),
            "orc_sd_get_n_obs": (i, [vp]),
            "orc_sd_get_n_kernels": (i, [vp]),
            "orc_sd_get_href": (d, [vp]),
            "orc_sd_get_rnorm": (d, [vp]),
            "orc_sd_get_lnnorm": (d, [vp, i]),
            "orc_sd_peek_weights": (_dp, [vp]),
            "orc_sd_peek_cov_decomp": (_dp, [vp, i]),
            "orc_sd_peek_full_cov": (_dp, [vp]),
            "orc_sd_peek_full_cov_decomp": (_dp, [vp]),
            "orc_sd_peek_sample": (_dp, [vp, i]),
            "orc_sd_peek_IM": (_dp, [vp]),
            "orc_sd_peek_lnnorms": (_dp, [vp]),
            "orc_sd_peek_invUsample": (_dp, [vp]),
            "orc_sd_get_nnls_stats": (None, [vp, vp]),
            "orc_sd_compute_IM": (None, [vp, _dp]),
            "orc_sd_get_timers": (None, [vp, _dp]),
            "orc_target_m2lnL": (d, [vp, _dp]),
            "orc_apes_new": (vp, [i, i, i, i, d, d, i, d, d, d, i]),
            "orc_apes_free": (None, [vp]),
            "orc_apes_set_cov_type": (None, [vp, i, _dp, i]),
            "orc_apes_set_exploration": (None, [vp, C.c_uint]),
            "orc_apes_get_fallback_counts": (None, [vp, C.POINTER(C.c_long)]),
            "orc_apes_run": (None, [vp, vp, _dp, _dp, i, vp, C.POINTER(C.c_ubyte), i]),
            "orc_apes_get_timers": (None, [vp, _dp]),
            "orc_apes_peek_thetastar": (_dp, [vp]),
            "orc_apes_peek_m2lnp_star": (_dp, [vp]),
            "orc_apes_peek_m2lnp_cur": (_dp, [vp]),
            "orc_log_gaussian_integral": (d, [d, d, d, d, _dp]),
            "orc_fill_rand_cov": (None, [_dp, i, d, d, d, vp]),
            "orc_cholesky_decomp_U": (i, [_dp, i, i]),
            "orc_set_blas_threads": (None, [i]),
            "orc_get_max_threads": (i, []),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class _Kernel(C.Structure):
    _fields_ = [("kind", C.c_int), ("d", C.c_int), ("nu", C.c_double)]


class RNG:
    """NcmRNG restatement: gsl_rng_mt19937 seeded with gsl_rng_set."""

    def __init__(self, seed: int):
        self._s = _RNG()
        lib().orc_rng_set(C.byref(self._s), seed)

    @property
    def ptr(self):
        return C.byref(self._s)

    def get(self) -> int:
        return lib().orc_rng_get(self.ptr)

    def uniform(self) -> float:
        return lib().orc_rng_uniform(self.ptr)

    def uniform_pos(self) -> float:
        return lib().orc_rng_uniform_pos(self.ptr)

    def flat(self, a, b) -> float:
        return lib().orc_ran_flat(self.ptr, a, b)

    def gaussian(self, sigma=1.0) -> float:
        return lib().orc_ran_gaussian(self.ptr, sigma)

    def gaussian_ziggurat(self, sigma=1.0) -> float:
        return lib().orc_ran_gaussian_ziggurat(self.ptr, sigma)

    def chisq(self, nu) -> float:
        return lib().orc_ran_chisq(self.ptr, nu)

    def gamma(self, a, b=1.0) -> float:
        return lib().orc_ran_gamma(self.ptr, a, b)

    def beta(self, a, b) -> float:
        return lib().orc_ran_beta(self.ptr, a, b)

    def state(self):
        return (list(self._s.mt), self._s.mti)


class Kernel:
    def __init__(self, kind: int, d: int, nu: float = 3.0):
        self._k = _Kernel(kind, d, nu)
        self.kind, self.d, self.nu = kind, d, nu

    @property
    def ptr(self):
        return C.byref(self._k)

    def get_rot_bandwidth(self, n: float) -> float:
        return lib().orc_kernel_get_rot_bandwidth(self.ptr, float(n))

    def get_lnnorm(self, U: np.ndarray) -> float:
        U = np.ascontiguousarray(U, dtype=np.float64)
        return lib().orc_kernel_get_lnnorm(self.ptr, _p(U), U.shape[1])

    def eval_unnorm(self, chi2: float) -> float:
        return lib().orc_kernel_eval_unnorm(self.ptr, float(chi2))

    def eval_unnorm_vec(self, chi2: np.ndarray, stride: int = 1) -> np.ndarray:
        chi2 = np.ascontiguousarray(chi2, dtype=np.float64)
        n = (chi2.size + stride - 1) // stride
        out = np.zeros(n)
        lib().orc_kernel_eval_unnorm_vec(self.ptr, _p(chi2), stride, _p(out), 1, n)
        return out

    def eval_sum0_gamma_lambda(self, chi2, weights, lnnorms):
        chi2, weights, lnnorms = (np.ascontiguousarray(a, dtype=np.float64) for a in (chi2, weights, lnnorms))
        lnK = np.zeros_like(chi2)
        g, l = C.c_double(), C.c_double()
        lib().orc_kernel_eval_sum0_gamma_lambda(self.ptr, _p(chi2), _p(weights), _p(lnnorms), _p(lnK), chi2.size, C.byref(g), C.byref(l))
        return g.value, l.value

    def eval_sum1_gamma_lambda(self, chi2, weights, lnnorm: float):
        chi2, weights = (np.ascontiguousarray(a, dtype=np.float64) for a in (chi2, weights))
        lnK = np.zeros_like(chi2)
        g, l = C.c_double(), C.c_double()
        lib().orc_kernel_eval_sum1_gamma_lambda(self.ptr, _p(chi2), _p(weights), float(lnnorm), _p(lnK), chi2.size, C.byref(g), C.byref(l))
        return g.value, l.value

    def sample(self, U: np.ndarray, href: float, mu: np.ndarray, rng: RNG) -> np.ndarray:
        U = np.ascontiguousarray(U, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        x = np.zeros(self.d)
        lib().orc_kernel_sample(self.ptr, _p(U), U.shape[1], float(href), _p(mu), _p(x), rng.ptr)
        return x


def nnls_solve(A: np.ndarray, f: np.ndarray, reltol: float = np.finfo(float).eps):
    A = np.ascontiguousarray(A, dtype=np.float64)
    f = np.ascontiguousarray(f, dtype=np.float64)
    x = np.zeros(A.shape[1])
    st = _NNLSStats()
    rnorm = lib().orc_nnls_solve(_p(A), A.shape[0], A.shape[1], A.shape[1], _p(x), _p(f), reltol, C.byref(st))
    stats = {k: getattr(st, k) for k, _ in _NNLSStats._fields_}
    return x, rnorm, stats


def stats_Qn(x) -> float:
    """Rousseeuw-Croux Q_n scale of a sample (gsl_sort + gsl_stats_Qn_from_sorted_data)."""
    v = np.sort(np.ascontiguousarray(x, dtype=np.float64))
    return lib().orc_stats_Qn_from_sorted_data(_p(v), len(v))


def nmsimplex2_minimize(f, x0, step, size_tol=1e-3, max_iter=1000):
    """GSL nmsimplex2 restatement driven as ncm_stats_dist.c:681-692 does; f maps a numpy vector to a float."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    step = np.ascontiguousarray(step, dtype=np.float64)
    n = len(x0)
    cb = _FMIN_FN(lambda xp, nn, _: float(f(np.array([xp[k] for k in range(nn)]))))
    h = lib().orc_nmsimplex2_new(n)
    try:
        it = lib().orc_nmsimplex2_minimize(h, cb, None, _p(x0), _p(step), float(size_tol), int(max_iter))
        xp = lib().orc_nmsimplex2_x(h)
        return np.array([xp[k] for k in range(n)]), lib().orc_nmsimplex2_fval(h), lib().orc_nmsimplex2_size(h), it
    finally:
        lib().orc_nmsimplex2_free(h)


def lm_dif(func, p0, n, x=None, itmax=1000, opts=None, reference=False):
    """dlevmar_dif: the restatement (orc_lm_dif) or, with reference=True, the reference's own build.
    func(p) -> residual model hx[n]; returns (p, info[10], iterations)."""
    p = np.ascontiguousarray(p0, dtype=np.float64).copy()
    m = len(p)

    def _cb(pp, hx, mm, nn, _):
        v = np.asarray(func(np.array([pp[k] for k in range(mm)])), dtype=np.float64)
        for k in range(nn):
            hx[k] = v[k]

    cb = _LM_FN(_cb)
    info = np.zeros(10)
    xo = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    o = None if opts is None else np.ascontiguousarray(opts, dtype=np.float64)
    if reference:
        L = ref_levmar()
        if L is None:
            raise RuntimeError("oracle/_ref/liblevmar_ref.so not built")
        it = L.dlevmar_dif(cb, _p(p), None if xo is None else _p(xo), m, n, itmax, None if o is None else _p(o), _p(info), None, None, None)
    else:
        it = lib().orc_lm_dif(cb, _p(p), None if xo is None else _p(xo), m, n, itmax, None if o is None else _p(o), _p(info), None)
    return p, info, it


class StatsDist:
    """orc_sd wrapper mirroring the ncm_stats_dist_* call names."""

    def __init__(self, sd_type: int, kernel_kind: int, d: int, nu: float = 3.0, cv_type: int = CV_NONE):
        self._h = lib().orc_sd_new(sd_type, kernel_kind, float(nu), d, cv_type)
        self.d = d
        self.sd_type = sd_type

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sd_free(self._h)
            self._h = None

    def set_over_smooth(self, v): lib().orc_sd_set_over_smooth(self._h, float(v))
    def set_shrink(self, v): lib().orc_sd_set_shrink(self._h, float(v))
    def set_split_frac(self, v): lib().orc_sd_set_split_frac(self._h, float(v))
    def get_over_smooth(self): return lib().orc_sd_get_over_smooth(self._h)

    def cv_trace(self):
        """(ln over_smooth, objective) of every objective evaluation of the last CV prepare / prepare_interp."""
        n = lib().orc_sd_get_cv_trace(self._h, None, None, 0)
        a, b = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        lib().orc_sd_get_cv_trace(self._h, _p(a), _p(b), n)
        return a[:n], b[:n]
    def set_use_threads(self, v): lib().orc_sd_set_use_threads(self._h, int(v))
    def set_cov_type(self, v): lib().orc_sd_set_cov_type(self._h, int(v))
    def set_local_frac(self, v): lib().orc_sd_set_local_frac(self._h, float(v))
    def set_use_rot_href(self, v): lib().orc_sd_set_use_rot_href(self._h, int(v))
    def set_nearPD_maxiter(self, v): lib().orc_sd_set_nearPD_maxiter(self._h, int(v))

    def set_cov_fixed(self, cov):
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        lib().orc_sd_set_cov_fixed(self._h, _p(cov), cov.shape[1])

    def reset(self): lib().orc_sd_reset(self._h)

    def add_obs(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lib().orc_sd_add_obs(self._h, _p(x))

    def add_obs_matrix(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        for row in X:
            lib().orc_sd_add_obs(self._h, _p(np.ascontiguousarray(row)))

    def prepare(self) -> int:
        return lib().orc_sd_prepare(self._h)

    def prepare_interp(self, m2lnp) -> int:
        m2lnp = np.ascontiguousarray(m2lnp, dtype=np.float64)
        return lib().orc_sd_prepare_interp(self._h, _p(m2lnp), m2lnp.size)

    def eval(self, x) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_sd_eval(self._h, _p(x))

    def eval_m2lnp(self, x) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_sd_eval_m2lnp(self._h, _p(x))

    def eval_m2lnp_batch(self, X, nthreads: int = 1) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.zeros(X.shape[0])
        lib().orc_sd_eval_m2lnp_batch(self._h, _p(X), X.shape[1], X.shape[0], _p(out), nthreads)
        return out

    def eval_batch(self, X, nthreads: int = 1) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.zeros(X.shape[0])
        lib().orc_sd_eval_batch(self._h, _p(X), X.shape[1], X.shape[0], _p(out), nthreads)
        return out

    def set_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.size == self.get_n_kernels()
        lib().orc_sd_set_weights(self._h, _p(w))

    def kernel_choose(self, rng: RNG) -> int:
        return lib().orc_sd_kernel_choose(self._h, rng.ptr)

    def sample(self, rng: RNG) -> np.ndarray:
        x = np.zeros(self.d)
        lib().orc_sd_sample(self._h, _p(x), rng.ptr)
        return x

    def get_n_obs(self): return lib().orc_sd_get_n_obs(self._h)
    def get_n_kernels(self): return lib().orc_sd_get_n_kernels(self._h)
    def get_sample_size(self): return lib().orc_sd_get_sample_size(self._h)
    def get_href(self): return lib().orc_sd_get_href(self._h)
    def get_rnorm(self): return lib().orc_sd_get_rnorm(self._h)
    def get_lnnorm(self, i): return lib().orc_sd_get_lnnorm(self._h, i)

    def _arr(self, ptr, shape):
        return np.ctypeslib.as_array(ptr, shape=shape).copy()

    def peek_weights(self): return self._arr(lib().orc_sd_peek_weights(self._h), (self.get_n_kernels(),))
    def peek_cov_decomp(self, i): return self._arr(lib().orc_sd_peek_cov_decomp(self._h, i), (self.d, self.d))
    def peek_full_cov(self): return self._arr(lib().orc_sd_peek_full_cov(self._h), (self.d, self.d))
    def peek_full_cov_decomp(self): return self._arr(lib().orc_sd_peek_full_cov_decomp(self._h), (self.d, self.d))
    def peek_sample(self, i): return self._arr(lib().orc_sd_peek_sample(self._h, i), (self.d,))
    def peek_IM(self): return self._arr(lib().orc_sd_peek_IM(self._h), (self.get_n_obs(), self.get_n_kernels()))
    def peek_lnnorms(self): return self._arr(lib().orc_sd_peek_lnnorms(self._h), (self.get_n_kernels(),))
    def peek_invUsample(self): return self._arr(lib().orc_sd_peek_invUsample(self._h), (self.get_n_obs(), self.d))

    def peek_cov_array(self):
        n = self.get_n_kernels()
        if self.sd_type == SD_KDE:
            return np.repeat(self.peek_full_cov_decomp()[None], n, axis=0)
        return self._arr(lib().orc_sd_peek_cov_decomp(self._h, 0), (n, self.d, self.d))

    def compute_IM(self) -> np.ndarray:
        IM = np.zeros((self.get_n_obs(), self.get_n_kernels()))
        lib().orc_sd_compute_IM(self._h, _p(IM))
        return IM

    def nnls_stats(self):
        st = _NNLSStats()
        lib().orc_sd_get_nnls_stats(self._h, C.byref(st))
        return {k: getattr(st, k) for k, _ in _NNLSStats._fields_}

    def timers(self):
        t = np.zeros(3)
        lib().orc_sd_get_timers(self._h, _p(t))
        return {"prepare_kernel": t[0], "IM": t[1], "NNLS": t[2]}


class Target:
    def __init__(self, kind: int, d: int, lb, ub, mu=None, cov=None):
        self.kind, self.d = kind, d
        self.lb = np.ascontiguousarray(lb, dtype=np.float64)
        self.ub = np.ascontiguousarray(ub, dtype=np.float64)
        self.mu = np.ascontiguousarray(mu if mu is not None else np.zeros(d), dtype=np.float64)
        if cov is not None:
            U = np.ascontiguousarray(cov, dtype=np.float64).copy()
            assert lib().orc_cholesky_decomp_U(_p(U), d, d) == 0
            self.U = np.triu(U)
        else:
            self.U = np.eye(d)
        self.U = np.ascontiguousarray(self.U)
        self._t = _Target(kind, d, _p(self.mu), _p(self.U), _p(self.lb), _p(self.ub))

    @property
    def ptr(self):
        return C.byref(self._t)

    def m2lnL(self, x) -> float:
        x = np.ascontiguousarray(x, dtype=np.float64)
        return lib().orc_target_m2lnL(self.ptr, _p(x))


class APES:
    def __init__(self, nwalkers, d, sd_type=SD_VKDE, kernel_kind=KERNEL_ST, nu=1.0, over_smooth=1.0, use_interp=True,
                 shrink=0.01, random_walk_prob=0.02, local_frac=0.05, use_threads=False):
        self._h = lib().orc_apes_new(nwalkers, d, sd_type, kernel_kind, float(nu), float(over_smooth), int(use_interp),
                                     float(shrink), float(random_walk_prob), float(local_frac), int(use_threads))
        self.nwalkers, self.d = nwalkers, d

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_apes_free(self._h)
            self._h = None

    def run(self, target: Target, theta: np.ndarray, m2lnL: np.ndarray, iters: int, rng: RNG, nthreads: int = 1):
        assert theta.dtype == np.float64 and theta.flags["C_CONTIGUOUS"] and theta.shape == (self.nwalkers, self.d)
        acc = np.zeros((iters, self.nwalkers), dtype=np.uint8)
        lib().orc_apes_run(self._h, target.ptr, _p(theta), _p(m2lnL), iters, rng.ptr, acc.ctypes.data_as(C.POINTER(C.c_ubyte)), nthreads)
        return acc

    def set_cov_type(self, cov_type, cov_fixed=None):
        cf = np.ascontiguousarray(cov_fixed, dtype=np.float64) if cov_fixed is not None else None
        lib().orc_apes_set_cov_type(self._h, int(cov_type), _p(cf) if cf is not None else None, self.d)

    def set_exploration(self, n):
        lib().orc_apes_set_exploration(self._h, int(n))

    def fallback_counts(self):
        """(prepare_interp calls, dsysv fallbacks, dgels fallbacks) of the NNLS since the object was made."""
        n = (C.c_long * 3)()
        lib().orc_apes_get_fallback_counts(self._h, n)
        return int(n[0]), int(n[1]), int(n[2])

    def timers(self):
        t = np.zeros(6)
        lib().orc_apes_get_timers(self._h, _p(t))
        return dict(zip(["prepare_kernel", "IM", "NNLS", "sample", "eval", "likelihood_accept"], t))

    def peek_thetastar(self):
        return np.ctypeslib.as_array(lib().orc_apes_peek_thetastar(self._h), shape=(self.nwalkers, self.d)).copy()

    def peek_m2lnp_star(self):
        return np.ctypeslib.as_array(lib().orc_apes_peek_m2lnp_star(self._h), shape=(self.nwalkers,)).copy()

    def peek_m2lnp_cur(self):
        return np.ctypeslib.as_array(lib().orc_apes_peek_m2lnp_cur(self._h), shape=(self.nwalkers,)).copy()


def fill_rand_cov(n, sigma_min, sigma_max, cor_level, rng: RNG) -> np.ndarray:
    cm = np.zeros((n, n))
    lib().orc_fill_rand_cov(_p(cm), n, sigma_min, sigma_max, cor_level, rng.ptr)
    return cm


def log_gaussian_integral(xl, xu, mu, sigma) -> float:
    s = C.c_double()
    return lib().orc_log_gaussian_integral(xl, xu, mu, sigma, C.byref(s))
