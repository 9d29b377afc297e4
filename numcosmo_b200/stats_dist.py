"""Python face of the host mirror (libncm_stats_dist_b200.so), named after the GI objects numcosmo_py
exposes (numcosmo_py/ncm.pyi:11823-11930): Ncm.StatsDistKernelGauss / ST, Ncm.StatsDistKDE / VKDE, Ncm.RNG,
Ncm.FitESMCMCWalkerAPES.  Vectors and matrices are numpy arrays.

Every density evaluation, interpolation-matrix build and NNLS solve runs in the CUDA library behind the
C ABI; a g_error of the reference surfaces here as NcmError.
"""
from __future__ import annotations

import ctypes as C
import os
import enum

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libncm_stats_dist_b200.so")

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class NcmError(RuntimeError):
    """A g_error / g_assert of the reference path."""


class StatsDistCV(enum.IntEnum):
    NONE = 0
    SPLIT = 1
    SPLIT_NOFIT = 2
    LOO = 3


class StatsDistKDECovType(enum.IntEnum):
    SAMPLE = 0
    FIXED = 1
    ROBUST_DIAG = 2
    ROBUST = 3


class FitESMCMCWalkerAPESMethod(enum.IntEnum):
    KDE = 0
    VKDE = 1


class FitESMCMCWalkerAPESKType(enum.IntEnum):
    CAUCHY = 0
    ST3 = 1
    GAUSS = 2


class _NcmVector(C.Structure):
    _fields_ = [("data", _dp), ("len", C.c_uint), ("stride", C.c_uint), ("ref", C.c_int), ("own", C.c_bool)]


class _NcmMatrix(C.Structure):
    _fields_ = [("data", _dp), ("nrows", C.c_uint), ("ncols", C.c_uint), ("tda", C.c_uint), ("ref", C.c_int), ("own", C.c_bool)]


class _MVND(C.Structure):
    _fields_ = [("d", C.c_uint), ("mu", _dp), ("U", _dp)]


_ERR_CB = C.CFUNCTYPE(None, C.c_char_p, C.c_void_p)
_M2LNL_CB = C.CFUNCTYPE(None, _dp, C.c_uint, C.c_uint, _dp, C.c_void_p)
_last_error: list[str] = []


@_ERR_CB
def _on_error(msg, _):
    _last_error.append(msg.decode(errors="replace"))


_lib = None


def lib():
    global _lib
    if _lib is None:
        capi.load()  # fails loudly if the CUDA library is missing
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError(f"{HOST_LIB_PATH} is missing: build it with `make -C numcosmo_b200/host`.")
        L = C.CDLL(HOST_LIB_PATH)
        d, u, i = C.c_double, C.c_uint, C.c_int
        VP = C.POINTER(_NcmVector)
        MP = C.POINTER(_NcmMatrix)
        sig = {
            "ncm_b200_set_error_handler": (None, [_ERR_CB, _vp]),
            "ncm_b200_set_device": (None, [i]),
            "ncm_b200_set_num_threads": (None, [i]),
            "ncm_b200_set_host_prepare_kernel": (None, [i]),
            "ncm_vector_new_data_static": (VP, [_dp, u, u]),
            "ncm_vector_free": (None, [VP]),
            "ncm_matrix_free": (None, [MP]),
            "ncm_rng_seeded_new": (_vp, [C.c_char_p, C.c_ulong]),
            "ncm_rng_free": (None, [_vp]),
            "ncm_rng_set_seed": (None, [_vp, C.c_ulong]),
            "ncm_rng_gen_ulong": (C.c_ulong, [_vp]),
            "ncm_rng_uniform01_gen": (d, [_vp]),
            "ncm_rng_uniform01_pos_gen": (d, [_vp]),
            "ncm_rng_uniform_gen": (d, [_vp, d, d]),
            "ncm_rng_gaussian_gen": (d, [_vp, d, d]),
            "ncm_rng_ugaussian_gen": (d, [_vp]),
            "ncm_rng_chisq_gen": (d, [_vp, d]),
            "ncm_rng_beta_gen": (d, [_vp, d, d]),
            "ncm_stats_dist_kernel_gauss_new": (_vp, [u]),
            "ncm_stats_dist_kernel_st_new": (_vp, [u, d]),
            "ncm_stats_dist_kernel_free": (None, [_vp]),
            "ncm_stats_dist_kernel_get_dim": (u, [_vp]),
            "ncm_stats_dist_kernel_get_rot_bandwidth": (d, [_vp, d]),
            "ncm_stats_dist_kernel_get_lnnorm": (d, [_vp, MP]),
            "ncm_stats_dist_kernel_eval_unnorm": (d, [_vp, d]),
            "ncm_stats_dist_kernel_eval_unnorm_vec": (None, [_vp, VP, VP]),
            "ncm_stats_dist_kernel_eval_sum0_gamma_lambda": (None, [_vp, VP, VP, VP, VP, _dp, _dp]),
            "ncm_stats_dist_kernel_eval_sum1_gamma_lambda": (None, [_vp, VP, VP, d, VP, _dp, _dp]),
            "ncm_stats_dist_kernel_sample": (None, [_vp, MP, d, VP, VP, _vp]),
            "ncm_stats_dist_kernel_st_get_nu": (d, [_vp]),
            "ncm_stats_dist_kde_new": (_vp, [_vp, i]),
            "ncm_stats_dist_vkde_new": (_vp, [_vp, i]),
            "ncm_stats_dist_free": (None, [_vp]),
            "ncm_stats_dist_get_dim": (u, [_vp]),
            "ncm_stats_dist_get_sample_size": (u, [_vp]),
            "ncm_stats_dist_get_n_kernels": (u, [_vp]),
            "ncm_stats_dist_get_href": (d, [_vp]),
            "ncm_stats_dist_set_over_smooth": (None, [_vp, d]),
            "ncm_stats_dist_get_over_smooth": (d, [_vp]),
            "ncm_stats_dist_set_split_frac": (None, [_vp, d]),
            "ncm_stats_dist_get_split_frac": (d, [_vp]),
            "ncm_stats_dist_set_shrink": (None, [_vp, d]),
            "ncm_stats_dist_get_shrink": (d, [_vp]),
            "ncm_stats_dist_set_print_fit": (None, [_vp, i]),
            "ncm_stats_dist_get_print_fit": (i, [_vp]),
            "ncm_stats_dist_set_cv_type": (None, [_vp, i]),
            "ncm_stats_dist_get_cv_type": (i, [_vp]),
            "ncm_stats_dist_set_use_threads": (None, [_vp, i]),
            "ncm_stats_dist_get_use_threads": (i, [_vp]),
            "ncm_stats_dist_prepare": (None, [_vp]),
            "ncm_stats_dist_prepare_interp": (None, [_vp, VP]),
            "ncm_stats_dist_eval": (d, [_vp, VP]),
            "ncm_stats_dist_eval_m2lnp": (d, [_vp, VP]),
            "ncm_stats_dist_eval_m2lnp_array": (None, [_vp, MP, VP]),
            "ncm_stats_dist_eval_array": (None, [_vp, MP, VP]),
            "ncm_stats_dist_kernel_choose": (u, [_vp, _vp]),
            "ncm_stats_dist_sample": (None, [_vp, VP, _vp]),
            "ncm_stats_dist_get_rnorm": (d, [_vp]),
            "ncm_stats_dist_add_obs": (None, [_vp, VP]),
            "ncm_stats_dist_peek_cov_decomp": (MP, [_vp, u]),
            "ncm_stats_dist_peek_full_cov_decomp": (MP, [_vp]),
            "ncm_stats_dist_peek_full_cov": (MP, [_vp]),
            "ncm_stats_dist_get_lnnorm": (d, [_vp, u]),
            "ncm_stats_dist_peek_weights": (VP, [_vp]),
            "ncm_stats_dist_get_Ki": (None, [_vp, u, C.POINTER(VP), C.POINTER(MP), _dp, _dp]),
            "ncm_stats_dist_reset": (None, [_vp]),
            "ncm_stats_dist_kde_set_nearPD_maxiter": (None, [_vp, u]),
            "ncm_stats_dist_kde_get_nearPD_maxiter": (u, [_vp]),
            "ncm_stats_dist_kde_set_cov_type": (None, [_vp, i]),
            "ncm_stats_dist_kde_get_cov_type": (i, [_vp]),
            "ncm_stats_dist_kde_set_cov_fixed": (None, [_vp, MP]),
            "ncm_stats_dist_vkde_set_local_frac": (None, [_vp, d]),
            "ncm_stats_dist_vkde_get_local_frac": (d, [_vp]),
            "ncm_stats_dist_vkde_set_use_rot_href": (None, [_vp, i]),
            "ncm_stats_dist_vkde_get_use_rot_href": (i, [_vp]),
            "ncm_stats_dist_b200_get_nnls_stats": (None, [_vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
            "ncm_stats_dist_b200_get_nnls_lowrank_stats": (None, [_vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(i)]),
            "ncm_stats_dist_b200_get_nnls_fallback_stats": (None, [_vp, C.POINTER(i), C.POINTER(i)]),
            "ncm_stats_dist_b200_comm_unique_id": (i, [C.c_char_p]),
            "ncm_stats_dist_b200_comm_init": (i, [_vp, i, i, C.c_char_p]),
            "ncm_stats_dist_b200_get_cv_trace": (i, [_vp, _dp, _dp, i]),
            "ncm_stats_dist_b200_get_timers": (None, [_vp, _dp, C.POINTER(C.c_longlong), _dp]),
            "ncm_stats_dist_b200_enable_timers": (None, [_vp, i]),
            "ncm_stats_dist_b200_peek_ctx": (_vp, [_vp]),
            "ncm_fit_esmcmc_walker_apes_new_full": (_vp, [u, u, i, i, d, i]),
            "ncm_fit_esmcmc_walker_apes_free": (None, [_vp]),
            "ncm_fit_esmcmc_walker_apes_set_over_smooth": (None, [_vp, d]),
            "ncm_fit_esmcmc_walker_apes_set_shrink": (None, [_vp, d]),
            "ncm_fit_esmcmc_walker_apes_set_random_walk_prob": (None, [_vp, d]),
            "ncm_fit_esmcmc_walker_apes_set_random_walk_scale": (None, [_vp, d]),
            "ncm_fit_esmcmc_walker_apes_set_method": (None, [_vp, i]),
            "ncm_fit_esmcmc_walker_apes_set_k_type": (None, [_vp, i]),
            "ncm_fit_esmcmc_walker_apes_get_method": (i, [_vp]),
            "ncm_fit_esmcmc_walker_apes_get_k_type": (i, [_vp]),
            "ncm_fit_esmcmc_walker_apes_get_over_smooth": (d, [_vp]),
            "ncm_fit_esmcmc_walker_apes_get_shrink": (d, [_vp]),
            "ncm_fit_esmcmc_walker_apes_get_random_walk_prob": (d, [_vp]),
            "ncm_fit_esmcmc_walker_apes_get_random_walk_scale": (d, [_vp]),
            "ncm_fit_esmcmc_walker_apes_interp": (i, [_vp]),
            "ncm_fit_esmcmc_walker_apes_get_use_threads": (i, [_vp]),
            "ncm_fit_esmcmc_walker_apes_use_interp": (None, [_vp, i]),
            "ncm_fit_esmcmc_walker_apes_set_use_threads": (None, [_vp, i]),
            "ncm_fit_esmcmc_walker_apes_set_local_frac": (None, [_vp, d]),
            "ncm_fit_esmcmc_walker_apes_set_exploration": (None, [_vp, u]),
            "ncm_fit_esmcmc_walker_apes_set_cov_fixed_from_mset": (None, [_vp, C.POINTER(C.c_double)]),
            "ncm_fit_esmcmc_walker_apes_set_cov_robust_diag": (None, [_vp]),
            "ncm_fit_esmcmc_walker_apes_set_cov_robust": (None, [_vp]),
            "ncm_fit_esmcmc_walker_apes_ref": (_vp, [_vp]),
            "ncm_fit_esmcmc_walker_apes_b200_get_pregen_stats": (None, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
            "ncm_fit_esmcmc_walker_apes_peek_sds": (None, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
            "ncm_fit_esmcmc_walker_apes_setup": (None, [_vp, _dp, _dp, _dp, _dp, u, u, _vp]),
            "ncm_fit_esmcmc_walker_apes_step": (None, [_vp, _dp, _dp, u]),
            "ncm_fit_esmcmc_walker_apes_prob_norm": (d, [_vp, u]),
            "ncm_fit_esmcmc_walker_apes_peek_thetastar": (_dp, [_vp]),
            "ncm_fit_esmcmc_walker_apes_peek_m2lnp_star": (_dp, [_vp]),
            "ncm_fit_esmcmc_walker_apes_peek_m2lnp_cur": (_dp, [_vp]),
            "ncm_b200_esmcmc_run": (None, [_vp, _vp, _vp, _dp, _dp, _dp, _dp, u, _vp, C.POINTER(C.c_ubyte), _dp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        L.ncm_b200_set_error_handler(_on_error, None)
        _lib = L
    return _lib


def _check():
    if _last_error:
        msg = "; ".join(_last_error)
        _last_error.clear()
        raise NcmError(msg)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class _Vec:
    """Borrowed NcmVector view over a numpy array (kept alive for the call)."""

    def __init__(self, a: np.ndarray):
        assert a.dtype == np.float64 and a.ndim == 1 and a.flags["C_CONTIGUOUS"]
        self.a = a
        self.p = lib().ncm_vector_new_data_static(a.ctypes.data_as(_dp), a.size, 1)

    def __enter__(self):
        return self.p

    def __exit__(self, *exc):
        lib().ncm_vector_free(self.p)


class _Mat:
    def __init__(self, a: np.ndarray):
        assert a.dtype == np.float64 and a.ndim == 2 and a.flags["C_CONTIGUOUS"]
        self.a = a
        self.s = _NcmMatrix(a.ctypes.data_as(_dp), a.shape[0], a.shape[1], a.shape[1], 1, False)

    def __enter__(self):
        return C.pointer(self.s)

    def __exit__(self, *exc):
        pass


def _mat_to_np(mp) -> np.ndarray:
    m = mp.contents
    a = np.ctypeslib.as_array(m.data, shape=(m.nrows, m.tda))
    return a[:, : m.ncols].copy()


def _vec_to_np(vp) -> np.ndarray:
    v = vp.contents
    return np.ctypeslib.as_array(v.data, shape=(v.len * v.stride,))[:: v.stride].copy()


class RNG:
    """Ncm.RNG: gsl_rng_mt19937 (the GSL default generator NcmRNG uses)."""

    def __init__(self, seed: int = 0, algo: str | None = None):
        self._h = lib().ncm_rng_seeded_new(algo.encode() if algo else None, seed)
        _check()

    @classmethod
    def seeded_new(cls, algo, seed):
        return cls(seed, algo)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ncm_rng_free(self._h)
            self._h = None

    def set_seed(self, seed): lib().ncm_rng_set_seed(self._h, seed)
    def gen_ulong(self): return lib().ncm_rng_gen_ulong(self._h)
    def uniform01_gen(self): return lib().ncm_rng_uniform01_gen(self._h)
    def uniform01_pos_gen(self): return lib().ncm_rng_uniform01_pos_gen(self._h)
    def uniform_gen(self, a, b): return lib().ncm_rng_uniform_gen(self._h, a, b)
    def gaussian_gen(self, mu, sigma): return lib().ncm_rng_gaussian_gen(self._h, mu, sigma)
    def ugaussian_gen(self): return lib().ncm_rng_ugaussian_gen(self._h)
    def chisq_gen(self, nu): return lib().ncm_rng_chisq_gen(self._h, nu)
    def beta_gen(self, a, b): return lib().ncm_rng_beta_gen(self._h, a, b)


class StatsDistKernel:
    def __init__(self, handle, dim):
        self._h = handle
        self.dim = dim
        _check()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ncm_stats_dist_kernel_free(self._h)
            self._h = None

    def get_dim(self): return lib().ncm_stats_dist_kernel_get_dim(self._h)
    def get_rot_bandwidth(self, n): return lib().ncm_stats_dist_kernel_get_rot_bandwidth(self._h, float(n))

    def get_lnnorm(self, cov_decomp):
        with _Mat(_f64(cov_decomp)) as m:
            return lib().ncm_stats_dist_kernel_get_lnnorm(self._h, m)

    def eval_unnorm(self, chi2): return lib().ncm_stats_dist_kernel_eval_unnorm(self._h, float(chi2))

    def eval_unnorm_vec(self, chi2):
        chi2 = _f64(chi2)
        out = np.zeros_like(chi2)
        with _Vec(chi2) as c, _Vec(out) as o:
            lib().ncm_stats_dist_kernel_eval_unnorm_vec(self._h, c, o)
        _check()
        return out

    def eval_sum0_gamma_lambda(self, chi2, weights, lnnorms):
        chi2, weights, lnnorms = _f64(chi2), _f64(weights), _f64(lnnorms)
        lnK = np.zeros_like(chi2)
        g, l = C.c_double(), C.c_double()
        with _Vec(chi2) as c, _Vec(weights) as w, _Vec(lnnorms) as n, _Vec(lnK) as k:
            lib().ncm_stats_dist_kernel_eval_sum0_gamma_lambda(self._h, c, w, n, k, C.byref(g), C.byref(l))
        _check()
        return g.value, l.value

    def eval_sum1_gamma_lambda(self, chi2, weights, lnnorm):
        chi2, weights = _f64(chi2), _f64(weights)
        lnK = np.zeros_like(chi2)
        g, l = C.c_double(), C.c_double()
        with _Vec(chi2) as c, _Vec(weights) as w, _Vec(lnK) as k:
            lib().ncm_stats_dist_kernel_eval_sum1_gamma_lambda(self._h, c, w, float(lnnorm), k, C.byref(g), C.byref(l))
        _check()
        return g.value, l.value

    def sample(self, cov_decomp, href, mu, rng: RNG):
        y = np.zeros(self.dim)
        with _Mat(_f64(cov_decomp)) as m, _Vec(_f64(mu)) as mv, _Vec(y) as yv:
            lib().ncm_stats_dist_kernel_sample(self._h, m, float(href), mv, yv, rng._h)
        return y


class StatsDistKernelGauss(StatsDistKernel):
    def __init__(self, dim: int):
        super().__init__(lib().ncm_stats_dist_kernel_gauss_new(dim), dim)

    @classmethod
    def new(cls, dim):
        return cls(dim)


class StatsDistKernelST(StatsDistKernel):
    def __init__(self, dim: int, nu: float):
        super().__init__(lib().ncm_stats_dist_kernel_st_new(dim, float(nu)), dim)

    @classmethod
    def new(cls, dim, nu):
        return cls(dim, nu)

    def get_nu(self): return lib().ncm_stats_dist_kernel_st_get_nu(self._h)


class StatsDist:
    """Ncm.StatsDist: the calls APES and numcosmo_py/interpolation/stats_dist.py make."""

    def __init__(self, handle, kernel: StatsDistKernel):
        self._h = handle
        self._kernel = kernel
        self.dim = kernel.dim
        _check()

    def __del__(self):
        if getattr(self, "_h", None) and getattr(self, "_owned", True):
            lib().ncm_stats_dist_free(self._h)
            self._h = None

    # properties
    def get_dim(self): return lib().ncm_stats_dist_get_dim(self._h)
    def get_sample_size(self): return lib().ncm_stats_dist_get_sample_size(self._h)
    def get_n_kernels(self): return lib().ncm_stats_dist_get_n_kernels(self._h)
    def get_href(self): return lib().ncm_stats_dist_get_href(self._h)
    def set_over_smooth(self, v): lib().ncm_stats_dist_set_over_smooth(self._h, float(v))
    def get_over_smooth(self): return lib().ncm_stats_dist_get_over_smooth(self._h)
    def set_shrink(self, v):
        lib().ncm_stats_dist_set_shrink(self._h, float(v))
        _check()

    def get_shrink(self): return lib().ncm_stats_dist_get_shrink(self._h)
    def set_cv_type(self, v): lib().ncm_stats_dist_set_cv_type(self._h, int(v))
    def get_cv_type(self): return StatsDistCV(lib().ncm_stats_dist_get_cv_type(self._h))
    def set_use_threads(self, v): lib().ncm_stats_dist_set_use_threads(self._h, int(v))
    def get_use_threads(self): return bool(lib().ncm_stats_dist_get_use_threads(self._h))

    def set_split_frac(self, v):
        lib().ncm_stats_dist_set_split_frac(self._h, float(v))
        _check()

    def get_split_frac(self): return lib().ncm_stats_dist_get_split_frac(self._h)
    def peek_kernel(self): return self._kernel

    def add_obs(self, y):
        with _Vec(_f64(y)) as v:
            lib().ncm_stats_dist_add_obs(self._h, v)
        _check()

    def reset(self): lib().ncm_stats_dist_reset(self._h)

    def prepare(self):
        lib().ncm_stats_dist_prepare(self._h)
        _check()

    def prepare_interp(self, m2lnp):
        with _Vec(_f64(m2lnp)) as v:
            lib().ncm_stats_dist_prepare_interp(self._h, v)
        _check()

    def eval(self, x) -> float:
        with _Vec(_f64(x)) as v:
            r = lib().ncm_stats_dist_eval(self._h, v)
        _check()
        return r

    def eval_m2lnp(self, x) -> float:
        with _Vec(_f64(x)) as v:
            r = lib().ncm_stats_dist_eval_m2lnp(self._h, v)
        _check()
        return r

    def eval_m2lnp_array(self, X) -> np.ndarray:
        X = _f64(X)
        out = np.zeros(X.shape[0])
        with _Mat(X) as m, _Vec(out) as o:
            lib().ncm_stats_dist_eval_m2lnp_array(self._h, m, o)
        _check()
        return out

    def eval_array(self, X) -> np.ndarray:
        X = _f64(X)
        out = np.zeros(X.shape[0])
        with _Mat(X) as m, _Vec(out) as o:
            lib().ncm_stats_dist_eval_array(self._h, m, o)
        _check()
        return out

    def kernel_choose(self, rng: RNG) -> int:
        return lib().ncm_stats_dist_kernel_choose(self._h, rng._h)

    def sample(self, rng: RNG) -> np.ndarray:
        x = np.zeros(self.dim)
        with _Vec(x) as v:
            lib().ncm_stats_dist_sample(self._h, v, rng._h)
        _check()
        return x

    def get_rnorm(self): return lib().ncm_stats_dist_get_rnorm(self._h)

    def peek_cov_decomp(self, i) -> np.ndarray:
        m = lib().ncm_stats_dist_peek_cov_decomp(self._h, i)
        _check()
        return _mat_to_np(m)

    def peek_full_cov_decomp(self): return _mat_to_np(lib().ncm_stats_dist_peek_full_cov_decomp(self._h))
    def peek_full_cov(self): return _mat_to_np(lib().ncm_stats_dist_peek_full_cov(self._h))

    def get_lnnorm(self, i):
        r = lib().ncm_stats_dist_get_lnnorm(self._h, i)
        _check()
        return r

    def peek_weights(self): return _vec_to_np(lib().ncm_stats_dist_peek_weights(self._h))

    def get_Ki(self, i):
        L = lib()
        yv, cm = C.POINTER(_NcmVector)(), C.POINTER(_NcmMatrix)()
        n, w = C.c_double(), C.c_double()
        L.ncm_stats_dist_get_Ki(self._h, i, C.byref(yv), C.byref(cm), C.byref(n), C.byref(w))
        _check()
        y, cov = _vec_to_np(yv), _mat_to_np(cm)
        L.ncm_vector_free(yv)
        L.ncm_matrix_free(cm)
        return y, cov, n.value, w.value

    # instrumentation
    def nnls_stats(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().ncm_stats_dist_b200_get_nnls_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        out = {"n_chol": a.value, "n_lu": b.value, "n_outer": c.value, "n_passive": d.value}
        lib().ncm_stats_dist_b200_get_nnls_fallback_stats(self._h, C.byref(a), C.byref(b))
        out["n_qr"] = b.value
        lib().ncm_stats_dist_b200_get_nnls_lowrank_stats(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        out.update({"n_lowrank": a.value, "n_lowrank_fallback": b.value, "n_trinv": c.value, "max_lowrank_k": d.value})
        return out

    def cv_trace(self):
        """(ln over_smooth, objective or rnorm) of every objective evaluation of the last prepare / prepare_interp
        under a cross-validation mode (ncm_stats_dist.c:484-701, 1018-1072)."""
        n = lib().ncm_stats_dist_b200_get_cv_trace(self._h, None, None, 0)
        a, b = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        lib().ncm_stats_dist_b200_get_cv_trace(self._h, a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), n)
        return a[:n], b[:n]

    def enable_timers(self, on=True):
        lib().ncm_stats_dist_b200_enable_timers(self._h, int(on))
        _check()

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        """Multi-rank (SPMD) mode: shard the IM rows / query rows of this object over the ranks (NCCL inside the C ABI)."""
        assert len(unique_id) == 128
        lib().ncm_stats_dist_b200_comm_init(self._h, nranks, rank, unique_id)
        _check()

    def comm_init_from_torch(self, group=None):
        """comm_init with the NCCL id drawn on rank 0 and broadcast through torch.distributed (any backend)."""
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=group)
        self.comm_init(world, rank, uid[0])

    def get_timers(self):
        ms = np.zeros(len(capi.T_NAMES))
        n, h = C.c_longlong(), C.c_double()
        lib().ncm_stats_dist_b200_get_timers(self._h, ms.ctypes.data_as(_dp), C.byref(n), C.byref(h))
        out = dict(zip(capi.T_NAMES, ms.tolist()))
        out["host_prepare_kernel"] = h.value
        return out, n.value


class StatsDistKDE(StatsDist):
    def __init__(self, kernel: StatsDistKernel, cv_type: StatsDistCV = StatsDistCV.NONE):
        super().__init__(lib().ncm_stats_dist_kde_new(kernel._h, int(cv_type)), kernel)

    @classmethod
    def new(cls, kernel, cv_type):
        return cls(kernel, cv_type)

    def set_nearPD_maxiter(self, v): lib().ncm_stats_dist_kde_set_nearPD_maxiter(self._h, int(v))
    def get_nearPD_maxiter(self): return lib().ncm_stats_dist_kde_get_nearPD_maxiter(self._h)
    def set_cov_type(self, v): lib().ncm_stats_dist_kde_set_cov_type(self._h, int(v))
    def get_cov_type(self): return StatsDistKDECovType(lib().ncm_stats_dist_kde_get_cov_type(self._h))

    def set_cov_fixed(self, cov):
        with _Mat(_f64(cov)) as m:
            lib().ncm_stats_dist_kde_set_cov_fixed(self._h, m)
        _check()


class StatsDistVKDE(StatsDistKDE):
    def __init__(self, kernel: StatsDistKernel, cv_type: StatsDistCV = StatsDistCV.NONE):
        StatsDist.__init__(self, lib().ncm_stats_dist_vkde_new(kernel._h, int(cv_type)), kernel)

    def set_local_frac(self, v):
        lib().ncm_stats_dist_vkde_set_local_frac(self._h, float(v))
        _check()

    def get_local_frac(self): return lib().ncm_stats_dist_vkde_get_local_frac(self._h)
    def set_use_rot_href(self, v): lib().ncm_stats_dist_vkde_set_use_rot_href(self._h, int(v))
    def get_use_rot_href(self): return bool(lib().ncm_stats_dist_vkde_get_use_rot_href(self._h))


class _BorrowedSD(StatsDistVKDE):   # the walker builds NcmStatsDistVKDE objects for both methods (walker_apes.c:563-572)
    def __init__(self, handle, dim):
        self._h = handle
        self.dim = dim
        self._kernel = None
        self._owned = False


TARGET_MVND, TARGET_ROSENBROCK, TARGET_FUNNEL = "mvnd", "rosenbrock", "funnel"


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if lib().ncm_stats_dist_b200_comm_unique_id(buf) != 0:
        raise NcmError("ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
    return buf.raw


class FitESMCMCWalkerAPES:
    """Ncm.FitESMCMCWalkerAPES + the accept loop of Ncm.FitESMCMC around it (numcosmo_py/sampling/apes.py)."""

    def __init__(self, nwalkers: int, nparams: int, method=FitESMCMCWalkerAPESMethod.VKDE, k_type=FitESMCMCWalkerAPESKType.CAUCHY,
                 over_smooth: float = 1.0, use_interp: bool = True):
        self._h = lib().ncm_fit_esmcmc_walker_apes_new_full(nwalkers, nparams, int(method), int(k_type), float(over_smooth), int(use_interp))
        self.nwalkers, self.nparams = nwalkers, nparams
        _check()

    @classmethod
    def new_full(cls, nwalkers, nparams, method, k_type, over_smooth, use_interp):
        return cls(nwalkers, nparams, method, k_type, over_smooth, use_interp)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ncm_fit_esmcmc_walker_apes_free(self._h)
            self._h = None

    def set_over_smooth(self, v): lib().ncm_fit_esmcmc_walker_apes_set_over_smooth(self._h, float(v))

    def set_shrink(self, v):
        lib().ncm_fit_esmcmc_walker_apes_set_shrink(self._h, float(v))
        _check()

    def set_random_walk_prob(self, v):
        lib().ncm_fit_esmcmc_walker_apes_set_random_walk_prob(self._h, float(v))
        _check()

    def set_random_walk_scale(self, v):
        lib().ncm_fit_esmcmc_walker_apes_set_random_walk_scale(self._h, float(v))
        _check()

    def set_exploration(self, n): lib().ncm_fit_esmcmc_walker_apes_set_exploration(self._h, int(n))

    def set_method(self, m):
        lib().ncm_fit_esmcmc_walker_apes_set_method(self._h, int(m))
        _check()

    def set_k_type(self, k):
        lib().ncm_fit_esmcmc_walker_apes_set_k_type(self._h, int(k))
        _check()

    def get_method(self): return FitESMCMCWalkerAPESMethod(lib().ncm_fit_esmcmc_walker_apes_get_method(self._h))
    def get_k_type(self): return FitESMCMCWalkerAPESKType(lib().ncm_fit_esmcmc_walker_apes_get_k_type(self._h))
    def get_over_smooth(self): return lib().ncm_fit_esmcmc_walker_apes_get_over_smooth(self._h)
    def get_shrink(self): return lib().ncm_fit_esmcmc_walker_apes_get_shrink(self._h)
    def get_random_walk_prob(self): return lib().ncm_fit_esmcmc_walker_apes_get_random_walk_prob(self._h)
    def get_random_walk_scale(self): return lib().ncm_fit_esmcmc_walker_apes_get_random_walk_scale(self._h)
    def interp(self): return bool(lib().ncm_fit_esmcmc_walker_apes_interp(self._h))
    def get_use_threads(self):
        r = bool(lib().ncm_fit_esmcmc_walker_apes_get_use_threads(self._h))
        _check()
        return r

    def use_interp(self, v): lib().ncm_fit_esmcmc_walker_apes_use_interp(self._h, int(v))
    def set_use_threads(self, v): lib().ncm_fit_esmcmc_walker_apes_set_use_threads(self._h, int(v))

    def set_local_frac(self, v):
        lib().ncm_fit_esmcmc_walker_apes_set_local_frac(self._h, float(v))
        _check()

    def set_cov_fixed_from_mset(self, fparam_scales):
        """walker_apes.c:1450-1473; `fparam_scales` = [ncm_mset_fparam_get_scale (mset, i)] (the only thing the reference reads from the mset)."""
        s = np.ascontiguousarray(fparam_scales, dtype=np.float64)
        lib().ncm_fit_esmcmc_walker_apes_set_cov_fixed_from_mset(self._h, s.ctypes.data_as(C.POINTER(C.c_double)))
        _check()

    def set_cov_robust_diag(self): lib().ncm_fit_esmcmc_walker_apes_set_cov_robust_diag(self._h)

    def set_cov_robust(self): lib().ncm_fit_esmcmc_walker_apes_set_cov_robust(self._h)

    def peek_sds(self):
        a, b = C.c_void_p(), C.c_void_p()
        lib().ncm_fit_esmcmc_walker_apes_peek_sds(self._h, C.byref(a), C.byref(b))
        return _BorrowedSD(a, self.nparams), _BorrowedSD(b, self.nparams)

    def cv_trace(self):
        """(ln over_smooth, objective or rnorm) of every objective evaluation of the last prepare / prepare_interp
        under a cross-validation mode (ncm_stats_dist.c:484-701, 1018-1072)."""
        n = lib().ncm_stats_dist_b200_get_cv_trace(self._h, None, None, 0)
        a, b = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        lib().ncm_stats_dist_b200_get_cv_trace(self._h, a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), n)
        return a[:n], b[:n]

    def enable_timers(self, on=True):
        for sd in self.peek_sds():
            sd.enable_timers(on)

    def comm_init_from_torch(self, group=None):
        """Multi-rank (SPMD) APES: every rank runs the same chain with the same generator seed; the density work of the two NcmStatsDist
        objects behind the walker is sharded over the ranks (one NCCL communicator each)."""
        for sd in self.peek_sds():
            sd.comm_init_from_torch(group)

    def peek_thetastar(self):
        return np.ctypeslib.as_array(lib().ncm_fit_esmcmc_walker_apes_peek_thetastar(self._h), shape=(self.nwalkers, self.nparams)).copy()

    def peek_m2lnp_star(self):
        return np.ctypeslib.as_array(lib().ncm_fit_esmcmc_walker_apes_peek_m2lnp_star(self._h), shape=(self.nwalkers,)).copy()

    def peek_m2lnp_cur(self):
        return np.ctypeslib.as_array(lib().ncm_fit_esmcmc_walker_apes_peek_m2lnp_cur(self._h), shape=(self.nwalkers,)).copy()

    def pregen_stats(self):
        """(blocks whose proposal draws were generated while the GPU solved the NNLS, blocks that had to be replayed serially)."""
        a, b = C.c_longlong(), C.c_longlong()
        lib().ncm_fit_esmcmc_walker_apes_b200_get_pregen_stats(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def run(self, target, lb, ub, theta: np.ndarray, m2lnL: np.ndarray, iters: int, rng: RNG, target_args=None, record_accept: bool = True):
        """iters whole-ensemble iterations in place on theta [W x d] / m2lnL [W]; returns (accepted, timers_ms)."""
        L = lib()
        assert theta.dtype == np.float64 and theta.flags["C_CONTIGUOUS"] and theta.shape == (self.nwalkers, self.nparams)
        assert m2lnL.dtype == np.float64 and m2lnL.flags["C_CONTIGUOUS"]
        lb, ub = _f64(lb), _f64(ub)
        keep = []
        user = None
        if callable(target):
            def _cb(Xp, n, d, outp, _u):
                X = np.ctypeslib.as_array(Xp, shape=(n, d))
                np.ctypeslib.as_array(outp, shape=(n,))[:] = target(X)
            cb = _M2LNL_CB(_cb)
            keep.append(cb)
            fptr = C.cast(cb, _vp)
        elif target == TARGET_ROSENBROCK:
            fptr = C.cast(L.ncm_b200_target_rosenbrock, _vp)
        elif target == TARGET_FUNNEL:
            fptr = C.cast(L.ncm_b200_target_funnel, _vp)
        elif target == TARGET_MVND:
            mu, U = _f64(target_args[0]), _f64(target_args[1])
            st = _MVND(self.nparams, mu.ctypes.data_as(_dp), U.ctypes.data_as(_dp))
            keep += [mu, U, st]
            fptr = C.cast(L.ncm_b200_target_mvnd, _vp)
            user = C.cast(C.pointer(st), _vp)
        else:
            raise ValueError(target)
        acc = np.zeros((iters, self.nwalkers), dtype=np.uint8) if record_accept else None
        timers = np.zeros(8)
        L.ncm_b200_esmcmc_run(self._h, fptr, user, lb.ctypes.data_as(_dp), ub.ctypes.data_as(_dp), theta.ctypes.data_as(_dp), m2lnL.ctypes.data_as(_dp),
                              iters, rng._h, acc.ctypes.data_as(C.POINTER(C.c_ubyte)) if acc is not None else None, timers.ctypes.data_as(_dp))
        _check()
        names = ["prepare_kernel", "IM", "NNLS", "sample", "eval", "likelihood_accept", "copies", "total"]
        return acc, dict(zip(names, timers.tolist()))
