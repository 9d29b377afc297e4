// NcmFitESMCMCWalkerAPES mirror + the ESMCMC accept loop around it.
//
//   set_sys           ncm_fit_esmcmc_walker_apes.c:509-608  (BOTH methods build NcmStatsDistVKDE objects, :563-572)
//   prepare_random_walk / random_walk_sample / sample      :646-739
//   setup             :741-817   + NEW: one batched GPU eval of {theta*_k, theta_k} for the whole block
//   transition_prob   :819-859   (host: O(d) random-walk mixture on top of the cached density values)
//   step / prob_norm  :861-919   (step only reads the cache)
//   run               ncm_fit_esmcmc.c:2136-2148 (jumps), :2151-2232 (run_interval), :2235-2288 (run, ki = 0)
//
// Proposal draws ahead of the weights.  The serial stream of a block -- per walker one uniform for the random-walk decision, then either the
// bounded Gaussian steps of the random walk or {one uniform for kernel_choose, d unit normals, one chi-square} -- does not depend on the
// interpolation weights; only the kernel index picked by that uniform and the affine map do.  With use_interp the draws of the block are therefore
// generated on a second host thread WHILE the GPU solves the NNLS (the host thread of prepare_interp mostly waits on the device), and applied
// afterwards.  The speculation assumes what is true for all but pathological bounds: no proposal falls outside the box (walker_apes.c:735-737
// would then draw again).  If one does, or if prepare_interp re-prepared on a cut sample (dynamic-range guard), the generator state of the block
// start is restored and the block is sampled serially, exactly as before: the stream, hence the accepted sequence, is the same either way.
#include <chrono>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>
#include "internal.h"

namespace {
constexpr double LN2PI     = 1.8378770664093454835606594728112352797227949472755668;
constexpr double ERF_BOUND = 1.0;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ncm_util.c:645-716
double log_gaussian_integral(double xl, double xu, double mu, double sigma) {
  const double zl = (xl - mu) / sigma, zu = (xu - mu) / sigma;
  if (zl == zu) return -INFINITY;
  double ul, uu;
  if (zl < zu) {
    ul = zl * M_SQRT1_2;
    uu = zu * M_SQRT1_2;
  } else {
    ul = zu * M_SQRT1_2;
    uu = zl * M_SQRT1_2;
  }
  if (ul > ERF_BOUND) return log(fabs(0.5 * (erfc(ul) - erfc(uu))));
  if (uu < -ERF_BOUND) return log(fabs(0.5 * (erfc(-ul) - erfc(-uu))));
  if ((uu > ERF_BOUND) && (ul < -ERF_BOUND)) return log1p(-0.5 * (erfc(uu) + erfc(-ul)));
  return log(fabs(0.5 * (erf(uu) - erf(ul))));
}

struct RandomWalk {
  std::vector<double> std, lb, ub;
};
}   // namespace

struct _NcmFitESMCMCWalkerAPES {
  guint size, size_2, nparams;
  guint a_size = 0, a_nparams = 0;   // the configuration the objects below were built for (walker_apes.c:515-523)
  int a_mk = -1;
  NcmFitESMCMCWalkerAPESMethod method;
  NcmFitESMCMCWalkerAPESKType k_type;
  double over_smooth, shrink, random_walk_prob, random_walk_scale, local_frac;
  gboolean use_interp, use_threads;
  guint exploration;
  NcmStatsDist *sd0, *sd1;
  std::vector<double> thetastar, m2lnp_star, m2lnp_cur, m2lnL_s0, m2lnL_s1, jumps;
  RandomWalk rw0, rw1;
  double t_sample_ms, t_eval_ms;
  // draws generated ahead of the weights (one block): random-walk flag, kernel_choose uniform, unit normals, chi-square
  std::vector<unsigned char> pre_rw;
  std::vector<double> pre_p, pre_z, pre_chisq;
  long long n_spec_blocks, n_spec_fallbacks;
  NcmB200PinnedVec evalQ, evalOut;   // staging of the batched density evaluation of a block
  int ref = 1;
};

namespace {

// _ncm_fit_esmcmc_walker_apes_vkde_check_sizes (walker_apes.c:489-507): reads the local fraction the OBJECTS hold
void vkde_check_sizes(NcmFitESMCMCWalkerAPES *a) {
  for (NcmStatsDist *sd : {a->sd0, a->sd1}) {
    const guint cov_estimates = (guint) (ncm_stats_dist_vkde_get_local_frac(sd) * a->size_2);
    if (cov_estimates < 2) {
      ncm_b200_error("Number of walkers per block (%d) is too low for the current dimension (%d).\n\tToo few points (%d) to estimate local covariances.",
                     a->size_2, a->nparams, cov_estimates);
      return;
    }
  }
}

// _ncm_fit_esmcmc_walker_apes_set_sys (walker_apes.c:509-598): the objects are rebuilt only when size, dimension, method or kernel
// type changed; fresh objects start from their own defaults (local_frac 0.05, covariance type SAMPLE) and receive the walker's
// over_smooth, shrink and use_threads
void set_sys(NcmFitESMCMCWalkerAPES *a) {
  const int mk = (int) a->method * 1000 + (int) a->k_type;
  if (a->size == a->a_size && a->nparams == a->a_nparams && mk == a->a_mk) return;
  a->a_size    = a->size;
  a->a_nparams = a->nparams;
  a->a_mk      = mk;
  ncm_stats_dist_clear(&a->sd0);
  ncm_stats_dist_clear(&a->sd1);
  if (a->size % 2 != 0) {
    ncm_b200_error("_ncm_fit_esmcmc_walker_apes_set_sys: assertion failed (self->size %% 2 == 0)");
    return;
  }
  a->size_2 = a->size / 2;
  a->m2lnp_star.assign(a->size, 0.0);
  a->m2lnp_cur.assign(a->size, 0.0);
  a->m2lnL_s0.assign(a->size_2, 0.0);
  a->m2lnL_s1.assign(a->size_2, 0.0);
  a->jumps.assign(a->size, 0.0);
  a->thetastar.assign((size_t) a->size * a->nparams, 0.0);
  NcmStatsDistKernel *kernel = nullptr;
  switch (a->k_type) {
    case NCM_FIT_ESMCMC_WALKER_APES_KTYPE_CAUCHY: kernel = ncm_stats_dist_kernel_st_new(a->nparams, 1.0); break;
    case NCM_FIT_ESMCMC_WALKER_APES_KTYPE_ST3: kernel = ncm_stats_dist_kernel_st_new(a->nparams, 3.0); break;
    default: kernel = ncm_stats_dist_kernel_gauss_new(a->nparams); break;
  }
  if (kernel == nullptr) return;
  // METHOD_KDE and METHOD_VKDE both construct NcmStatsDistVKDE (walker_apes.c:563-572)
  a->sd0 = ncm_stats_dist_vkde_new(kernel, NCM_STATS_DIST_CV_NONE);
  a->sd1 = ncm_stats_dist_vkde_new(kernel, NCM_STATS_DIST_CV_NONE);
  if (a->method == NCM_FIT_ESMCMC_WALKER_APES_METHOD_VKDE) vkde_check_sizes(a);
  ncm_stats_dist_kernel_free(kernel);
  for (NcmStatsDist *sd : {a->sd0, a->sd1}) {
    ncm_stats_dist_set_over_smooth(sd, a->over_smooth);
    ncm_stats_dist_set_shrink(sd, a->shrink);
    ncm_stats_dist_set_use_threads(sd, a->use_threads);
  }
}

bool valid_bounds(const double *lb, const double *ub, const double *x, guint d) {
  for (guint i = 0; i < d; i++)
    if ((x[i] < lb[i]) || (x[i] > ub[i])) return false;
  return true;
}

void prepare_random_walk(NcmFitESMCMCWalkerAPES *a, NcmStatsDist *sd, RandomWalk &rw, const double *lb, const double *ub) {
  if (a->random_walk_prob > 0.0) {
    NcmMatrix *cov = ncm_stats_dist_peek_full_cov(sd);
    rw.std.resize(a->nparams);
    rw.lb.assign(lb, lb + a->nparams);
    rw.ub.assign(ub, ub + a->nparams);
    for (guint i = 0; i < a->nparams; i++) {
      const double var = ncm_matrix_get(cov, i, i);
      if (var <= 0.0) {
        ncm_b200_error("Invalid covariance matrix: diagonal element %d is non-positive.", i);
        return;
      }
      rw.std[i] = sqrt(var) * 0.25;   // literal 0.25; random_walk_scale is stored but unused (walker_apes.c:684-688)
    }
  }
}

void apes_sample(NcmFitESMCMCWalkerAPES *a, NcmStatsDist *sd, const RandomWalk &rw, const double *lb, const double *ub, const double *theta,
                 double *thetastar, NcmRNG *rng) {
  const guint d = a->nparams;
  NcmVector *ts = ncm_vector_new_data_static(thetastar, d, 1);
  do {
    if (a->random_walk_prob != 0.0 && ncm_rng_uniform01_pos_gen(rng) < a->random_walk_prob) {
      for (guint i = 0; i < d; i++) {
        double x;
        do {
          x = ncm_rng_gaussian_gen(rng, theta[i], rw.std[i]);
        } while ((x < rw.lb[i]) || (x > rw.ub[i]));
        thetastar[i] = x;
      }
    } else {
      ncm_stats_dist_sample(sd, ts, rng);
    }
  } while (!valid_bounds(lb, ub, thetastar, d));
  ncm_vector_free(ts);
}

double transition_prob(NcmFitESMCMCWalkerAPES *a, const RandomWalk &rw, const double *theta, const double *thetastar, double m2lnp_sd) {
  if (!(a->random_walk_prob > 0.0)) return m2lnp_sd;
  double m2lnp_rw = 0.0;
  for (guint i = 0; i < a->nparams; i++) {
    const double sd      = rw.std[i];
    const double ln_norm = 0.5 * LN2PI + log(sd) + log_gaussian_integral(rw.lb[i], rw.ub[i], theta[i], sd);
    const double r       = (thetastar[i] - theta[i]) / sd;
    m2lnp_rw += r * r + 2.0 * ln_norm;
  }
  m2lnp_rw += -2.0 * log(a->random_walk_prob);
  m2lnp_sd += -2.0 * log1p(-a->random_walk_prob);
  if (m2lnp_sd < m2lnp_rw) return m2lnp_sd - 2.0 * log1p(exp(-0.5 * (m2lnp_rw - m2lnp_sd)));
  return m2lnp_rw - 2.0 * log1p(exp(-0.5 * (m2lnp_sd - m2lnp_rw)));
}

// the stream of one block, consumed in the order of _ncm_fit_esmcmc_walker_apes_sample (walker_apes.c:716-739) under the assumption that no
// proposal has to be redrawn; random-walk proposals are complete here (they never depend on the density estimate)
void pregenerate_block(NcmFitESMCMCWalkerAPES *a, NcmStatsDistKernel *sdk, const RandomWalk &rw, const double *theta, guint ki, guint kf, NcmRNG *rng) {
  const guint d  = a->nparams;
  const bool st  = sdk->kind == NCM_SD_GPU_KERNEL_ST;
  const guint nb = kf - ki;
  a->pre_rw.assign(nb, 0);
  a->pre_p.assign(nb, 0.0);
  a->pre_z.assign((size_t) nb * d, 0.0);
  a->pre_chisq.assign(nb, 0.0);
  for (guint k = ki; k < kf; k++) {
    const guint q = k - ki;
    if (a->random_walk_prob != 0.0 && ncm_rng_uniform01_pos_gen(rng) < a->random_walk_prob) {
      a->pre_rw[q] = 1;
      const double *th = &theta[(size_t) k * d];
      double *ts       = &a->thetastar[(size_t) k * d];
      for (guint i = 0; i < d; i++) {
        double x;
        do {
          x = ncm_rng_gaussian_gen(rng, th[i], rw.std[i]);
        } while ((x < rw.lb[i]) || (x > rw.ub[i]));
        ts[i] = x;
      }
    } else {
      a->pre_p[q] = ncm_rng_uniform_gen(rng, 0.0, 1.0);
      for (guint i = 0; i < d; i++) a->pre_z[(size_t) q * d + i] = ncm_rng_ugaussian_gen(rng);
      if (st) a->pre_chisq[q] = ncm_rng_chisq_gen(rng, sdk->nu);
    }
  }
}

// kernel index and affine map for the pre-drawn proposals; false when one of them leaves the box (the reference would draw again)
bool apply_pregenerated(NcmFitESMCMCWalkerAPES *a, NcmStatsDist *sd, const double *lb, const double *ub, guint ki, guint kf) {
  const guint d            = a->nparams;
  NcmStatsDistKernel *sdk  = ncm_stats_dist_peek_kernel(sd);
  const double href        = ncm_stats_dist_get_href(sd);
  GPtrArray *sample        = ncm_stats_dist_peek_sample_array(sd);
  (void) ncm_b200_kernel_choose_p(sd, 0.0);   // builds the cumulative-weight table once; the loop below only reads it
  int all_inside = 1;
  // with the draws in hand the proposals are independent of each other: threads change nothing in the values
#pragma omp parallel for schedule(static) reduction(&& : all_inside) if (a->use_threads)
  for (long k = (long) ki; k < (long) kf; k++) {
    const guint q = (guint) (k - (long) ki);
    double *ts    = &a->thetastar[(size_t) k * d];
    if (!a->pre_rw[q]) {
      const guint i = ncm_b200_kernel_choose_p(sd, a->pre_p[q]);
      NcmVector tsv{ts, d, 1, 1, false};
      ncm_b200_kernel_sample_from(sdk, ncm_stats_dist_peek_cov_decomp(sd, i), href, (NcmVector *) sample->pdata[i], &tsv, &a->pre_z[(size_t) q * d], a->pre_chisq[q]);
    }
    all_inside = all_inside && valid_bounds(lb, ub, ts, d);
  }
  return all_inside != 0;
}

void setup_block(NcmFitESMCMCWalkerAPES *a, int block, const double *lb, const double *ub, const double *theta, const double *m2lnL, guint ki,
                 guint kf, NcmRNG *rng) {
  const guint d          = a->nparams;
  NcmStatsDist *sd       = block == 0 ? a->sd0 : a->sd1;
  RandomWalk &rw         = block == 0 ? a->rw0 : a->rw1;
  std::vector<double> &s = block == 0 ? a->m2lnL_s0 : a->m2lnL_s1;
  const guint c0         = block == 0 ? a->size_2 : 0;   // centres come from the OTHER half

  NcmB200ProfScope prof_block("setup_block(total)");
  NcmB200HostProf *pslot = ncm_b200_prof_on() ? ncm_b200_prof_slot("reset+add_obs") : nullptr;
  const double tp0 = pslot ? ncm_b200_now_ms() : 0.0;
  ncm_stats_dist_reset(sd);
  for (guint i = 0; i < a->size_2; i++) {
    s[i]         = m2lnL[c0 + i];
    NcmVector *v = ncm_vector_new_data_static(const_cast<double *>(&theta[(size_t) (c0 + i) * d]), d, 1);
    ncm_stats_dist_add_obs(sd, v);
    ncm_vector_free(v);
  }
  if (pslot) {
    pslot->ms += ncm_b200_now_ms() - tp0;
    pslot->calls++;
  }
  bool speculated = false;
  NcmRNG rng0;   // generator state at the start of the block's draws
  if (a->use_interp) {
    NcmVector *mv = ncm_vector_new_data_static(s.data(), a->size_2, 1);
    static const bool spec_on = getenv("NCM_B200_APES_PREGEN") == nullptr || atoi(getenv("NCM_B200_APES_PREGEN")) != 0;
    bool begun = false;
    if (spec_on) {
      NcmB200ProfScope prof("prepare_interp_begin(prepare_kernel)");
      begun = ncm_b200_prepare_interp_begin(sd, mv);
    }
    if (spec_on && begun) {
      // the kernel (hence the full covariance the random walk scales with) is prepared; the weights are what the GPU works on next
      prepare_random_walk(a, sd, rw, lb, ub);
      if (!ncm_b200_error_pending()) {
        rng0       = *rng;
        speculated = true;
        a->n_spec_blocks++;
        const RandomWalk rw_used = rw;
        std::thread gen([&]() { pregenerate_block(a, ncm_stats_dist_peek_kernel(sd), rw_used, theta, ki, kf, rng); });
        {
          NcmB200ProfScope prof("prepare_interp_finish");
          ncm_b200_prepare_interp_finish(sd, mv);
        }
        {
          NcmB200ProfScope prof("join pregenerate thread");
          gen.join();
        }
        if (!ncm_b200_error_pending()) {
          prepare_random_walk(a, sd, rw, lb, ub);   // the dynamic-range guard may have re-prepared the object on a cut sample
          if (rw.std != rw_used.std || ncm_stats_dist_get_n_kernels(sd) != a->size_2) {
            *rng       = rng0;
            speculated = false;
            a->n_spec_fallbacks++;
          }
        }
      } else {
        ncm_b200_prepare_interp_finish(sd, mv);
      }
    } else if (!spec_on) {
      ncm_stats_dist_prepare_interp(sd, mv);
    }
    ncm_vector_free(mv);
  } else {
    ncm_stats_dist_prepare(sd);
  }
  if (ncm_b200_error_pending()) return;
  if (!speculated) prepare_random_walk(a, sd, rw, lb, ub);

  double t0 = now_ms();
  if (speculated && !apply_pregenerated(a, sd, lb, ub, ki, kf)) {
    *rng       = rng0;   // a proposal left the box: the reference redraws it, which shifts the stream -- replay the block serially
    speculated = false;
    a->n_spec_fallbacks++;
  }
  if (!speculated)
    for (guint k = ki; k < kf; k++) apes_sample(a, sd, rw, lb, ub, &theta[(size_t) k * d], &a->thetastar[(size_t) k * d], rng);
  double t1 = now_ms();
  a->t_sample_ms += t1 - t0;

  // NEW: both density evaluations of _apes_step (walker_apes.c:872-873) for every walker of the block in ONE call
  const guint nb = kf - ki;
  // page-locked staging kept by the walker: the queries go up and the densities come back at PCIe rate, no bounce buffer
  a->evalQ.resize((size_t) 2 * a->size_2 * d);
  a->evalOut.resize((size_t) 2 * a->size_2);
  NcmMatrix Qm, *Q = &Qm;
  Qm.data = a->evalQ.data(); Qm.nrows = 2 * nb; Qm.ncols = d; Qm.tda = d; Qm.ref = 1; Qm.own = false;
  NcmVector outv, *out = &outv;
  outv.data = a->evalOut.data(); outv.len = 2 * nb; outv.stride = 1; outv.ref = 1; outv.own = false;
  memcpy(ncm_matrix_data(Q), &a->thetastar[(size_t) ki * d], sizeof(double) * nb * d);
  memcpy(ncm_matrix_data(Q) + (size_t) nb * d, &theta[(size_t) ki * d], sizeof(double) * nb * d);
  {
    NcmB200ProfScope prof("eval_m2lnp_array");
    ncm_stats_dist_eval_m2lnp_array(sd, Q, out);
  }
  // per-walker and independent (erf-heavy when the random-walk mixture is on): threads change nothing in the values
#pragma omp parallel for schedule(static) if (a->use_threads)
  for (long k = (long) ki; k < (long) kf; k++) {
    const double *th = &theta[(size_t) k * d], *ts = &a->thetastar[(size_t) k * d];
    a->m2lnp_star[k] = transition_prob(a, rw, th, ts, ncm_vector_get(out, (guint) (k - ki)));
    a->m2lnp_cur[k]  = transition_prob(a, rw, ts, th, ncm_vector_get(out, (guint) (nb + k - ki)));
  }
  a->t_eval_ms += now_ms() - t1;
}

}   // namespace

extern "C" {

NcmFitESMCMCWalkerAPES *ncm_fit_esmcmc_walker_apes_new_full(guint nwalkers, guint nparams, NcmFitESMCMCWalkerAPESMethod method,
                                                            NcmFitESMCMCWalkerAPESKType k_type, gdouble over_smooth, gboolean use_interp) {
  NcmFitESMCMCWalkerAPES *a = new NcmFitESMCMCWalkerAPES();
  a->size             = nwalkers;
  a->nparams          = nparams;
  a->method           = method;
  a->k_type           = k_type;
  a->over_smooth      = over_smooth;
  a->shrink           = 0.01;
  a->random_walk_prob = 0.02;
  a->random_walk_scale = 1.0;
  a->local_frac       = 0.05;
  a->use_interp       = use_interp;
  a->use_threads      = FALSE;
  a->exploration      = 0;
  a->sd0 = a->sd1 = nullptr;
  a->t_sample_ms = a->t_eval_ms = 0.0;
  a->n_spec_blocks = a->n_spec_fallbacks = 0;
  set_sys(a);
  return a;
}
// defaults of the property specs, walker_apes.c:358-475
NcmFitESMCMCWalkerAPES *ncm_fit_esmcmc_walker_apes_new(guint nwalkers, guint nparams) {
  return ncm_fit_esmcmc_walker_apes_new_full(nwalkers, nparams, NCM_FIT_ESMCMC_WALKER_APES_METHOD_VKDE, NCM_FIT_ESMCMC_WALKER_APES_KTYPE_CAUCHY,
                                             1.0, TRUE);
}
// ncm_fit_esmcmc_walker_apes.c:1052-1056
NcmFitESMCMCWalkerAPES *ncm_fit_esmcmc_walker_apes_ref(NcmFitESMCMCWalkerAPES *a) {
  a->ref++;
  return a;
}
void ncm_fit_esmcmc_walker_apes_free(NcmFitESMCMCWalkerAPES *a) {
  if (a == nullptr || --a->ref > 0) return;
  ncm_stats_dist_clear(&a->sd0);
  ncm_stats_dist_clear(&a->sd1);
  delete a;
}
void ncm_fit_esmcmc_walker_apes_clear(NcmFitESMCMCWalkerAPES **a) {
  if (a != nullptr && *a != nullptr) {
    ncm_fit_esmcmc_walker_apes_free(*a);
    *a = nullptr;
  }
}
// walker_apes.c:1110-1144 (the second message is the reference's own, copy-and-paste of the first included)
void ncm_fit_esmcmc_walker_apes_set_method(NcmFitESMCMCWalkerAPES *a, NcmFitESMCMCWalkerAPESMethod m) {
  if ((guint) m >= (guint) NCM_FIT_ESMCMC_WALKER_APES_METHOD_LEN) {
    ncm_b200_error("ncm_fit_esmcmc_walker_apes_set_method: invalid method `%d'.", (int) m);
    return;
  }
  a->method = m;
  set_sys(a);
}
void ncm_fit_esmcmc_walker_apes_set_k_type(NcmFitESMCMCWalkerAPES *a, NcmFitESMCMCWalkerAPESKType k) {
  if ((guint) k >= (guint) NCM_FIT_ESMCMC_WALKER_APES_KTYPE_LEN) {
    ncm_b200_error("ncm_fit_esmcmc_walker_apes_set_method: invalid method `%d'.", (int) k);
    return;
  }
  a->k_type = k;
  set_sys(a);
}
void ncm_fit_esmcmc_walker_apes_set_over_smooth(NcmFitESMCMCWalkerAPES *a, const gdouble os) {
  a->over_smooth = os;
  ncm_stats_dist_set_over_smooth(a->sd0, os);   // forwarded immediately (walker_apes.c:1162-1166)
  ncm_stats_dist_set_over_smooth(a->sd1, os);
}
// range checks and messages of walker_apes.c:1178-1227; shrink is only stored (the objects got theirs in set_sys)
void ncm_fit_esmcmc_walker_apes_set_shrink(NcmFitESMCMCWalkerAPES *a, const gdouble s) {
  if ((s < 0.0) || (s > 1.0)) {
    ncm_b200_error("ncm_fit_esmcmc_walker_apes_set_shrink: invalid shrink `%f'.", s);
    return;
  }
  a->shrink = s;
}
void ncm_fit_esmcmc_walker_apes_set_random_walk_prob(NcmFitESMCMCWalkerAPES *a, const gdouble p) {
  if ((p < 0.0) || (p > 1.0)) {
    ncm_b200_error("ncm_fit_esmcmc_walker_apes_set_random_walk_prob: invalid probability `%f'.", p);
    return;
  }
  a->random_walk_prob = p;
}
void ncm_fit_esmcmc_walker_apes_set_random_walk_scale(NcmFitESMCMCWalkerAPES *a, const gdouble s) {
  if (s <= 0.0) {
    ncm_b200_error("ncm_fit_esmcmc_walker_apes_set_random_walk_scale: invalid scale `%f'.", s);
    return;
  }
  a->random_walk_scale = s;
}
NcmFitESMCMCWalkerAPESMethod ncm_fit_esmcmc_walker_apes_get_method(NcmFitESMCMCWalkerAPES *a) { return a->method; }
NcmFitESMCMCWalkerAPESKType ncm_fit_esmcmc_walker_apes_get_k_type(NcmFitESMCMCWalkerAPES *a) { return a->k_type; }
gdouble ncm_fit_esmcmc_walker_apes_get_over_smooth(NcmFitESMCMCWalkerAPES *a) { return a->over_smooth; }
gdouble ncm_fit_esmcmc_walker_apes_get_shrink(NcmFitESMCMCWalkerAPES *a) { return a->shrink; }
gdouble ncm_fit_esmcmc_walker_apes_get_random_walk_prob(NcmFitESMCMCWalkerAPES *a) { return a->random_walk_prob; }
gdouble ncm_fit_esmcmc_walker_apes_get_random_walk_scale(NcmFitESMCMCWalkerAPES *a) { return a->random_walk_scale; }
void ncm_fit_esmcmc_walker_apes_use_interp(NcmFitESMCMCWalkerAPES *a, gboolean u) { a->use_interp = u; }
gboolean ncm_fit_esmcmc_walker_apes_interp(NcmFitESMCMCWalkerAPES *a) { return a->use_interp; }
void ncm_fit_esmcmc_walker_apes_set_use_threads(NcmFitESMCMCWalkerAPES *a, gboolean u) {
  a->use_threads = u;
  ncm_stats_dist_set_use_threads(a->sd0, u);
  ncm_stats_dist_set_use_threads(a->sd1, u);
}
// walker_apes.c:1378-1391: the two objects must agree with the walker (someone may have reached them through peek_sds)
gboolean ncm_fit_esmcmc_walker_apes_get_use_threads(NcmFitESMCMCWalkerAPES *a) {
  const gboolean u0 = ncm_stats_dist_get_use_threads(a->sd0), u1 = ncm_stats_dist_get_use_threads(a->sd1);
  if (!a->use_threads != !u0) ncm_b200_error("ncm_fit_esmcmc_walker_apes_get_use_threads: assertion failed (self->use_threads == use_threads0)");
  if (!u0 != !u1) ncm_b200_error("ncm_fit_esmcmc_walker_apes_get_use_threads: assertion failed (use_threads0 == use_threads1)");
  return u0;
}
void ncm_fit_esmcmc_walker_apes_peek_sds(NcmFitESMCMCWalkerAPES *a, NcmStatsDist **sd0, NcmStatsDist **sd1) {
  *sd0 = a->sd0;
  *sd1 = a->sd1;
}
// walker_apes.c:1428-1439: VKDE method only, forwarded to both objects (which keep it until set_sys rebuilds them), sizes re-checked
void ncm_fit_esmcmc_walker_apes_set_local_frac(NcmFitESMCMCWalkerAPES *a, gdouble lf) {
  if (a->method != NCM_FIT_ESMCMC_WALKER_APES_METHOD_VKDE) {
    ncm_b200_error("ncm_fit_esmcmc_walker_apes_set_local_frac: cannot set local fraction for a non-VKDE method.");
    return;
  }
  ncm_stats_dist_vkde_set_local_frac(a->sd0, lf);
  if (ncm_stats_dist_vkde_get_local_frac(a->sd0) != lf) return;   // refused by the object's own range assert (reported there)
  ncm_stats_dist_vkde_set_local_frac(a->sd1, lf);
  a->local_frac = lf;
  vkde_check_sizes(a);
}
void ncm_fit_esmcmc_walker_apes_set_exploration(NcmFitESMCMCWalkerAPES *a, guint e) { a->exploration = e; }
// ncm_fit_esmcmc_walker_apes.c:1450-1473.  The reference reads nothing from the NcmMSet but the scales of its free parameters
// (ncm_mset_fparam_get_scale, one per walker dimension): the mirror takes that array.  cov_fixed = diag (scale_i^2) on both halves.
void ncm_fit_esmcmc_walker_apes_set_cov_fixed_from_mset(NcmFitESMCMCWalkerAPES *a, const gdouble *fparam_scales) {
  NcmMatrix *cov_fixed = ncm_matrix_new(a->nparams, a->nparams);
  for (guint i = 0; i < a->nparams; i++)
    for (guint j = 0; j < a->nparams; j++) ncm_matrix_set(cov_fixed, i, j, i == j ? fparam_scales[i] * fparam_scales[i] : 0.0);
  ncm_stats_dist_kde_set_cov_type(a->sd0, NCM_STATS_DIST_KDE_COV_TYPE_FIXED);
  ncm_stats_dist_kde_set_cov_type(a->sd1, NCM_STATS_DIST_KDE_COV_TYPE_FIXED);
  ncm_stats_dist_kde_set_cov_fixed(a->sd0, cov_fixed);
  ncm_stats_dist_kde_set_cov_fixed(a->sd1, cov_fixed);
  ncm_matrix_free(cov_fixed);
}
// :1483-1490
void ncm_fit_esmcmc_walker_apes_set_cov_robust_diag(NcmFitESMCMCWalkerAPES *a) {
  ncm_stats_dist_kde_set_cov_type(a->sd0, NCM_STATS_DIST_KDE_COV_TYPE_ROBUST_DIAG);
  ncm_stats_dist_kde_set_cov_type(a->sd1, NCM_STATS_DIST_KDE_COV_TYPE_ROBUST_DIAG);
}
// :1500-1507
void ncm_fit_esmcmc_walker_apes_set_cov_robust(NcmFitESMCMCWalkerAPES *a) {
  ncm_stats_dist_kde_set_cov_type(a->sd0, NCM_STATS_DIST_KDE_COV_TYPE_ROBUST);
  ncm_stats_dist_kde_set_cov_type(a->sd1, NCM_STATS_DIST_KDE_COV_TYPE_ROBUST);
}
// instrumentation: blocks whose draws were generated ahead of the weights, and how many of those had to be replayed serially
void ncm_fit_esmcmc_walker_apes_b200_get_pregen_stats(NcmFitESMCMCWalkerAPES *a, long long *n_blocks, long long *n_fallbacks) {
  if (n_blocks) *n_blocks = a->n_spec_blocks;
  if (n_fallbacks) *n_fallbacks = a->n_spec_fallbacks;
}

void ncm_fit_esmcmc_walker_apes_setup(NcmFitESMCMCWalkerAPES *a, const gdouble *lb, const gdouble *ub, const gdouble *theta, const gdouble *m2lnL,
                                      guint ki, guint kf, NcmRNG *rng) {
  if (ki < a->size_2) setup_block(a, 0, lb, ub, theta, m2lnL, ki, kf < a->size_2 ? kf : a->size_2, rng);
  if (kf > a->size_2) setup_block(a, 1, lb, ub, theta, m2lnL, ki > a->size_2 ? ki : a->size_2, kf, rng);
  if (a->exploration > 0) a->exploration--;
}

void ncm_fit_esmcmc_walker_apes_step(NcmFitESMCMCWalkerAPES *a, const gdouble *theta, gdouble *thetastar, guint k) {
  (void) theta;
  memcpy(thetastar, &a->thetastar[(size_t) k * a->nparams], sizeof(double) * a->nparams);
  if (!(std::isfinite(a->m2lnp_star[k]) && std::isfinite(a->m2lnp_cur[k])))
    ncm_b200_error("_ncm_fit_esmcmc_walker_apes_step: assertion failed (gsl_finite (m2lnapes_star) && gsl_finite (m2lnapes_cur))");
}

gdouble ncm_fit_esmcmc_walker_apes_prob_norm(NcmFitESMCMCWalkerAPES *a, guint k) {
  if (a->exploration) return 0.0;
  return -0.5 * (a->m2lnp_cur[k] - a->m2lnp_star[k]);
}
const gdouble *ncm_fit_esmcmc_walker_apes_peek_thetastar(NcmFitESMCMCWalkerAPES *a) { return a->thetastar.data(); }
const gdouble *ncm_fit_esmcmc_walker_apes_peek_m2lnp_star(NcmFitESMCMCWalkerAPES *a) { return a->m2lnp_star.data(); }
const gdouble *ncm_fit_esmcmc_walker_apes_peek_m2lnp_cur(NcmFitESMCMCWalkerAPES *a) { return a->m2lnp_cur.data(); }

void ncm_b200_esmcmc_run(NcmFitESMCMCWalkerAPES *a, NcmB200M2lnLFunc m2lnL_func, void *user_data, const gdouble *lb, const gdouble *ub,
                         gdouble *theta, gdouble *m2lnL, guint iters, NcmRNG *rng, unsigned char *accepted, gdouble *timers_ms) {
  const guint d = a->nparams, W = a->size, W2 = a->size_2;
  std::vector<double> m2lnL_star(W2), thetastar(d);
  double t_like = 0.0;
  const double t_begin = now_ms();
  a->t_sample_ms = a->t_eval_ms = 0.0;
  ncm_b200_error_clear();
  double s0[NCM_SD_GPU_T_LEN] = {0}, s1[NCM_SD_GPU_T_LEN] = {0}, sh0 = 0.0, sh1 = 0.0;
  long long sn0 = 0, sn1 = 0;
  ncm_stats_dist_b200_get_timers(a->sd0, s0, &sn0, &sh0);   // snapshot: the context timers are cumulative
  ncm_stats_dist_b200_get_timers(a->sd1, s1, &sn1, &sh1);
  for (guint it = 0; it < iters; it++) {
    unsigned char *acc = accepted != nullptr ? &accepted[(size_t) it * W] : nullptr;
    if (acc != nullptr) memset(acc, 0, W);
    for (guint k = 0; k < W; k++) a->jumps[k] = ncm_rng_uniform01_gen(rng);   // _ncm_fit_esmcmc_get_jumps (0, W)
    for (int block = 0; block < 2; block++) {
      const guint ki = block == 0 ? 0 : W2, kf = block == 0 ? W2 : W;
      ncm_fit_esmcmc_walker_apes_setup(a, lb, ub, theta, m2lnL, ki, kf, rng);
      if (ncm_b200_error_pending()) return;
      const double t0 = now_ms();
      m2lnL_func(&a->thetastar[(size_t) ki * d], kf - ki, d, m2lnL_star.data(), user_data);
      for (guint k = ki; k < kf; k++) {
        ncm_fit_esmcmc_walker_apes_step(a, theta, thetastar.data(), k);
        double prob = 0.0;
        if (valid_bounds(lb, ub, thetastar.data(), d)) {
          const double ms = m2lnL_star[k - ki];
          if (std::isfinite(ms)) {
            const double lnq   = ncm_fit_esmcmc_walker_apes_prob_norm(a, k);
            const double m2lnq = -2.0 * lnq;
            const double m2lnp = ms - m2lnL[k] + m2lnq;
            prob               = exp(-0.5 * m2lnp);
            prob               = prob < 1.0 ? prob : 1.0;
          }
          if (a->jumps[k] < prob) {
            memcpy(&theta[(size_t) k * d], thetastar.data(), sizeof(double) * d);
            m2lnL[k] = ms;
            if (acc != nullptr) acc[k] = 1;
          }
        }
      }
      t_like += now_ms() - t0;
    }
  }
  if (timers_ms != nullptr) {
    double g0[NCM_SD_GPU_T_LEN] = {0}, g1[NCM_SD_GPU_T_LEN] = {0}, h0 = 0.0, h1 = 0.0;
    long long n0 = 0, n1 = 0;
    ncm_stats_dist_b200_get_timers(a->sd0, g0, &n0, &h0);
    ncm_stats_dist_b200_get_timers(a->sd1, g1, &n1, &h1);
    for (int i = 0; i < NCM_SD_GPU_T_LEN; i++) {
      g0[i] -= s0[i];
      g1[i] -= s1[i];
    }
    h0 -= sh0;
    h1 -= sh1;
    timers_ms[0] = h0 + h1;
    timers_ms[1] = g0[NCM_SD_GPU_T_IM] + g1[NCM_SD_GPU_T_IM];
    timers_ms[2] = g0[NCM_SD_GPU_T_SYRK] + g1[NCM_SD_GPU_T_SYRK] + g0[NCM_SD_GPU_T_CHOL] + g1[NCM_SD_GPU_T_CHOL] + g0[NCM_SD_GPU_T_NNLS_MISC] +
                   g1[NCM_SD_GPU_T_NNLS_MISC] + g0[NCM_SD_GPU_T_LOWRANK] + g1[NCM_SD_GPU_T_LOWRANK];
    timers_ms[3] = a->t_sample_ms;
    timers_ms[4] = a->t_eval_ms;
    timers_ms[5] = t_like;
    timers_ms[6] = g0[NCM_SD_GPU_T_H2D] + g1[NCM_SD_GPU_T_H2D] + g0[NCM_SD_GPU_T_D2H] + g1[NCM_SD_GPU_T_D2H];
    timers_ms[7] = now_ms() - t_begin;
  }
}

void ncm_b200_target_rosenbrock(const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *) {
  for (guint i = 0; i < n; i++) {
    const double x1 = X[(size_t) i * nparams], x2 = X[(size_t) i * nparams + 1];
    const double a = x2 - x1 * x1, b = 1.0 - x1;
    m2lnL[i]       = (100.0 * (a * a) + (b * b)) * 1.0e-1;
  }
}

void ncm_b200_target_funnel(const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *) {
  for (guint i = 0; i < n; i++) {
    const double *x       = &X[(size_t) i * nparams];
    const double nu       = x[0];
    const double sigma_nu = exp(0.5 * nu);
    const guint x_len     = nparams - 1;
    double v              = x_len * nu + (nu / 3.0) * (nu / 3.0);
    for (guint j = 0; j < x_len; j++) {
      const double r = x[1 + j] / sigma_nu;
      v += r * r;
    }
    m2lnL[i] = v;
  }
}

void ncm_b200_target_mvnd(const gdouble *X, guint n, guint nparams, gdouble *m2lnL, void *user_data) {
  const NcmB200MVND *t = (const NcmB200MVND *) user_data;
  const guint d        = nparams;
  double v[NCM_SD_GPU_MAX_DIM];
  for (guint i = 0; i < n; i++) {
    const double *x = &X[(size_t) i * d];
    double s        = 0.0;
    for (guint k = 0; k < d; k++) {
      double acc = x[k] - t->mu[k];
      for (guint j = 0; j < k; j++) acc -= t->U[j * d + k] * v[j];
      v[k] = acc / t->U[k * d + k];
      s += v[k] * v[k];
    }
    m2lnL[i] = s;
  }
}

}   // extern "C"
