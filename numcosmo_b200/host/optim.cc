// One-parameter optimisers behind the cross-validation modes of NcmStatsDist.
//
// The reference tunes a single scalar, ln(over_smooth), in two ways:
//   * CV_SPLIT_NOFIT / CV_LOO: gsl_multimin_fminimizer_nmsimplex2 allocated for ONE parameter
//     (ncm_stats_dist.c:175) and driven by _ncm_stats_dist_minimize_obj (ncm_stats_dist.c:660-701);
//   * CV_SPLIT: levmar's dlevmar_dif with m = 1 (ncm_stats_dist.c:1062-1064, numcosmo/external/levmar/lm_core.c:436-851).
// Both are written here for exactly that case -- a two-corner simplex on the real line and a scalar
// Levenberg-Marquardt with a forward-difference / secant-updated derivative -- keeping the operation order of the
// general algorithms so that the sequence of trial points (each one a full GPU IM + NNLS + eval pass) is the one
// the reference would visit.
#include <cfloat>
#include <cmath>
#include <vector>
#include "internal.h"

// ---------------------------------------------------------------------------------------------------------------
// Nelder-Mead on the line: corners a[0], a[1] with values y[0], y[1]; c = running centre, S2 = running squared size
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Simplex1 {
  double a[2], y[2], c, S2;
  const std::function<double(double)> &f;
  explicit Simplex1(const std::function<double(double)> &fn) : f(fn) {}

  void recentre() {
    double s = 0.0;
    s += 1.0 * a[0];
    s += 1.0 * a[1];
    c = s * (1.0 / 2);
  }
  double resize() {
    double ss = 0.0;
    for (int k = 0; k < 2; k++) {
      const double t = std::fabs(a[k] + (-1.0) * c);
      ss += t * t;
    }
    S2 = ss / 2;
    return std::sqrt(ss / 2);
  }
  // trial point: the corner moved through the other one by `coeff` (-1 mirror, -2 mirror and stretch, 0.5 pull in)
  double trial(double coeff, int corner, double &xc) {
    const double alpha = (1 - coeff) * 2 / (2 - 1.0);
    const double beta  = (2 * coeff - 1.0) / (2 - 1.0);
    xc = c;
    xc *= alpha;
    xc += beta * a[corner];
    return f(xc);
  }
  void replace(int k, double x, double val) {
    const double delta = x + (-1.0) * a[k];
    const double xmc   = a[k] + (-1.0) * c;
    const double dn    = std::fabs(delta);
    S2 += (2.0 / 2) * (xmc * delta) + ((2 - 1.0) / 2) * (dn * dn / 2);
    const double alpha = 1.0 / 2;
    c += (-alpha) * a[k];
    c += alpha * x;
    a[k] = x;
    y[k] = val;
  }
};

}   // namespace

int ncm_b200_simplex1_minimize(const std::function<double(double)> &f, double x0, double step, double size_tol, int max_iter, double *x_best,
                               double *f_best) {
  Simplex1 s(f);
  s.a[0] = x0;
  s.y[0] = f(x0);
  if (!std::isfinite(s.y[0])) return -1;
  s.a[1] = x0 + step;
  s.y[1] = f(s.a[1]);
  if (!std::isfinite(s.y[1])) return -1;
  s.recentre();
  double size = s.resize();

  int iter = 0;
  for (;;) {
    iter++;
    // with two corners the scan leaves "second highest" on corner 1 when corner 1 is the lowest, on corner 0 otherwise
    int hi = 0, lo = 0, s_hi = 1;
    if (s.y[1] < s.y[0]) {
      lo = 1;
    } else if (s.y[1] > s.y[0]) {
      s_hi = 0;
      hi   = 1;
    }
    double xc, xc2;
    const double val = s.trial(-1.0, hi, xc);
    bool failed      = false;
    if (std::isfinite(val) && val < s.y[lo]) {
      const double val2 = s.trial(-2.0, hi, xc2);
      if (std::isfinite(val2) && val2 < s.y[lo])
        s.replace(hi, xc2, val2);
      else
        s.replace(hi, xc, val);
    } else if (!std::isfinite(val) || val > s.y[s_hi]) {
      if (std::isfinite(val) && val <= s.y[hi]) s.replace(hi, xc, val);
      const double val2 = s.trial(0.5, hi, xc2);
      if (std::isfinite(val2) && val2 <= s.y[hi]) {
        s.replace(hi, xc2, val2);
      } else {
        // shrink towards the best corner
        const int other = 1 - lo;
        s.a[other]      = 0.5 * (s.a[other] + s.a[lo]);
        s.y[other]      = f(s.a[other]);
        if (!std::isfinite(s.y[other])) failed = true;
        s.recentre();
        s.resize();
      }
    } else {
      s.replace(hi, xc, val);
    }
    if (failed) break;
    const int best = (s.y[1] < s.y[0]) ? 1 : 0;
    *x_best        = s.a[best];
    *f_best        = s.y[best];
    size           = (s.S2 > 0) ? std::sqrt(s.S2) : s.resize();
    if (size < size_tol) break;
    if (!(iter < max_iter)) break;
  }
  return iter;
}

// ---------------------------------------------------------------------------------------------------------------
// Scalar Levenberg-Marquardt, residual model hx(p) in R^n, target x (nullptr = 0), derivative by forward differences
// refreshed by secant updates
// ---------------------------------------------------------------------------------------------------------------
namespace {

// e = x - y, ||e||^2 in levmar's four interleaved accumulators (blocks of 8 from the top, then the tail)
double sq_err(std::vector<double> &e, const double *x, const std::vector<double> &y) {
  const int n = (int) y.size(), blockn = (n >> 3) << 3;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int top = blockn - 1; top > 0; top -= 8)
    for (int k = 0; k < 8; k++) {
      const int j = top - k;
      e[j]        = (x != nullptr) ? x[j] - y[j] : -y[j];
      acc[k & 3] += e[j] * e[j];
    }
  for (int j = blockn; j < n; j++) {
    e[j] = (x != nullptr) ? x[j] - y[j] : -y[j];
    acc[(7 - (n - j)) & 3] += e[j] * e[j];
  }
  return acc[0] + acc[1] + acc[2] + acc[3];
}

}   // namespace

int ncm_b200_lm1_dif(const std::function<void(double, double *)> &func, double *p_io, const double *x, int n, int itmax, const double opts[5],
                     double info[10]) {
  if (n < 1) return -1;
  const double tau = opts[0], eps1 = opts[1], eps2 = opts[2], eps2_sq = opts[2] * opts[2], eps3 = opts[3];
  const double delta = std::fabs(opts[4]);
  const bool forward = !(opts[4] < 0.0);
  const int K        = 10;   // secant updates between two finite-difference derivatives
  std::vector<double> e(n), hx(n), J(n), trial(n), e_trial(n);
  double p = *p_io, mu = 0.0, JtJ = 0.0, JtJ_aug = 0.0, Jte = 0.0, Jte_inf = 0.0, p_L2 = 0.0, Dp_L2 = DBL_MAX;
  int nu = 20, stop = 0, nfev, njap = 0, nlss = 0, updjac = 0, updp = 1, newjac = 0, k;

  func(p, hx.data());
  nfev               = 1;
  double p_eL2       = sq_err(e, x, hx);
  const double init2 = p_eL2;
  if (!std::isfinite(p_eL2)) stop = 7;

  for (k = 0; k < itmax && !stop; k++) {
    if (p_eL2 <= eps3) {
      stop = 6;
      break;
    }
    if ((updp && nu > 16) || updjac == K) {
      double h = std::fabs(1e-04 * p);
      if (h < delta) h = delta;
      if (forward) {
        func(p + h, trial.data());
        const double ih = 1.0 / h;
        for (int i = 0; i < n; i++) J[i] = (trial[i] - hx[i]) * ih;
        nfev += 1;
      } else {
        func(p - h, trial.data());
        func(p + h, e_trial.data());
        const double ih = 0.5 / h;
        for (int i = 0; i < n; i++) J[i] = (e_trial[i] - trial[i]) * ih;
        nfev += 2;
      }
      njap++;
      nu     = 2;
      updjac = 0;
      updp   = 0;
      newjac = 1;
    }
    if (newjac) {
      newjac = 0;
      JtJ = Jte = 0.0;
      if (n <= 32 * 32) {
        for (int l = n; l-- > 0;) {
          JtJ += J[l] * J[l];
          Jte += J[l] * e[l];
        }
      } else {
        for (int kk = 0; kk < n; kk += 32) {
          const int kend = (kk + 32 <= n) ? kk + 32 : n;
          double part    = 0.0;
          for (int l = kk; l < kend; l++) part += J[l] * J[l];
          JtJ += part;
        }
        for (int l = 0; l < n; l++) Jte += J[l] * e[l];
      }
      Jte_inf = std::fabs(Jte);
      if (!(0.0 < Jte_inf)) Jte_inf = 0.0;
      p_L2    = p * p;
      JtJ_aug = JtJ;
    }
    if (Jte_inf <= eps1) {
      Dp_L2 = 0.0;
      stop  = 1;
      break;
    }
    if (k == 0) mu = tau * JtJ;
    JtJ_aug += mu;
    nlss++;
    if (JtJ_aug != 0.0) {
      const double Dp  = Jte / JtJ_aug;
      const double pDp = p + Dp;
      Dp_L2            = Dp * Dp;
      if (Dp_L2 <= eps2_sq * p_L2) {
        stop = 2;
        break;
      }
      if (Dp_L2 >= (p_L2 + eps2) / (1e-12 * 1e-12)) {
        stop = 4;
        break;
      }
      func(pDp, trial.data());
      nfev++;
      const double pDp_eL2 = sq_err(e_trial, x, trial);
      if (!std::isfinite(pDp_eL2)) {
        stop = 7;
        break;
      }
      const double dF = p_eL2 - pDp_eL2;
      if (updp || dF > 0) {
        for (int i = 0; i < n; i++) {
          double t = 0.0;
          t += J[i] * Dp;
          t = (trial[i] - hx[i] - t) / Dp_L2;
          J[i] += t * Dp;
        }
        updjac++;
        newjac = 1;
      }
      double dL = 0.0;
      dL += Dp * (mu * Dp + Jte);
      if (dL > 0.0 && dF > 0.0) {
        double t = (2.0 * dF / dL - 1.0);
        t        = 1.0 - t * t * t;
        mu       = mu * ((t >= 0.3333333334) ? t : 0.3333333334);
        nu       = 2;
        p        = pDp;
        e.swap(e_trial);
        hx.swap(trial);
        p_eL2 = pDp_eL2;
        updp  = 1;
        continue;
      }
    }
    mu *= nu;
    const int nu2 = nu << 1;
    if (nu2 <= nu) {
      stop = 5;
      break;
    }
    nu      = nu2;
    JtJ_aug = JtJ;
  }
  if (k >= itmax) stop = 3;
  if (info != nullptr) {
    info[0] = init2;
    info[1] = p_eL2;
    info[2] = Jte_inf;
    info[3] = Dp_L2;
    info[4] = mu / JtJ;
    info[5] = (double) k;
    info[6] = (double) stop;
    info[7] = (double) nfev;
    info[8] = (double) njap;
    info[9] = (double) nlss;
  }
  *p_io = p;
  return (stop != 4 && stop != 7) ? k : -1;
}

// plain-C faces for tests (ctypes): a C callback instead of std::function
extern "C" {

int ncm_b200_test_simplex1(double (*f)(double, void *), void *data, double x0, double step, double size_tol, int max_iter, double *x_best, double *f_best) {
  std::function<double(double)> fn = [&](double x) { return f(x, data); };
  return ncm_b200_simplex1_minimize(fn, x0, step, size_tol, max_iter, x_best, f_best);
}

int ncm_b200_test_lm1_dif(void (*f)(double, double *, int, void *), void *data, double *p, const double *x, int n, int itmax, const double *opts,
                          double *info) {
  std::function<void(double, double *)> fn = [&](double pp, double *hx) { f(pp, hx, n, data); };
  return ncm_b200_lm1_dif(fn, p, x, n, itmax, opts, info);
}
}
