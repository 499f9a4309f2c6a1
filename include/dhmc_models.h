/* dhmc_models.h — scalar formulas of the shipped device log-density family.
 *
 * These play the role of the USER's `LogDensityProblems.logdensity_and_gradient`
 * (reference call site src/hamiltonian.jl:204).  They are inputs to the sampler,
 * not part of it, so the CUDA path and the CPU oracle share the per-element
 * formulas below; every cross-element sum is done by each side's own reduction
 * in the canonical order documented in DESIGN.md §"canonical reduction".
 *
 * Families:
 *   STD_NORMAL  l(q) = -1/2 sum q_i^2                      grad = -q
 *   DIAG_NORMAL l(q) = -1/2 sum (q_i-mu_i)^2 * prec_i      grad = -(q_i-mu_i)*prec_i
 *               params = [mu(D), prec(D)]
 *   FUNNEL      Neal's funnel, theta = (v, x_1..x_{D-1}):
 *               l = -v^2/18 - 1/2 e^{-v} sum x_i^2 - (D-1)/2 v
 *   LOGISTIC    logistic regression with a N(0, I) prior, X [N x p], y in {0,1}:
 *               eta = X beta;  l = sum_n [y_n eta_n - log1pexp(eta_n)] - 1/2 |beta|^2
 *               grad = X' (y - sigma(eta)) - beta
 *               params = [N, X row-major (N*p), y (N)];  eta_n and (X' r)_j are sequential fused
 *               multiply-adds over j resp. n, the two scalar sums use the canonical reduction.
 */
#ifndef DHMC_MODELS_H
#define DHMC_MODELS_H
#include "dhmc_math.h"

enum {
  DHMC_FAMILY_STD_NORMAL = 0,
  DHMC_FAMILY_DIAG_NORMAL = 1,
  DHMC_FAMILY_FUNNEL = 2,
  DHMC_FAMILY_LOGISTIC = 3,
  DHMC_FAMILY_COUNT = 4
};

/* --- STD_NORMAL: term of the sum and gradient element */
DHMC_HD double dhmc_std_term(double q) { return q * q; }
DHMC_HD double dhmc_std_grad(double q) { return -q; }
DHMC_HD double dhmc_std_lq(double sum) { return -0.5 * sum; }

/* --- DIAG_NORMAL */
DHMC_HD double dhmc_diag_scaled(double q, double mu, double prec) {
  return prec * (q - mu);
}
DHMC_HD double dhmc_diag_term(double q, double mu, double t) {
  return (q - mu) * t;
}
DHMC_HD double dhmc_diag_grad(double t) { return -t; }
DHMC_HD double dhmc_diag_lq(double sum) { return -0.5 * sum; }

/* --- FUNNEL: S = sum_{i>=1} x_i^2 (element 0 contributes +0.0) */
DHMC_HD double dhmc_funnel_term(int i, double x) { return i == 0 ? 0.0 : x * x; }
DHMC_HD double dhmc_funnel_lq(double v, double ev, double S, int D) {
  double a = (v * v) / 18.0;
  double b = (0.5 * ev) * S;
  double c = (0.5 * (double)(D - 1)) * v;
  return ((-a) - b) - c;
}
DHMC_HD double dhmc_funnel_grad(int i, double x, double v, double ev, double S,
                                int D) {
  if (i == 0) return ((-v) / 9.0 + (0.5 * ev) * S) - 0.5 * (double)(D - 1);
  return -(ev * x);
}

/* --- LOGISTIC */
/* one multiply-accumulate of η_n = Σ_j X_nj β_j and of (Xᵀr)_j = Σ_n X_nj r_n: fused, as a BLAS would */
DHMC_HD double dhmc_logit_mac(double acc, double x, double b) { return dm_fma(x, b, acc); }
DHMC_HD double dhmc_logit_sigma(double eta) { return 1.0 / (1.0 + dm_exp(-eta)); }
/* y·η − log(1+e^η) with log(1+e^η) = max(η,0) + log(1+e^{−|η|}): one table-driven softplus (absolute
 * accuracy, no division) instead of log1p(exp(η)) */
DHMC_HD double dhmc_logit_ll(double y, double eta) {
  return y * eta - ((eta > 0.0 ? eta : 0.0) + dm_softplus_neg(dm_fabs(eta)));
}
DHMC_HD double dhmc_logit_resid(double y, double eta) { return y - dhmc_logit_sigma(eta); }
DHMC_HD double dhmc_logit_lq(double sum_ll, double sum_b2) { return sum_ll - 0.5 * sum_b2; }
DHMC_HD double dhmc_logit_grad(double xtr, double beta) { return xtr - beta; }

#endif
