/* mixture_normals.h — a two-component normal mixture as a USER model (include/dhmc_models.h):
 *   l(q) = log( alpha N(q; 0, I) + (1 - alpha) N(q; mu, Sigma) )      (up to the common -D/2 log 2 pi)
 * the "mixture of two normals" target of the reference's sample-correctness tests (test/sample-correctness_tests.jl:93-102).
 *   a = log alpha - 1/2 q'q,   b = log(1 - alpha) - 1/2 (q-mu)' P (q-mu) - log det L,   l = logaddexp(a, b)
 *   grad = -w_a q - w_b P (q - mu),   w_a = e^{a - l}, w_b = e^{b - l}
 * params = [alpha, log det L, mu (D), P row-major (D*D)]   (Sigma = L L', P = Sigma^{-1}). */
#define DHMC_USER_NAME "mixture_normals"
#define DHMC_USER_NSUMS 2      /* S0 = q'q, S1 = (q - mu)' P (q - mu) */
#define DHMC_USER_NSCALARS 3   /* S2 = l, S3 = w_a, S4 = w_b */

DHMC_HD double dhmc_mix_row(int i, int D, const double* q, const double* params) {   /* (P (q - mu))_i */
  const double* mu = params + 2;
  const double* row = params + 2 + D + (long)i * D;
  double acc = 0.0;
  for (int j = 0; j < D; ++j) acc = acc + row[j] * (q[j] - mu[j]);
  return acc;
}
DHMC_HD void dhmc_user_terms(int i, int D, const double* q, const double* params, double* t) {
  t[0] = q[i] * q[i];
  t[1] = (q[i] - params[2 + i]) * dhmc_mix_row(i, D, q, params);
}
DHMC_HD void dhmc_user_prepare(int D, const double* q, double* S, const double* params) {
  (void)D; (void)q;
  const double a = dm_log(params[0]) - 0.5 * S[0];
  const double b = (dm_log(1.0 - params[0]) - 0.5 * S[1]) - params[1];
  const double l = dm_logaddexp(a, b);
  S[2] = l;
  S[3] = dm_exp(a - l);
  S[4] = dm_exp(b - l);
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)D; (void)q; (void)params;
  return S[2];
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  return (-(S[3] * q[i])) - S[4] * dhmc_mix_row(i, D, q, params);
}
