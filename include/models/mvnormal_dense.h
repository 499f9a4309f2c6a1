/* mvnormal_dense.h — a correlated multivariate normal N(mu, Sigma) as a USER model (include/dhmc_models.h):
 *   l(q) = -1/2 (q - mu)' P (q - mu),   grad = -P (q - mu),   P = Sigma^{-1}
 * the target family of the reference's sample-correctness tests (test/sample-correctness_tests.jl: random, ill-conditioned
 * and "kept" multivariate normals with dense adaptation), which the shipped DIAG_NORMAL family cannot express.  Every
 * element evaluates its own row of P (q - mu) — O(D^2) per gradient, meant for small D.
 * params = [mu (D), P row-major (D*D)]. */
#define DHMC_USER_NAME "mvnormal_dense"
#define DHMC_USER_NSUMS 1      /* S[0] = (q - mu)' P (q - mu) */

DHMC_HD double dhmc_mvn_row(int i, int D, const double* q, const double* params) {   /* (P (q - mu))_i, sequential over j */
  const double* mu = params;
  const double* row = params + D + (long)i * D;
  double acc = 0.0;
  for (int j = 0; j < D; ++j) acc = acc + row[j] * (q[j] - mu[j]);
  return acc;
}
DHMC_HD void dhmc_user_terms(int i, int D, const double* q, const double* params, double* t) {
  t[0] = (q[i] - params[i]) * dhmc_mvn_row(i, D, q, params);
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)D; (void)q; (void)params;
  return -0.5 * S[0];
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  (void)S;
  return -dhmc_mvn_row(i, D, q, params);
}
