/* rosenbrock.h — a neighbour-coupled ("banana chain") density as a USER model (include/dhmc_models.h):
 *   l(q) = - sum_{i=0}^{D-2} [ b (q_{i+1} - q_i^2)^2 + (a - q_i)^2 ]
 * Every gradient element looks at its two neighbours, which is what the whole-position view of the contract is for.
 * params = [a, b]. */
#define DHMC_USER_NAME "rosenbrock"
#define DHMC_USER_NSUMS 1
#define DHMC_USER_MIN_DIM 2

DHMC_HD void dhmc_user_terms(int i, int D, const double* q, const double* params, double* t) {
  if (i >= D - 1) { t[0] = 0.0; return; }
  const double a = params[0], b = params[1];
  const double u = q[i + 1] - q[i] * q[i];
  const double w = a - q[i];
  t[0] = b * (u * u) + w * w;
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)D; (void)q; (void)params;
  return -S[0];
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  (void)S;
  const double a = params[0], b = params[1];
  double g = 0.0;
  if (i < D - 1) {                    /* d/dq_i of term i */
    const double u = q[i + 1] - q[i] * q[i];
    g = (4.0 * b) * (q[i] * u) + 2.0 * (a - q[i]);
  }
  if (i > 0) {                        /* d/dq_i of term i-1 */
    const double u = q[i] - q[i - 1] * q[i - 1];
    g = g - (2.0 * b) * u;
  }
  return g;
}
