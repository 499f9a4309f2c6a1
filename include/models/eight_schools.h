/* eight_schools.h — non-centred hierarchical normal ("eight schools" with J groups) as a USER model
 * (include/dhmc_models.h):  theta = (mu, log tau, eta_1..eta_J), D = J + 2,
 *   y_j ~ N(mu + tau eta_j, sigma_j),  eta_j ~ N(0, 1),  mu ~ N(0, 5),  tau ~ HalfCauchy(5)  (sampled as log tau)
 *   l = -1/2 sum r_j^2 - 1/2 sum eta_j^2 - mu^2/50 - log(1 + tau^2/25) + log tau,   r_j = (y_j - mu - tau eta_j)/sigma_j
 * params = [y_1..y_J, sigma_1..sigma_J]. */
#define DHMC_USER_NAME "eight_schools"
#define DHMC_USER_NSUMS 4      /* S0 = sum r^2, S1 = sum eta^2, S2 = sum r/sigma, S3 = sum r eta/sigma */
#define DHMC_USER_NSCALARS 1   /* S4 = tau */
#define DHMC_USER_MIN_DIM 3

DHMC_HD void dhmc_user_terms(int i, int D, const double* q, const double* params, double* t) {
  if (i < 2) { t[0] = t[1] = t[2] = t[3] = 0.0; return; }
  const int J = D - 2, j = i - 2;
  const double tau = dm_exp(q[1]);
  const double sig = params[J + j];
  const double r = ((params[j] - q[0]) - tau * q[i]) / sig;
  t[0] = r * r;
  t[1] = q[i] * q[i];
  t[2] = r / sig;
  t[3] = (r * q[i]) / sig;
}
DHMC_HD void dhmc_user_prepare(int D, const double* q, double* S, const double* params) {
  (void)D; (void)params;
  S[4] = dm_exp(q[1]);
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)D; (void)params;
  const double tau = S[4];
  return (((-0.5 * S[0]) - 0.5 * S[1]) - (q[0] * q[0]) / 50.0) - dm_log1p((tau * tau) / 25.0) + q[1];
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  const double tau = S[4];
  if (i == 0) return S[2] - q[0] / 25.0;
  if (i == 1) return (tau * S[3] - (2.0 * (tau * tau)) / (25.0 + tau * tau)) + 1.0;
  const int J = D - 2, j = i - 2;
  const double sig = params[J + j];
  const double r = ((params[j] - q[0]) - tau * q[i]) / sig;
  return (tau * r) / sig - q[i];
}
