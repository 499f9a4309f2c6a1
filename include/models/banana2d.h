/* banana2d.h — the two-dimensional "banana" (twisted normal) as a USER model with NO cross-element sums
 * (include/dhmc_models.h; DHMC_USER_NSUMS 0: l is computed from the position directly):
 *   l(q) = -1/2 [ q0^2 / s^2 + (q1 + b q0^2 - b s^2)^2 ],   params = [s, b]        (Haario et al. 1999)
 *   dl/dq0 = -q0 / s^2 - 2 b q0 (q1 + b q0^2 - b s^2),   dl/dq1 = -(q1 + b q0^2 - b s^2) */
#define DHMC_USER_NAME "banana2d"
#define DHMC_USER_NSUMS 0
#define DHMC_USER_MIN_DIM 2

DHMC_HD double dhmc_banana_u(const double* q, const double* params) {
  const double s = params[0], b = params[1];
  return (q[1] + b * (q[0] * q[0])) - b * (s * s);
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)D; (void)S;
  const double s = params[0], u = dhmc_banana_u(q, params);
  return -0.5 * ((q[0] * q[0]) / (s * s) + u * u);
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  (void)D; (void)S;
  const double s = params[0], b = params[1], u = dhmc_banana_u(q, params);
  if (i == 0) return (-(q[0] / (s * s))) - ((2.0 * b) * q[0]) * u;
  if (i == 1) return -u;
  return 0.0;        /* further coordinates (D > 2) are improper-flat: use D = 2 */
}
