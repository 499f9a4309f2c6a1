/* funnel_user.h — Neal's funnel written against the USER model contract (include/dhmc_models.h).
 *
 * Same formulas as the shipped FUNNEL family (dhmc_funnel_*), so a library built from this header must reproduce
 * family 2 bit for bit — the test that pins the user-model path to a shipped one (tests/test_user_model.py).
 *   theta = (v, x_1..x_{D-1});  l = -v^2/18 - 1/2 e^{-v} sum x_i^2 - (D-1)/2 v
 * params: none. */
#define DHMC_USER_NAME "funnel_user"
#define DHMC_USER_NSUMS 1      /* S[0] = sum_{i>=1} x_i^2 */
#define DHMC_USER_NSCALARS 1   /* S[1] = e^{-v} */
#define DHMC_USER_MIN_DIM 2

DHMC_HD void dhmc_user_terms(int i, int D, const double* q, const double* params, double* t) {
  (void)D; (void)params;
  t[0] = dhmc_funnel_term(i, q[i]);
}
DHMC_HD void dhmc_user_prepare(int D, const double* q, double* S, const double* params) {
  (void)D; (void)params;
  S[1] = dm_exp(-q[0]);
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)params;
  return dhmc_funnel_lq(q[0], S[1], S[0], D);
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  (void)params;
  return dhmc_funnel_grad(i, q[i], q[0], S[1], S[0], D);
}
