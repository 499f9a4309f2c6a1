/* std_normal_user.h — the standard multivariate normal written against the USER model contract (include/dhmc_models.h):
 *   l(q) = -1/2 sum q_i^2,  grad = -q.
 * Same formulas as the shipped STD_NORMAL family (dhmc_std_*), so a library built from this header reproduces family 0 bit
 * for bit; bench.py times it next to the shipped kernels on the C2 workload to show what the generic path (position staged
 * in shared memory, formulas behind the contract) costs.  params: none. */
#define DHMC_USER_NAME "std_normal_user"
#define DHMC_USER_NSUMS 1      /* S[0] = sum q_i^2 */

DHMC_HD void dhmc_user_terms(int i, int D, const double* q, const double* params, double* t) {
  (void)D; (void)params;
  t[0] = dhmc_std_term(q[i]);
}
DHMC_HD double dhmc_user_logdensity(int D, const double* q, const double* S, const double* params) {
  (void)D; (void)q; (void)params;
  return dhmc_std_lq(S[0]);
}
DHMC_HD double dhmc_user_grad(int i, int D, const double* q, const double* S, const double* params) {
  (void)D; (void)S; (void)params;
  return dhmc_std_grad(q[i]);
}
