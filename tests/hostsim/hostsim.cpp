// hostsim.cpp — host simulation of the DEVICE state machine, for `-m "not gpu"` tests.
//
// Compiles the product's control code (dynamichmc.jl_b200/csrc/nuts_machine.cuh,
// the flattened NUTS tree the GPU runs) with g++ against a plain-loop backend so
// that its bookkeeping (explicit stack, slot pool, RNG order, early exits) can
// be compared bit-for-bit with the recursive oracle on a machine without a GPU.
// This is a TEST HARNESS: it is not linked into libdhmc_b200.so and the product
// has no CPU fallback.
#include <cstring>
#include <set>
#include <vector>

#include "../../dynamichmc.jl_b200/csrc/nuts_machine.cuh"
#include "../../include/dhmc_models.h"

using vec = std::vector<double>;
using namespace dhmc;

namespace {

template <class F>
double canon_sum(int T, int D, F term) {
  double part[1024];
  for (int v = 0; v < T; ++v) {
    double acc = 0.0;
    for (int i = v; i < D; i += T) acc = acc + term(i);
    part[v] = acc;
  }
  for (int base = 0; base < T; base += 32)
    for (int off = 16; off >= 1; off >>= 1)
      for (int v = base; v < base + off; ++v) part[v] = part[v] + part[v + off];
  for (int off = 32; off < T; off <<= 1)
    for (int v = 0; v < T; v += 2 * off) part[v] = part[v] + part[v + off];
  return part[0];
}

struct HostBackend {
  int D, T, family;
  const double* params;
  vec q, p, g, minv, rhoL, Rnew;
  double lq = 0;
  std::vector<vec> slots;
  Entry entries[kMaxLevels + 2];
  vec wmean, wm2;
  int max_slots_used = 0;

  HostBackend(int D_, int T_, int fam, const double* pr, int nslots)
      : D(D_), T(T_), family(fam), params(pr), q(D_), p(D_), g(D_), minv(D_, 1.0), rhoL(D_),
        Rnew(D_), slots(nslots, vec(D_)), wmean(D_), wm2(D_) {}

  TopState top_;
  TopState& top() { return top_; }
  void top_sync() {}
  static constexpr bool kDeep = true;              // exercise the spill words of the slot pool on the host as well
  int reserved_first() const { return 1 << 20; }   // no reserved slots in the host simulation
  double cur_lq() const { return lq; }
  void set_cur_lq(double v) { lq = v; }
  void st_q(int s) { slots[s] = q; }
  void st_p(int s) { slots[s] = p; }
  void st_g(int s) { slots[s] = g; }
  void st_rho(int s) { slots[s] = rhoL; }
  void ld_q(int s) { q = slots[s]; }
  void ld_p(int s) { p = slots[s]; }
  void ld_g(int s) { g = slots[s]; }
  void swap_cur(int sq, int sp, int sg) { q.swap(slots[sq]); p.swap(slots[sp]); g.swap(slots[sg]); }
  void rho_from_p() { rhoL = p; }
  void rho_commit() { rhoL = Rnew; }
  void logaddexp2(double a0, double b0, double a1, double b1, double* r0, double* r1) {
    *r0 = dm_logaddexp(a0, b0); *r1 = dm_logaddexp(a1, b1);
  }
  double randexp(dm_rng_key key, uint32_t t, uint32_t j) { return dm_randexp(key, t, j); }
  void put_entry(int j, const Entry& e) { entries[j] = e; }
  Entry get_entry(int j) const { return entries[j]; }

  void draw_momentum(dm_rng_key key, uint32_t t, const double* p_override) {
    for (int i = 0; i < D; ++i)
      p[i] = p_override ? p_override[i]
                        : dm_sqrt(1.0 / minv[i]) * dm_normal_elem(key, DHMC_STREAM_P, t, (uint32_t)i);
  }
  void draw_search_momentum(dm_rng_key key, const double* p_override) {
    for (int i = 0; i < D; ++i)
      p[i] = p_override ? p_override[i]
                        : dm_sqrt(1.0 / minv[i]) * dm_normal_elem(key, DHMC_STREAM_PSEARCH, 0, (uint32_t)i);
  }
  double kinetic() const {
    return canon_sum(T, D, [&](int i) { double ps = minv[i] * p[i]; return p[i] * ps; }) / 2.0;
  }
  double phase_logdensity() const {
    if (!dm_isfinite(lq)) return -dm_inf();
    double K = kinetic();
    return lq - (dm_isfinite(K) ? K : dm_inf());
  }
  // model: fills g from q, returns sanitised ℓq (evaluate_ℓ, hamiltonian.jl:202-217)
  double eval_model(bool* gbad) {
    double l = 0;
    switch (family) {
      case DHMC_FAMILY_STD_NORMAL: {
        double s = canon_sum(T, D, [&](int i) { return dhmc_std_term(q[i]); });
        for (int i = 0; i < D; ++i) g[i] = dhmc_std_grad(q[i]);
        l = dhmc_std_lq(s);
        break;
      }
      case DHMC_FAMILY_DIAG_NORMAL: {
        const double* mu = params; const double* pr = params + D;
        vec t(D);
        for (int i = 0; i < D; ++i) t[i] = dhmc_diag_scaled(q[i], mu[i], pr[i]);
        double s = canon_sum(T, D, [&](int i) { return dhmc_diag_term(q[i], mu[i], t[i]); });
        for (int i = 0; i < D; ++i) g[i] = dhmc_diag_grad(t[i]);
        l = dhmc_diag_lq(s);
        break;
      }
      case DHMC_FAMILY_LOGISTIC: {
        const int N = (int)params[0];
        const double* X = params + 1; const double* y = X + (size_t)N * D;
        vec r(N), ll(N);
        for (int n = 0; n < N; ++n) {
          const double eta = dhmc_logit_eta(X + (size_t)n * D, &q[0], D);
          dhmc_logit_ll_resid(y[n], eta, &ll[n], &r[n]);
        }
        double sll = canon_sum(T, N, [&](int n) { return ll[n]; });
        double sb = canon_sum(T, D, [&](int i) { return q[i] * q[i]; });
        for (int j = 0; j < D; ++j) {
          double acc = 0.0;
          for (int n = 0; n < N; ++n) acc = dhmc_logit_mac(acc, X[(size_t)n * D + j], r[n]);
          g[j] = dhmc_logit_grad(acc, q[j]);
        }
        l = dhmc_logit_lq(sll, sb);
        break;
      }
      default: {
        double v = q[0], ev = dm_exp(-v);
        double S = canon_sum(T, D, [&](int i) { return dhmc_funnel_term(i, q[i]); });
        for (int i = 0; i < D; ++i) g[i] = dhmc_funnel_grad(i, q[i], v, ev, S, D);
        l = dhmc_funnel_lq(v, ev, S, D);
      }
    }
    bool bad = false;
    for (int i = 0; i < D; ++i) bad = bad || !dm_isfinite(g[i]);
    *gbad = bad;
    if ((dm_isfinite(l) && !bad) || l == -dm_inf()) return l;
    return -dm_inf();
  }
  double leapfrog(double eps, int* flags) {
    const double h = eps / 2;
    for (int i = 0; i < D; ++i) p[i] = p[i] + h * g[i];
    bool qbad = false;
    for (int i = 0; i < D; ++i) {
      double vel = minv[i] * p[i];
      q[i] = q[i] + eps * vel;
      qbad = qbad || !dm_isfinite(q[i]);
    }
    bool gbad;
    lq = eval_model(&gbad);
    if (qbad) { *flags |= 1; lq = -dm_inf(); }
    for (int i = 0; i < D; ++i) p[i] = p[i] + h * g[i];
    return phase_logdensity();
  }
  bool merge_check(int sEf, int sEl, int sEr, int sLf, bool L_leaf) {
    const vec& Ef = slots[sEf]; const vec& El = slots[sEl]; const vec& Er = slots[sEr];
    const vec& Lf = L_leaf ? p : slots[sLf];
    const vec& Lr = L_leaf ? p : rhoL;
    vec A(D), Bv(D);
    for (int i = 0; i < D; ++i) { A[i] = Er[i] + Lf[i]; Bv[i] = El[i] + Lr[i]; Rnew[i] = Er[i] + Lr[i]; }
    auto dot = [&](const vec& a, const vec& r) {
      return canon_sum(T, D, [&](int i) { double ps = minv[i] * a[i]; return ps * r[i]; });
    };
    double d1 = dot(Ef, A), d2 = dot(Lf, A), d3 = dot(El, Bv), d4 = dot(p, Bv), d5 = dot(Ef, Rnew),
           d6 = dot(p, Rnew);
    return d1 < 0 || d2 < 0 || d3 < 0 || d4 < 0 || d5 < 0 || d6 < 0;
  }
  void metric_reset(int) { std::fill(wmean.begin(), wmean.end(), 0.0); std::fill(wm2.begin(), wm2.end(), 0.0); }
  void metric_push(int, int n) {
    for (int i = 0; i < D; ++i) {
      double d = q[i] - wmean[i];
      wmean[i] = wmean[i] + d / (double)n;
      wm2[i] = wm2[i] + d * (q[i] - wmean[i]);
    }
  }
  void metric_finish(int, int n) { for (int i = 0; i < D; ++i) minv[i] = wm2[i] / (double)(n - 1); }
};

struct Sink {
  HostBackend& b; double* post; dhmc_tree_stats* stats; double* logd; double* eps_used;
  void operator()(int n, const dhmc_tree_stats& ts, double e) {
    if (post) std::memcpy(post + (size_t)n * b.D, b.q.data(), sizeof(double) * b.D);
    if (stats) stats[n] = ts;
    if (logd) logd[n] = b.lq;
    if (eps_used) eps_used[n] = e;
  }
};

}  // namespace

// ---- DummyTrajectory backend (reference test/test_trees.jl:28-103) for the flattened machine:
// positions are integers, a "vector" is the position it belongs to, a tree is turning when all
// of its leaves are in the `turning` set, a leaf is divergent when it is in `divergent`.
struct DummyBackend {
  std::set<long> turning, divergent;
  long z = 0, z0 = 0;
  std::vector<long> slots;
  std::vector<long> visited;
  Entry entries[kMaxLevels + 2];
  TopState top_;
  static double l(long zz) { return -((double)(zz - 3) * (double)(zz - 3)) * 0.1; }   // testℓ, :106
  explicit DummyBackend(int nslots) : slots(nslots, 0) {}
  TopState& top() { return top_; }
  void top_sync() {}
  static constexpr bool kDeep = true;              // exercise the spill words of the slot pool on the host as well
  int reserved_first() const { return 1 << 20; }   // no reserved slots in the host simulation
  double cur_lq() const { return 0.0; }
  void set_cur_lq(double) {}
  void st_q(int s) { slots[s] = z; }
  void st_p(int s) { slots[s] = z; }
  void st_g(int s) { slots[s] = z; }
  void st_rho(int) {}
  void ld_q(int s) { z = slots[s]; }
  void ld_p(int s) { z = slots[s]; }
  void ld_g(int) {}
  void swap_cur(int sq, int sp, int sg) { long t = slots[sq]; slots[sq] = z; slots[sp] = z; slots[sg] = z; z = t; }
  void rho_from_p() {}
  void rho_commit() {}
  void put_entry(int j, const Entry& e) { entries[j] = e; }
  Entry get_entry(int j) const { return entries[j]; }
  void logaddexp2(double a0, double b0, double a1, double b1, double* r0, double* r1) {
    *r0 = dm_logaddexp(a0, b0); *r1 = dm_logaddexp(a1, b1);
  }
  double randexp(dm_rng_key key, uint32_t t, uint32_t j) { return dm_randexp(key, t, j); }
  void draw_momentum(dm_rng_key, uint32_t, const double*) {}
  void draw_search_momentum(dm_rng_key, const double*) {}
  double phase_logdensity() const { return l(z0); }
  double leapfrog(double eps, int*) {                       // move: z ± 1, leaf: Δ = ℓ(z)
    z += eps > 0 ? 1 : -1;
    visited.push_back(z);
    return divergent.count(z) ? -dm_inf() : l(z);
  }
  bool all_turning(long a, long b) const {
    if (a > b) std::swap(a, b);
    for (long k = a; k <= b; ++k) if (!turning.count(k)) return false;
    return true;
  }
  bool merge_check(int sEf, int sEl, int, int sLf, bool L_leaf) {
    const long lf = L_leaf ? z : slots[sLf];
    return all_turning(slots[sEf], slots[sEl]) && all_turning(lf, z);
  }
  void metric_reset(int) {}
  void metric_push(int, int) {}
  void metric_finish(int, int) {}
};

extern "C" {

// The product's flattened tree on the reference's DummyTrajectory: visited order, depth,
// termination and steps for a given direction word.
int hs_dummy_sample_trajectory(long z0, int max_depth, uint32_t flags, const long* turning, int nt,
                               const long* divergent, int nd, long* visited, int* n_visited, int* depth,
                               long* left, long* right, long* steps) {
  const int ns = slots_needed(max_depth);
  DummyBackend b(ns);
  for (int i = 0; i < nt; ++i) b.turning.insert(turning[i]);
  for (int i = 0; i < nd; ++i) b.divergent.insert(divergent[i]);
  b.z = b.z0 = z0;
  NutsMachine<DummyBackend> m(b, dm_make_key(1, 1), max_depth, -1000.0, ns);
  dhmc_tree_stats ts;
  m.transition(0, 1.0, nullptr, &flags, &ts);
  *n_visited = (int)b.visited.size();
  for (size_t i = 0; i < b.visited.size() && i < 4096; ++i) visited[i] = b.visited[i];
  *depth = (int)ts.depth; *left = ts.left; *right = ts.right; *steps = ts.steps;
  return 0;
}

int hs_slots_needed(int max_depth) { return slots_needed(max_depth); }

// N transitions of one chain starting at q with metric minv and step size eps.
// adapt4 = {delta, gamma, kappa, t0} or NULL; metric = DHMC_METRIC_*.
// Returns the chain status bits; eps_out = final ϵ; minv is updated in place.
int hs_run(int family, int D, const double* params, int T, double* minv, int max_depth,
           double min_delta, uint64_t seed, uint64_t chain, uint32_t t0, int N, double* q,
           double eps, const double* adapt4, int metric, const double* p_override,
           const uint32_t* dir_override, double* post, dhmc_tree_stats* stats, double* logd,
           double* eps_used, double* eps_out, double* lq_out, double* g_out) {
  const int ns = slots_needed(max_depth);
  HostBackend b(D, T, family, params, ns);
  b.q.assign(q, q + D);
  b.minv.assign(minv, minv + D);
  bool gbad;
  b.lq = b.eval_model(&gbad);
  NutsMachine<HostBackend> m(b, dm_make_key(seed, chain), max_depth, min_delta, ns);
  AdaptConfig cfg{};
  cfg.adapt = adapt4 != nullptr;
  if (adapt4) { cfg.delta = adapt4[0]; cfg.gamma = adapt4[1]; cfg.kappa = adapt4[2]; cfg.t0 = (int)adapt4[3]; }
  cfg.metric = metric;
  Sink sink{b, post, stats, logd, eps_used};
  double e = m.run(t0, N, eps, cfg, p_override, dir_override, sink);
  if (eps_out) *eps_out = e;
  std::memcpy(q, b.q.data(), sizeof(double) * D);
  std::memcpy(minv, b.minv.data(), sizeof(double) * D);
  if (lq_out) *lq_out = b.lq;
  if (g_out) std::memcpy(g_out, b.g.data(), sizeof(double) * D);
  return m.status;
}

int hs_find_initial_stepsize(int family, int D, const double* params, int T, const double* minv,
                             uint64_t seed, uint64_t chain, const double* q, const double* p_override,
                             double initial_eps, double log_threshold, int maxiter, double* eps_out) {
  const int ns = slots_needed(10);
  HostBackend b(D, T, family, params, ns);
  b.q.assign(q, q + D);
  b.minv.assign(minv, minv + D);
  bool gbad;
  b.lq = b.eval_model(&gbad);
  NutsMachine<HostBackend> m(b, dm_make_key(seed, chain), 10, -1000.0, ns);
  *eps_out = m.find_initial_stepsize(initial_eps, log_threshold, maxiter, p_override);
  return m.status;
}

}  // extern "C"
