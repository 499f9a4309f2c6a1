// dhmc_b200.cu — sm_100a kernels and the C ABI (include/dhmc.h) of the
// many-chain NUTS engine.  Build: see csrc/Makefile (nvcc -fmad=false, sm_100a).
//
// Kernels (one chain group of T threads = one CTA; persistent, chains pulled
// from an atomic queue so that ragged tree depths balance across SMs):
//   k_nuts      sample_tree / warmup(::TuningNUTS) / mcmc     NUTS.jl:232-241, mcmc.jl:258-286,366-381
//   k_search    warmup(::InitialStepsizeSearch)               mcmc.jl:134-148, stepsize.jl:46-85
//   k_leapfrog  leapfrog (streaming, HBM-bound)               hamiltonian.jl:273-282
//   k_eval      evaluate_ℓ(strict) / random_position          hamiltonian.jl:202-217, mcmc.jl:108
//   k_phase     logdensity(H, z)                              hamiltonian.jl:251-256
// Logistic family, dim <= 256: k_nuts / k_search run as "packed chain groups" — 8 chains per
// CTA, each with its own warps and state machine, the likelihood evaluated by the whole CTA
// (device_backend.cuh: coop_core).
//
// There is NO CPU fallback: without a CUDA device dhmc_create fails with
// DHMC_ECUDA and nothing else can be called.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only: the library is bound at run time (nccl_api below)

#include <pthread.h>
#include <sched.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/dhmc.h"
#include "kernels.cuh"      // KArgs, shared-memory planning helpers; the kernels themselves are instantiated in family_tu.cu

using namespace dhmc;

// kernel lookup, one function per (family, part) translation unit (family_tu.cu).  The references are WEAK: the stock
// library links every shipped family, a user-model library (`make user`, include/dhmc_models.h) links the USER family's two
// units only, and a family whose units are absent is refused by dhmc_create ("family not built into this library").
// (hidden: never bound across two copies of the library loaded into one process)
#define DHMC_TU_LINKAGE __attribute__((weak, visibility("hidden")))
#define DHMC_DECL_TU(f, p) const void* dhmc_family_kernel_##f##_##p(int W, int epl, int kernel, int dense) DHMC_TU_LINKAGE;
DHMC_DECL_TU(0, 0) DHMC_DECL_TU(1, 0) DHMC_DECL_TU(2, 0) DHMC_DECL_TU(3, 0) DHMC_DECL_TU(3, 1) DHMC_DECL_TU(3, 2)
DHMC_DECL_TU(0, 3) DHMC_DECL_TU(1, 3) DHMC_DECL_TU(2, 3) DHMC_DECL_TU(3, 3)
#undef DHMC_DECL_TU
extern "C" {
const void* dhmc_user_family_kernel_0(int W, int epl, int kernel, int dense) DHMC_TU_LINKAGE;
const void* dhmc_user_family_kernel_3(int W, int epl, int kernel, int dense) DHMC_TU_LINKAGE;
const char* dhmc_user_family_name_str(void) DHMC_TU_LINKAGE;
int dhmc_user_family_min_dim(void) DHMC_TU_LINKAGE;
}
// part: 0 = one chain per CTA, 1 = packed chain groups with the FMA likelihood, 2 = packed groups on the tensor cores,
// 3 = one chain per CTA with max_depth > 12 (persistent kernels only)
typedef const void* (*family_tu_fn)(int, int, int, int);
static family_tu_fn family_tu(int fam, int part) {
  switch (fam * 4 + part) {
    case 0: return dhmc_family_kernel_0_0;
    case 4: return dhmc_family_kernel_1_0;
    case 8: return dhmc_family_kernel_2_0;
    case 12: return dhmc_family_kernel_3_0;
    case 13: return dhmc_family_kernel_3_1;
    case 14: return dhmc_family_kernel_3_2;
    case 3: return dhmc_family_kernel_0_3;
    case 7: return dhmc_family_kernel_1_3;
    case 11: return dhmc_family_kernel_2_3;
    case 15: return dhmc_family_kernel_3_3;
    case 16: return dhmc_user_family_kernel_0;
    case 19: return dhmc_user_family_kernel_3;
  }
  return nullptr;
}
static const void* lookup_kernel(int fam, int part, int W, int epl, KernelId k, bool dense) {
  const family_tu_fn f = family_tu(fam, part);
  return f ? f(W, epl, k, dense) : nullptr;
}

// ------------------------------------------------------------------ Symmetric metric
// Lower Cholesky factor of the row-major symmetric A, stored transposed:
// Lt[k*D + i] = L[i][k].  One CTA; per output element the subtraction order is
// k = 0..j-1, as in the oracle (cholesky_lower).  Returns false if not positive definite.
__device__ bool chol_lower_t(const double* A, double* Lt, int D) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int j = 0; j < D; ++j) {
    double s = A[(size_t)j * D + j];
    for (int k = 0; k < j; ++k) { const double l = Lt[(size_t)k * D + j]; s = s - l * l; }
    if (!(s > 0.0) || !dm_isfinite(s)) return false;      // uniform across the CTA
    const double d = dm_sqrt(s);
    __syncthreads();
    if (tid == 0) Lt[(size_t)j * D + j] = d;
    for (int i = j + 1 + tid; i < D; i += nt) {
      double t = A[(size_t)i * D + j];
      for (int k = 0; k < j; ++k) t = t - Lt[(size_t)k * D + i] * Lt[(size_t)k * D + j];
      Lt[(size_t)j * D + i] = t / d;
    }
    __syncthreads();
  }
  return true;
}
// W = cholesky(inv(M⁻¹)).L — hamiltonian.jl:73, restated as in oracle dense_factor():
// C = chol(M⁻¹), Ci = C⁻¹, M = CiᵀCi, W = chol(M); W is stored transposed in wt.
__global__ void k_dense_factor(const double* minv_dense, double* wt, double* tmp, int* status, int D, int B) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const size_t dd = (size_t)D * D;
  double* Ct = tmp + (size_t)blockIdx.x * 3 * dd;
  double* Ci = Ct + dd;
  double* M = Ci + dd;
  for (int c = blockIdx.x; c < B; c += gridDim.x) {
    const double* A = minv_dense + (size_t)c * dd;
    bool ok = chol_lower_t(A, Ct, D);
    if (ok) {
      for (int j = tid; j < D; j += nt) {                 // Ci = C⁻¹, one column per thread
        Ci[(size_t)j * D + j] = 1.0 / Ct[(size_t)j * D + j];
        for (int i = j + 1; i < D; ++i) {
          double sacc = 0.0;
          for (int k = j; k < i; ++k) sacc = sacc - Ct[(size_t)k * D + i] * Ci[(size_t)k * D + j];
          Ci[(size_t)i * D + j] = sacc / Ct[(size_t)i * D + i];
        }
      }
      __syncthreads();
      for (int i = tid; i < D; i += nt)                   // M = Ciᵀ Ci
        for (int j = 0; j <= i; ++j) {
          double sacc = 0.0;
          for (int k = i; k < D; ++k) sacc = sacc + Ci[(size_t)k * D + i] * Ci[(size_t)k * D + j];
          M[(size_t)i * D + j] = sacc; M[(size_t)j * D + i] = sacc;
        }
      __syncthreads();
      ok = chol_lower_t(M, wt + (size_t)c * dd, D);
    }
    if (!ok && tid == 0) atomicOr(status + c, (int)DHMC_CHAIN_NOT_POSDEF);
    __syncthreads();
  }
}
// M⁻¹ = regularize_M⁻¹(Symmetric(cov(X; dims = 2)), λ) — mcmc.jl:211, :218-221, from the
// streamed co-moments (transposed lower) of a window of n draws.
__global__ void k_cov_finish(const double* covt, double* minv_dense, int n, double lambda, int D, int B) {
  const size_t dd = (size_t)D * D;
  const double dn1 = (double)(n - 1);
  for (int c = blockIdx.x; c < B; c += gridDim.x) {
    const double* ct = covt + (size_t)c * dd;
    double* out = minv_dense + (size_t)c * dd;
    for (int i = threadIdx.x; i < D; i += blockDim.x)
      for (int j = 0; j <= i; ++j) {
        const double sij = ct[(size_t)j * D + i] / dn1;
        double v = (1 - lambda) * sij;
        if (i == j) v = v + lambda * sij;
        out[(size_t)i * D + j] = v; out[(size_t)j * D + i] = v;
      }
  }
}
// Pooled metric of every group of 8 chains (DHMC_METRIC_SYMMETRIC_POOLED): the chains' streaming window means and co-moments
// (n draws each) merged in the oracle's fixed order (pooled_regularized_cov), shrunk as regularize_M⁻¹, written to all 8 chains.
__global__ void k_cov_pool(const double* covt, const double* means, double* minv_dense, int n, double lambda, int D, int B) {
  extern __shared__ double gm[];              // group mean [D]
  const size_t dd = (size_t)D * D;
  const double dn = (double)n;
  constexpr int G = 8;
  for (int g = blockIdx.x; g < B / G; g += gridDim.x) {
    const size_t c0 = (size_t)g * G;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      double m = means[c0 * D + i];
      for (int c = 1; c < G; ++c) m = m + means[(c0 + c) * D + i];
      gm[i] = m / (double)G;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x)
      for (int j = 0; j <= i; ++j) {
        double acc = 0.0;
        for (int c = 0; c < G; ++c) {
          const double di = means[(c0 + c) * D + i] - gm[i], dj = means[(c0 + c) * D + j] - gm[j];
          acc = acc + (covt[(c0 + c) * dd + (size_t)j * D + i] + (dn * di) * dj);
        }
        const double sij = acc / ((double)G * dn - 1.0);
        double v = (1 - lambda) * sij;
        if (i == j) v = v + lambda * sij;
        for (int c = 0; c < G; ++c) {
          double* out = minv_dense + (c0 + c) * dd;
          out[(size_t)i * D + j] = v; out[(size_t)j * D + i] = v;
        }
      }
    __syncthreads();
  }
}
__global__ void k_broadcast_mat(double* dst, const double* src, size_t dd, size_t B) {
  const size_t n = dd * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i % dd];
}

// Mp[c][i][j] = M[c][i][j] for i, j < D, zero elsewhere ([B][rows][xs])
__global__ void k_pad_metric(const double* M, double* Mp, size_t D, size_t rows, size_t xs, size_t B) {
  const size_t per = rows * xs, tot = per * B;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
    const size_t c = t / per, r = t % per, i = r / xs, j = r % xs;
    Mp[t] = (i < D && j < D) ? M[c * D * D + i * D + j] : 0.0;
  }
}
// Xp[n][j] = X[n][j] for n < N, j < D, zero elsewhere (rows x xs)
__global__ void k_pad_rows(const double* X, double* Xp, size_t N, size_t D, size_t rows, size_t xs) {
  const size_t tot = rows * xs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / xs, j = i % xs;
    Xp[i] = (n < N && j < D) ? X[n * D + j] : 0.0;
  }
}
// Xt[j][n] = X[n][j], rows of Xt padded to ld >= N (pad = 0)
__global__ void k_transpose(const double* X, double* Xt, size_t N, size_t D, size_t ld) {
  const size_t tot = ld * D;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = i / ld, n = i % ld;
    Xt[i] = n < N ? X[n * D + j] : 0.0;
  }
}
// Device-side reduction of tree statistics (Diagnostics.summarize_tree_statistics /
// EBFMI, diagnostics.jl:29-32, 65-106): one warp per chain over its N records.
__global__ void k_tree_summary(const dhmc_tree_stats* stats, int N, int B, unsigned long long* depth_counts /*[33]*/,
                               unsigned long long* term_counts /*[3]: max_depth, divergence, turning*/,
                               double* acc_sum, unsigned long long* step_sum, double* ebfmi /*[B] or null*/) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int c = warp; c < B; c += nwarps) {
    const dhmc_tree_stats* s = stats + (size_t)c * N;
    double a = 0.0, sp = 0.0, sd2 = 0.0;
    unsigned long long st = 0;
    for (int n = lane; n < N; n += 32) {
      const dhmc_tree_stats r = s[n];
      a += r.acceptance_rate; st += (unsigned long long)r.steps; sp += r.pi;
      if (n + 1 < N) { const double d = s[n + 1].pi - r.pi; sd2 += d * d; }
      atomicAdd(depth_counts + (r.depth < 32 ? r.depth : 32), 1ull);
      const int k = (r.left == 1 && r.right == 0) ? 0 : (r.left == r.right ? 1 : 2);
      atomicAdd(term_counts + k, 1ull);
    }
    for (int o = 16; o; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); sp += __shfl_xor_sync(0xffffffffu, sp, o);
      sd2 += __shfl_xor_sync(0xffffffffu, sd2, o); st += __shfl_xor_sync(0xffffffffu, st, o);
    }
    const double mean = sp / N;
    double ss = 0.0;
    for (int n = lane; n < N; n += 32) { const double d = s[n].pi - mean; ss += d * d; }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) {
      atomicAdd(acc_sum, a); atomicAdd(step_sum, st);
      if (ebfmi) ebfmi[c] = (N > 1) ? (sd2 / (N - 1)) / (ss / (N - 1)) : dm_nan();
    }
  }
}

// ---- cross-chain convergence diagnostics on device-resident draws [B][N][D] (§8f-2; the reference's tests use
// MCMCDiagnosticTools.ess_rhat on the same quantities, sample-correctness_utilities.jl:40-43).  Every chain is split in
// two halves of n = N/2 draws (m = 2B sequences).  One warp = 32 consecutive parameters of one sequence: mean, then the
// biased autocovariances at lags 0…L; the per-parameter sums over sequences are accumulated with atomics:
//   acc[d][0] = Σ (μ − pilot_d), [1] = Σ (μ − pilot_d)², [2 + t] = Σ acov(t)        (pilot_d = mean of sequence 0: a shift
//   that keeps the variance of the means free of cancellation).  The host finishes R̂ and the Geyer sum (tiny).
__global__ void k_pilot_mean(const double* draws, int n, int N, int D, double* pilot) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += draws[(size_t)i * D + d];
  pilot[d] = s / n;
}
__global__ void k_ess_rhat(const double* draws, int N, int n, int D, int B, int L, const double* pilot, double* acc) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  const int tiles = (D + 31) / 32;
  const long work = (long)2 * B * tiles;                    // (sequence, parameter tile)
  for (long w = warp; w < work; w += nwarps) {
    const long seq = w / tiles;
    const int d = (int)(w % tiles) * 32 + lane;
    if (d >= D) continue;
    const double* x = draws + ((size_t)(seq >> 1) * N + (size_t)(seq & 1) * n) * D + d;
    double mu = 0.0;
    for (int i = 0; i < n; ++i) mu += x[(size_t)i * D];
    mu /= n;
    double* a = acc + (size_t)d * (L + 3);
    const double dm = mu - pilot[d];
    atomicAdd(a, dm);
    atomicAdd(a + 1, dm * dm);
    for (int t = 0; t <= L; ++t) {
      double c = 0.0;
      for (int i = 0; i + t < n; ++i) c += (x[(size_t)i * D] - mu) * (x[(size_t)(i + t) * D] - mu);
      atomicAdd(a + 2 + t, c / n);
    }
  }
}
// histogram of the acceptance rates (4096 bins on [0, 1]) of a device statistics buffer
__global__ void k_acceptance_hist(const dhmc_tree_stats* stats, size_t n, unsigned long long* hist, int bins) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double a = stats[i].acceptance_rate;
    int b = a >= 1.0 ? bins - 1 : a <= 0.0 ? 0 : (int)(a * bins);
    if (!(a == a)) b = bins;                                  // NaN bucket
    atomicAdd(hist + b, 1ull);
  }
}

// broadcast a D-vector (or scalar when D == 1) to all chains
__global__ void k_broadcast(double* dst, const double* src, size_t D, size_t B) {
  const size_t n = D * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i % D];
}
__global__ void k_fill(double* dst, double v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = v;
}

// ================================================================== host side
struct dhmc_handle {
  dhmc_config cfg;
  int T = 0, W = 0, EPL = 0;
  int G = 1;                        // chains per CTA of the persistent kernels (packed chain groups)
  bool deep = false;                // max_depth > 12: kernels whose slot pool spills past 64 slots (part 3), one chain per CTA
  bool coop_mma = false;            // packed groups: likelihood rounds on the FP64 tensor cores, X streamed by TMA
  size_t stride = 0;
  int n_slots = 0, n_sm = 0, grid = 0, sm_count = 0, light_grid = 0;
  int levels = 13, ntab = 64;       // stack entries per warp (max_depth + 1), slot-table entries (>= n_slots)
  size_t smem_bytes = 0, smem_light = 0;
  size_t scratch_per_cta = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr, h2d_stream = nullptr;
  cudaEvent_t h2d_ev[16] = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t chunk_ev[16] = {};
  cudaEvent_t copy_ev[16] = {};
  bool trace = false;
  double* tmp_b = nullptr;          // [B] scratch (phase log densities) and [B·D] momentum override / [D·D] broadcast source,
  double* tmp_bd = nullptr;         // allocated once instead of per call
  unsigned* tmp_dir = nullptr;
  size_t tmp_bd_doubles = 0;
  std::vector<void*> registered;    // caller buffers page-locked on the fly (direct host writes of draws that exceed HBM)
  void* stage[4] = {nullptr, nullptr, nullptr, nullptr};   // grow-only device staging for host outputs
  size_t stage_bytes[4] = {0, 0, 0, 0};
  double *q = nullptr, *g = nullptr, *lq = nullptr, *p = nullptr, *minv = nullptr, *eps = nullptr;
  double* mparams = nullptr;
  int* status = nullptr;
  double* scratch = nullptr;
  unsigned* counter = nullptr;
  unsigned long long* total_steps = nullptr;
  uint32_t t = 0;
  int64_t launches = 0;
  double last_ms = 0;
  int64_t last_steps = 0;
  bool has_position = false, has_eps = false;
  bool dense = false;               // κ is a Symmetric (dense) metric
  double *minv_dense = nullptr, *wt = nullptr, *covt = nullptr, *dense_tmp = nullptr;
  double* mean_pool = nullptr;      // pooled Symmetric stages: window mean of every chain [B][D]
  bool pooled = false;              // the dense metric is shared by every group of 8 chains
  double* minv_pad = nullptr;       // packed groups on the tensor cores: zero-padded row blocks of every chain's M⁻¹
  double *lX = nullptr, *lXt = nullptr, *ly = nullptr, *lr = nullptr;   // logistic regression
  double* lXp = nullptr;            // … zero-padded row blocks of X for the tensor-core likelihood
  int lN = 0, lLd = 0;
  int reg_ctas[2] = {0, 0};         // occupancy of k_nuts (diag, dense)
  size_t smem_sm = 0, smem_cta_max = 0;
  ncclComm_t comm = nullptr;        // multi-GPU: one communicator per handle (dhmc_comm_init)
  int comm_nranks = 1, comm_rank = 0;
  double last_comm_ms = 0;
  size_t l2_persist_max = 0, l2_window_max = 0;   // persisting-L2 carve-out and largest access-policy window of the device
  std::string err;
};

static std::string g_create_err;

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
      return e_ == cudaErrorMemoryAllocation ? DHMC_ENOMEM : DHMC_ECUDA;              \
    }                                                                                 \
  } while (0)

static void set_l2_window(dhmc_handle* h, const void* ptr, size_t bytes);
// a per-chain staging vector in shared memory: mat-vec input (Symmetric metric), β (logistic), the whole position (USER)
static bool needs_staging(const dhmc_handle* h) {
  return h->minv_dense || h->cfg.family == DHMC_FAMILY_LOGISTIC || h->cfg.family == DHMC_FAMILY_USER;
}
static int kernel_part(const dhmc_handle* h, KernelId k, int G);

// rows of N doubles in the logistic scratch: one per CTA of the light kernels, 2·G per CTA
// (residuals and ll terms of every packed chain) of the persistent kernels
static size_t lr_rows(const dhmc_handle* h) {
  return std::max<size_t>((size_t)h->grid * (h->G > 1 ? 2 * (size_t)h->G : 1), (size_t)h->light_grid);
}

// Plan the persistent kernels for the current metric kind: CTAs per SM (register
// limited), how many slots fit in shared memory, and the global scratch arena.
static int plan(dhmc_handle* h) {
  const int T = h->T;
  const size_t B = (size_t)h->cfg.n_chains;
  const size_t slot_doubles = h->stride * (h->dense ? 2 : 1);
  const size_t slot_bytes = sizeof(double) * slot_doubles;
  const size_t xs = needs_staging(h) ? h->stride : 0;
  const int G = h->G;
  auto heavy_smem = [&](int n_sm) -> size_t {
    return G > 1 ? (size_t)G * group_smem_bytes(h->W, n_sm, slot_doubles, xs, h->levels, h->ntab) + coop_smem_bytes(G, h->coop_mma, (int)h->cfg.dim)
                 : smem_layout(h->W, n_sm, slot_doubles, xs, h->levels, h->ntab).total;
  };
  struct { size_t total; } L0{heavy_smem(0)};
  h->smem_light = smem_layout(h->W, 0, slot_doubles, xs).total;      // light kernels: standard layout
  int& reg_ctas = h->reg_ctas[h->dense ? 1 : 0];
  if (reg_ctas == 0) {
    const void* fn = lookup_kernel(h->cfg.family, kernel_part(h, K_NUTS, G), h->W, h->EPL, K_NUTS, h->dense);
    if (!fn) { h->err = "kernel not built into this library for this layout (max_depth > 12 needs a user-model library built with its deep part: USER_PARTS=\"0 3\" / deep=True)"; return DHMC_EARG; }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L0.total);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&reg_ctas, fn, T * G, L0.total);
    if (e != cudaSuccess) { h->err = std::string("occupancy query: ") + cudaGetErrorString(e); return DHMC_ECUDA; }
  }
  if (reg_ctas < 1) { h->err = "kernel does not fit on an SM"; return DHMC_ECUDA; }
  const int ctas = h->cfg.ctas_per_sm > 0 ? std::min(h->cfg.ctas_per_sm, reg_ctas) : reg_ctas;
  size_t per_cta = h->smem_sm / ctas - 1024;   // 1 KB system reservation per CTA
  if (per_cta > h->smem_cta_max) per_cta = h->smem_cta_max;
  long n_sm = per_cta > L0.total ? (long)((per_cta - L0.total) / (slot_bytes * (size_t)G)) : 0;
  const int pool = h->n_slots - kWelfordSlots;   // the two highest slots stay in global memory
  if (n_sm > pool) n_sm = pool;
  h->n_sm = (int)n_sm;
  h->smem_bytes = heavy_smem(h->n_sm);
  h->grid = (int)std::min<size_t>((size_t)ctas * h->sm_count, (B + G - 1) / G);
  h->light_grid = (int)std::min<size_t>((size_t)h->sm_count * 16, B);
  h->scratch_per_cta = (size_t)(h->n_slots - h->n_sm) * slot_doubles;
  cudaFree(h->scratch); h->scratch = nullptr;
  CK(cudaMalloc(&h->scratch, sizeof(double) * h->scratch_per_cta * (size_t)h->grid * (size_t)G));
  // (an access-policy window over the slot arena was measured and rejected: 1.10e8 vs 1.17e8 leapfrog-steps/s at C2)
  if (h->lN) {   // residual scratch of the logistic family follows the grid
    cudaFree(h->lr); h->lr = nullptr;
    CK(cudaMalloc(&h->lr, sizeof(double) * (size_t)h->lN * lr_rows(h)));
  }
  return DHMC_OK;
}

// Keep a re-read working set resident in L2 (access-policy window on the compute stream): the design matrix of the
// logistic family (every SM sweeps it once per gradient; evicted by the metric / co-moment streams otherwise), else the
// slot arena of the persistent kernels (its dirty lines were written back and re-fetched between tree levels).
static void set_l2_window(dhmc_handle* h, const void* ptr, size_t bytes) {
  if (!ptr || !bytes || h->l2_persist_max == 0) return;
  cudaStreamAttrValue v;
  std::memset(&v, 0, sizeof v);
  const size_t win = std::min(bytes, h->l2_window_max);
  v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  v.accessPolicyWindow.num_bytes = win;
  v.accessPolicyWindow.hitRatio = (float)std::min(1.0, 0.9 * (double)h->l2_persist_max / (double)win);
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();
}

static int kernel_part(const dhmc_handle* h, KernelId k, int G) {
  const bool heavy = (k == K_NUTS || k == K_SEARCH);
  if (heavy && h->deep) return 3;
  return G > 1 ? (h->coop_mma ? 2 : 1) : 0;
}

static int ensure_tmp(dhmc_handle* h, size_t doubles) {      // grow-only device scratch shared by the small entry points
  if (!h->tmp_b) CK(cudaMalloc(&h->tmp_b, sizeof(double) * (size_t)h->cfg.n_chains));
  if (!h->tmp_dir) CK(cudaMalloc(&h->tmp_dir, sizeof(unsigned) * (size_t)h->cfg.n_chains));
  if (doubles > h->tmp_bd_doubles) {
    cudaFree(h->tmp_bd); h->tmp_bd = nullptr; h->tmp_bd_doubles = 0;
    CK(cudaMalloc(&h->tmp_bd, sizeof(double) * doubles));
    h->tmp_bd_doubles = doubles;
  }
  return DHMC_OK;
}

static KArgs base_args(dhmc_handle* h) {
  KArgs a;
  std::memset(&a, 0, sizeof(a));
  a.D = (int)h->cfg.dim; a.B = (int)h->cfg.n_chains; a.T = h->T; a.W = h->W;
  a.seed = h->cfg.seed; a.chain_offset = h->cfg.chain_offset;
  a.q = h->q; a.g = h->g; a.lq = h->lq; a.p = h->p; a.minv = h->minv; a.eps = h->eps;
  a.mparams = h->mparams; a.status = h->status;
  a.max_depth = h->cfg.max_depth; a.min_delta = h->cfg.min_delta;
  a.t0 = h->t;
  a.scratch = h->scratch; a.scratch_per_cta = h->scratch_per_cta;
  a.n_sm = h->n_sm; a.n_slots = h->n_slots; a.stride = h->stride * (h->dense ? 2 : 1);
  a.levels = h->levels; a.ntab = h->ntab;
  a.counter = h->counter; a.total_steps = h->total_steps;
  a.chain_begin = 0; a.chain_end = (int)h->cfg.n_chains;
  a.minv_dense = h->minv_dense; a.wt = h->wt; a.covt = nullptr; a.minv_pad = h->minv_pad; a.mean_out = nullptr; a.pooled = h->pooled ? 1 : 0;
  a.xs_doubles = needs_staging(h) ? (int)((size_t)h->T * h->EPL) : 0;
  a.lX = h->lX; a.lXt = h->lXt; a.ly = h->ly; a.lr = h->lr; a.lN = h->lN; a.lLd = h->lLd; a.lXp = h->lXp;
  return a;
}

// heavy = persistent kernels that use the slot pool (k_nuts, k_search)
// timing: 0 = none, 1 = record ev0 before / ev1 after and read the time (synchronises),
// 2 = record ev0 only (first of a series), 3 = record ev1 only (last; caller reads),
// 4 = middle of a series.  reset_steps: zero the Σ steps counter first.
static int read_timer(dhmc_handle* h) {
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_ms = ms;
  return DHMC_OK;
}
static int launch(dhmc_handle* h, KernelId k, KArgs a, int timing, bool reset_steps = true) {
  const bool heavy = (k == K_NUTS || k == K_SEARCH);
  if (!heavy) { a.n_sm = 0; }
  const size_t smem = heavy ? h->smem_bytes : smem_layout(h->W, 0, a.stride, (size_t)a.xs_doubles).total;
  int grid = heavy ? h->grid : h->light_grid;
  const int G = heavy ? h->G : 1;
  if (heavy) grid = std::max(1, std::min(grid, (a.chain_end - a.chain_begin + G - 1) / G));
  if (heavy) {
    CK(cudaMemsetAsync(h->counter, 0, sizeof(unsigned), h->stream));
    if (reset_steps) CK(cudaMemsetAsync(h->total_steps, 0, sizeof(unsigned long long), h->stream));
  }
  if (timing == 1 || timing == 2) CK(cudaEventRecord(h->ev0, h->stream));
#ifdef DHMC_PROFILE_ROUNDS
  unsigned long long* prof_buf = nullptr;
  if (k == K_NUTS && G > 1) {
    CK(cudaMalloc(&prof_buf, sizeof(unsigned long long) * (size_t)grid * 32 * 16));
    CK(cudaMemsetAsync(prof_buf, 0, sizeof(unsigned long long) * (size_t)grid * 32 * 16, h->stream));
    a.prof = prof_buf;
  }
#endif
  {
    const void* fn = lookup_kernel(h->cfg.family, kernel_part(h, k, G), h->W, h->EPL, k, h->dense);
    if (!fn) { h->err = "kernel not built into this library for this layout (max_depth > 12 needs a user-model library built with its deep part: USER_PARTS=\"0 3\" / deep=True)"; return DHMC_EARG; }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { h->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return DHMC_ECUDA; }
    void* params[] = {(void*)&a};
    e = cudaLaunchKernel(fn, dim3(grid), dim3(h->T * G), params, smem, h->stream);
    if (e != cudaSuccess) { h->err = std::string("cudaLaunchKernel: ") + cudaGetErrorString(e); return DHMC_ECUDA; }
  }
  h->launches += 1;
#ifdef DHMC_PROFILE_ROUNDS
  if (prof_buf) {   // profiling build: dump the per-warp cycle counters of this launch as one JSON line
    CK(cudaStreamSynchronize(h->stream));
    std::vector<unsigned long long> hp((size_t)grid * 32 * 16);
    CK(cudaMemcpy(hp.data(), prof_buf, hp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(prof_buf);
    const char* path = std::getenv("DHMC_PROF_DUMP");
    if (FILE* f = std::fopen(path ? path : "dhmc_prof.jsonl", "a")) {
      const int nw = h->T * G / 32;
      std::fprintf(f, "{\"kernel\": %d, \"grid\": %d, \"warps\": %d, \"dense\": %d, \"sum\": [", (int)k, grid, nw, (int)h->dense);
      for (int w = 0; w < nw; ++w) {
        std::fprintf(f, "%s[", w ? ", " : "");
        for (int i = 0; i < 10; ++i) {
          unsigned long long sacc = 0;
          for (int c = 0; c < grid; ++c) sacc += hp[((size_t)c * 32 + w) * 16 + i];
          std::fprintf(f, "%s%llu", i ? ", " : "", sacc);
        }
        std::fprintf(f, "]");
      }
      std::fprintf(f, "]}\n");
      std::fclose(f);
    }
  }
#endif
  if (timing == 1 || timing == 3) CK(cudaEventRecord(h->ev1, h->stream));
  if (timing == 1) return read_timer(h);
  return DHMC_OK;
}

static int sync_and_check_status(dhmc_handle* h, int mask, const char* what) {
  CK(cudaStreamSynchronize(h->stream));
  const size_t B = (size_t)h->cfg.n_chains;
  std::vector<int> st(B);
  CK(cudaMemcpy(st.data(), h->status, sizeof(int) * B, cudaMemcpyDeviceToHost));
  long bad = 0, first = -1;
  for (size_t i = 0; i < B; ++i)
    if (st[i] & mask) { if (first < 0) first = (long)i; ++bad; }
  if (bad) {
    char buf[256];
    std::snprintf(buf, sizeof buf, "%s: %ld chain(s) failed (first: local chain %ld, status 0x%x)",
                  what, bad, first, st[first]);
    h->err = buf;
    return DHMC_ENUMERIC;
  }
  return DHMC_OK;
}

static void choose_layout(int64_t D, int req_T, int* T, int* EPL) {
  if (req_T > 0) {
    *T = req_T;
    const int W = req_T / 32;
    int e = (int)((D + req_T - 1) / req_T);
    e = e <= 1 ? 1 : e <= 2 ? 2 : e <= 4 ? 4 : e <= 8 ? 8 : e <= 16 ? 16 : e <= 32 ? 32 : 0;
    if (e && !layout_supported(W, e)) e = (e < 4 && layout_supported(W, 4)) ? 4 : (e < 8 && layout_supported(W, 8)) ? 8 : 0;
    *EPL = e;
    return;
  }
  if (D <= 32) { *T = 32; *EPL = 1; }
  else if (D <= 64) { *T = 32; *EPL = 2; }
  else if (D <= 128) { *T = 32; *EPL = 4; }
  else if (D <= 256) { *T = 64; *EPL = 4; }
  else if (D <= 512) { *T = 128; *EPL = 4; }
  else if (D <= 1024) { *T = 128; *EPL = 8; }
  else if (D <= 2048) { *T = 256; *EPL = 8; }
  else if (D <= 4096) { *T = 256; *EPL = 16; }
  else if (D <= 8192) { *T = 256; *EPL = 32; }    // the vectors no longer fit the register file: correct, not fast
  else { *T = 0; *EPL = 0; }
}

extern "C" {

const char* dhmc_last_error(dhmc_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int dhmc_comm_destroy(dhmc_handle* h);

int dhmc_destroy(dhmc_handle* h) {
  if (!h) return DHMC_OK;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->q); cudaFree(h->g); cudaFree(h->lq); cudaFree(h->p); cudaFree(h->minv); cudaFree(h->eps);
  cudaFree(h->mparams); cudaFree(h->status); cudaFree(h->scratch); cudaFree(h->counter);
  cudaFree(h->total_steps);
  cudaFree(h->minv_dense); cudaFree(h->wt); cudaFree(h->covt); cudaFree(h->dense_tmp); cudaFree(h->minv_pad); cudaFree(h->mean_pool);
  cudaFree(h->lX); cudaFree(h->lXt); cudaFree(h->ly); cudaFree(h->lr); cudaFree(h->lXp);
  cudaFree(h->tmp_b); cudaFree(h->tmp_bd); cudaFree(h->tmp_dir);
  for (void* r : h->registered) cudaHostUnregister(r);
  if (h->comm) dhmc_comm_destroy(h);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (auto& e : h->chunk_ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->copy_ev) if (e) cudaEventDestroy(e);
  for (auto& b : h->stage) cudaFree(b);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto& e : h->h2d_ev) if (e) cudaEventDestroy(e);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return DHMC_OK;
}

int dhmc_create(const dhmc_config* cfg, dhmc_handle** out) {
  if (!cfg || !out) { g_create_err = "null argument"; return DHMC_EARG; }
  *out = nullptr;
  // @argcheck sites: NUTS.jl:190-191
  if (!(cfg->max_depth > 0 && cfg->max_depth <= kMaxLevels)) { g_create_err = "0 < max_depth <= MAX_DIRECTIONS_DEPTH (32)"; return DHMC_EARG; }   // NUTS.jl:190, trees.jl:10
  if (!(cfg->min_delta < 0)) { g_create_err = "min_delta < 0"; return DHMC_EARG; }
  if (cfg->dim < 1 || cfg->n_chains < 1 || cfg->n_chains > (1ll << 30)) { g_create_err = "dim >= 1, 1 <= n_chains <= 2^30"; return DHMC_EARG; }
  if (cfg->family < 0 || cfg->family >= DHMC_FAMILY_COUNT) { g_create_err = "unknown family"; return DHMC_EARG; }
  if (cfg->family == DHMC_FAMILY_FUNNEL && cfg->dim < 2) { g_create_err = "funnel needs dim >= 2"; return DHMC_EARG; }
  if (cfg->family == DHMC_FAMILY_USER && !family_tu(DHMC_FAMILY_USER, 0)) {
    g_create_err = "this library was built without a user model (make user USER_HEADER=…; api.compile_user_model)";
    return DHMC_EARG;
  }
  if (!family_tu(cfg->family, 0)) { g_create_err = "family not built into this library (a user-model library carries the USER family only)"; return DHMC_EARG; }
  if (cfg->family == DHMC_FAMILY_USER && dhmc_user_family_min_dim && cfg->dim < dhmc_user_family_min_dim()) {
    g_create_err = "dim below the user model's DHMC_USER_MIN_DIM";
    return DHMC_EARG;
  }
  int T = 0, EPL = 0;
  const int rt = cfg->threads_per_chain;
  if (rt != 0 && !(rt == 32 || rt == 64 || rt == 128 || rt == 256)) { g_create_err = "threads_per_chain in {0,32,64,128,256}"; return DHMC_EARG; }
  choose_layout(cfg->dim, rt, &T, &EPL);
  // logistic regression, dim <= 256: kPack chains per CTA (one warp per chain up to dim 128, two
  // above) sharing every pass over X (packed chain groups); an explicit threads_per_chain keeps
  // one chain per CTA
  int pack = 1;
  const bool deep = cfg->max_depth > 12;
  if (cfg->family == DHMC_FAMILY_LOGISTIC && rt == 0 && cfg->dim <= 32 * kPack && !deep) {
    const char* ev = std::getenv("DHMC_PACK");
    // one warp per chain (8 elements per lane above dim 128): eight warps at 255 registers — the tensor-core rounds
    // need only two warps per sub-partition, and the state machine does not spill; DHMC_PACK_WARPS=2: two warps per chain
    const char* evw = std::getenv("DHMC_PACK_WARPS");
    if (!(evw && std::atoi(evw) == 2) && cfg->dim > 128) { T = 32; EPL = 8; }
    if (!(ev && std::atoi(ev) == 0) && packed_layout(T / 32, EPL)) pack = kPack;
  }
  // packed groups evaluate the likelihood on the FP64 tensor cores (coop_core_tma); DHMC_COOP_MMA=0 selects
  // the FMA formulation (coop_core) — same results bit for bit
  const char* evm = std::getenv("DHMC_COOP_MMA");
  const bool coop_mma = pack > 1 && !(evm && std::atoi(evm) == 0);
  if (T == 0 || EPL == 0) { g_create_err = "dim too large for this build (dim <= 32 * threads_per_chain, 32 only for 256 threads: dim <= 8192)"; return DHMC_EARG; }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0") +
                   " (libdhmc_b200 has no CPU fallback)";
    return DHMC_ECUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { g_create_err = "bad device ordinal"; return DHMC_EARG; }
  dhmc_handle* h = new dhmc_handle();
  h->cfg = *cfg; h->T = T; h->W = T / 32; h->EPL = EPL; h->stride = (size_t)T * EPL; h->G = pack; h->coop_mma = coop_mma; h->deep = deep;
  h->n_slots = slots_needed(cfg->max_depth);
  h->levels = deep ? cfg->max_depth + 1 : kStdLevels;        // deep persistent kernels size their stack / slot table at run time
  h->ntab = deep ? std::max(kStdTab, (h->n_slots + 7) & ~7) : kStdTab;
  auto fail = [&](int rc) { g_create_err = h->err; dhmc_destroy(h); return rc; };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return fail(e_ == cudaErrorMemoryAllocation ? DHMC_ENOMEM : DHMC_ECUDA); } } while (0)
  CKC(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CKC(cudaGetDeviceProperties(&prop, cfg->device));
  h->sm_count = prop.multiProcessorCount;
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
  const bool trace = std::getenv("DHMC_TRACE") != nullptr;      // DHMC_TRACE=1: timeline of the chunk pipeline on stderr
  for (auto& e : h->h2d_ev) CKC(cudaEventCreateWithFlags(&e, trace ? cudaEventDefault : cudaEventDisableTiming));
  for (auto& e : h->chunk_ev) CKC(cudaEventCreateWithFlags(&e, trace ? cudaEventDefault : cudaEventDisableTiming));
  for (auto& e : h->copy_ev) CKC(cudaEventCreateWithFlags(&e, trace ? cudaEventDefault : cudaEventDisableTiming));
  h->trace = trace;
  CKC(cudaEventCreate(&h->ev0));
  CKC(cudaEventCreate(&h->ev1));
  const size_t B = (size_t)cfg->n_chains, D = (size_t)cfg->dim;
  CKC(cudaMalloc(&h->q, sizeof(double) * B * D));
  CKC(cudaMalloc(&h->g, sizeof(double) * B * D));
  CKC(cudaMalloc(&h->p, sizeof(double) * B * D));
  CKC(cudaMalloc(&h->minv, sizeof(double) * B * D));
  CKC(cudaMalloc(&h->lq, sizeof(double) * B));
  CKC(cudaMalloc(&h->eps, sizeof(double) * B));
  CKC(cudaMalloc(&h->status, sizeof(int) * B));
  CKC(cudaMalloc(&h->counter, sizeof(unsigned)));
  CKC(cudaMalloc(&h->total_steps, sizeof(unsigned long long)));
  CKC(cudaMalloc(&h->mparams, sizeof(double) * 2 * D));
  CKC(cudaMemsetAsync(h->q, 0, sizeof(double) * B * D, h->stream));
  CKC(cudaMemsetAsync(h->g, 0, sizeof(double) * B * D, h->stream));
  CKC(cudaMemsetAsync(h->p, 0, sizeof(double) * B * D, h->stream));
  CKC(cudaMemsetAsync(h->lq, 0, sizeof(double) * B, h->stream));
  CKC(cudaMemsetAsync(h->eps, 0, sizeof(double) * B, h->stream));
  CKC(cudaMemsetAsync(h->status, 0, sizeof(int) * B, h->stream));
  CKC(cudaMemsetAsync(h->mparams, 0, sizeof(double) * 2 * D, h->stream));
  k_fill<<<1024, 256, 0, h->stream>>>(h->minv, 1.0, B * D);   // κ = GaussianKineticEnergy(D), mcmc.jl:130
  h->launches += 1;

  // Opt-in (DHMC_L2_WINDOW=1): persisting-L2 window over the logistic design matrix.  Measured on the B200: no gain for C4
  // (the matrix stays L2-resident anyway), and the carve-out costs the other kernels L2 capacity (C2: 1.17e8 -> 1.09e8
  // leapfrog-steps/s, streaming leapfrog 0.96 -> 0.42 of the HBM peak), hence off by default.
  if (std::getenv("DHMC_L2_WINDOW") && std::atoi(std::getenv("DHMC_L2_WINDOW")) == 1) {
    h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    if (h->l2_persist_max && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_max) != cudaSuccess) {
      cudaGetLastError();
      h->l2_persist_max = 0;
    }
  }
  h->smem_sm = (size_t)prop.sharedMemPerMultiprocessor;          // 228 KB
  h->smem_cta_max = (size_t)prop.sharedMemPerBlockOptin;         // 227 KB
  {
    int rcp = plan(h);
    if (rcp != DHMC_OK) return fail(rcp);
  }
  CKC(cudaStreamSynchronize(h->stream));
#undef CKC
  *out = h;
  return DHMC_OK;
}

int dhmc_get_layout(dhmc_handle* h, int32_t* T, int32_t* epl) {
  if (!h) return DHMC_EARG;
  if (T) *T = h->T;
  if (epl) *epl = h->EPL;
  return DHMC_OK;
}

int dhmc_family_available(int32_t family, int32_t* available) {
  if (!available) return DHMC_EARG;
  *available = (family >= 0 && family < DHMC_FAMILY_COUNT && family_tu(family, 0)) ? 1 : 0;
  return DHMC_OK;
}

int dhmc_user_family_name(char* name, size_t cap) {
  if (!dhmc_user_family_name_str || !name || cap == 0) return DHMC_EARG;
  std::strncpy(name, dhmc_user_family_name_str(), cap - 1);
  name[cap - 1] = 0;
  return DHMC_OK;
}

int dhmc_set_problem(dhmc_handle* h, const double* params, size_t n) {
  if (!h) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const size_t D = (size_t)h->cfg.dim;
  if (h->cfg.family == DHMC_FAMILY_LOGISTIC) {
    // params = [N, X row-major (N×D), y (N)]
    if (!params || n < 1) { h->err = "dhmc_set_problem: logistic regression needs [N, X, y]"; return DHMC_EARG; }
    const size_t N = (size_t)params[0];
    if (N < 1 || n != 1 + N * D + N) { h->err = "dhmc_set_problem: expected 1 + N*D + N values"; return DHMC_EARG; }
    for (size_t i = 0; i < N; ++i) {       // the model is a Bernoulli likelihood: responses (or their means) in [0, 1]
      const double yv = params[1 + N * D + i];
      if (!(yv >= 0.0 && yv <= 1.0)) { h->err = "dhmc_set_problem: logistic regression needs 0 <= y <= 1"; return DHMC_EARG; }
    }
    cudaFree(h->lX); cudaFree(h->lXt); cudaFree(h->ly); cudaFree(h->lr); cudaFree(h->lXp);
    h->lX = h->lXt = h->ly = h->lr = h->lXp = nullptr;
    const size_t ld = (N + 1) & ~(size_t)1;                 // even leading dimension: 16-byte aligned row segments
    CK(cudaMalloc(&h->lX, sizeof(double) * (N * D + 2)));  // slack for the last 16-byte piece of a tile
    CK(cudaMemsetAsync(h->lX, 0, sizeof(double) * (N * D + 2), h->stream));
    CK(cudaMalloc(&h->lXt, sizeof(double) * ld * D));
    CK(cudaMalloc(&h->ly, sizeof(double) * N));
    CK(cudaMalloc(&h->lr, sizeof(double) * N * lr_rows(h)));
    CK(cudaMemcpyAsync(h->lX, params + 1, sizeof(double) * N * D, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->ly, params + 1 + N * D, sizeof(double) * N, cudaMemcpyHostToDevice, h->stream));
    k_transpose<<<1024, 256, 0, h->stream>>>(h->lX, h->lXt, N, D, ld);
    h->launches += 1;
    if (h->coop_mma) {   // row blocks [32][XS] with zero padding (rows >= N, columns >= D): one bulk copy per block
      const size_t xs = (size_t)tma_xs((int)D), rows = (N + kTmaRows - 1) / kTmaRows * kTmaRows;
      CK(cudaMalloc(&h->lXp, sizeof(double) * rows * xs));
      k_pad_rows<<<1024, 256, 0, h->stream>>>(h->lX, h->lXp, N, D, rows, xs);
      h->launches += 1;
      set_l2_window(h, h->lXp, sizeof(double) * rows * xs);
    }
    h->lN = (int)N; h->lLd = (int)ld;
    CK(cudaStreamSynchronize(h->stream));
    return DHMC_OK;
  }
  if (h->cfg.family == DHMC_FAMILY_USER) {      // any number of doubles, interpreted by the user's formulas
    if (n && !params) { h->err = "dhmc_set_problem: null parameter block"; return DHMC_EARG; }
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(h->mparams); h->mparams = nullptr;
    CK(cudaMalloc(&h->mparams, sizeof(double) * std::max<size_t>(n, 1)));
    if (n) CK(cudaMemcpyAsync(h->mparams, params, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DHMC_OK;
  }
  const size_t want = h->cfg.family == DHMC_FAMILY_DIAG_NORMAL ? 2 * D : 0;
  if (n != want || (want && !params)) { h->err = "dhmc_set_problem: wrong parameter count for this family"; return DHMC_EARG; }
  if (want) CK(cudaMemcpyAsync(h->mparams, params, sizeof(double) * want, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DHMC_OK;
}

static int eval_position(dhmc_handle* h, bool randomize) {
  CK(cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)h->cfg.n_chains, h->stream));
  KArgs a = base_args(h);
  a.strict = 1; a.randomize = randomize ? 1 : 0;
  int rc = launch(h, K_EVAL, a, 0);
  if (rc != DHMC_OK) return rc;
  h->has_position = true;
  return sync_and_check_status(h, DHMC_CHAIN_BAD_INITIAL, "initialize_warmup_state: invalid log density or gradient at the initial position");
}

int dhmc_set_position(dhmc_handle* h, const double* q) {
  if (!h || !q) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->q, q, sizeof(double) * (size_t)h->cfg.n_chains * h->cfg.dim, cudaMemcpyHostToDevice, h->stream));
  return eval_position(h, false);
}
int dhmc_random_position(dhmc_handle* h) {
  if (!h) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  return eval_position(h, true);
}

static int ensure_dense(dhmc_handle* h) {
  if (h->minv_dense) return DHMC_OK;

  const size_t B = (size_t)h->cfg.n_chains, dd = (size_t)h->cfg.dim * h->cfg.dim;
  CK(cudaMalloc(&h->minv_dense, sizeof(double) * B * dd));
  CK(cudaMalloc(&h->wt, sizeof(double) * B * dd));
  CK(cudaMalloc(&h->covt, sizeof(double) * B * dd));
  const int fgrid = (int)std::min<size_t>((size_t)h->sm_count * 4, B);
  CK(cudaMalloc(&h->dense_tmp, sizeof(double) * 3 * dd * (size_t)fgrid));
  CK(cudaMemsetAsync(h->wt, 0, sizeof(double) * B * dd, h->stream));
  if (h->G > 1 && h->coop_mma) {   // [B][⌈D/32⌉·32][XS]: one bulk copy per 32-row block (coop_matvec_tma)
    const size_t rows = ((size_t)h->cfg.dim + kTmaRows - 1) / kTmaRows * kTmaRows;
    CK(cudaMalloc(&h->minv_pad, sizeof(double) * B * rows * (size_t)tma_xs((int)h->cfg.dim)));
  }
  return plan(h);   // the shared-memory layout now carries the mat-vec staging vector
}
// κ = GaussianKineticEnergy(Symmetric M⁻¹): W = cholesky(inv(M⁻¹)).L on device, then switch
// the handle to the dense kernels.
static int factor_and_switch(dhmc_handle* h) {
  const size_t B = (size_t)h->cfg.n_chains;
  const int fgrid = (int)std::min<size_t>((size_t)h->sm_count * 4, B);
  CK(cudaMemsetAsync(h->status, 0, sizeof(int) * B, h->stream));
  k_dense_factor<<<fgrid, 128, 0, h->stream>>>(h->minv_dense, h->wt, h->dense_tmp, h->status, (int)h->cfg.dim, (int)B);
  h->launches += 1;
  if (h->minv_pad) {
    const size_t rows = ((size_t)h->cfg.dim + kTmaRows - 1) / kTmaRows * kTmaRows;
    k_pad_metric<<<2048, 256, 0, h->stream>>>(h->minv_dense, h->minv_pad, (size_t)h->cfg.dim, rows, (size_t)tma_xs((int)h->cfg.dim), B);
    h->launches += 1;
  }
  CK(cudaGetLastError());
  const bool was = h->dense;
  h->dense = true;
  if (!was) { int rc = plan(h); if (rc != DHMC_OK) return rc; }
  return sync_and_check_status(h, DHMC_CHAIN_NOT_POSDEF, "GaussianKineticEnergy: M⁻¹ is not positive definite (PosDefException)");
}

int dhmc_set_metric_dense(dhmc_handle* h, const double* minv, int broadcast) {
  if (!h || !minv) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  int rc = ensure_dense(h);
  if (rc != DHMC_OK) return rc;
  h->pooled = false;
  const size_t B = (size_t)h->cfg.n_chains, dd = (size_t)h->cfg.dim * h->cfg.dim;
  if (broadcast) {
    int rct = ensure_tmp(h, dd);
    if (rct != DHMC_OK) return rct;
    CK(cudaMemcpyAsync(h->tmp_bd, minv, sizeof(double) * dd, cudaMemcpyHostToDevice, h->stream));
    k_broadcast_mat<<<1024, 256, 0, h->stream>>>(h->minv_dense, h->tmp_bd, dd, B);
    h->launches += 1;
    CK(cudaStreamSynchronize(h->stream));
  } else {
    CK(cudaMemcpyAsync(h->minv_dense, minv, sizeof(double) * B * dd, cudaMemcpyHostToDevice, h->stream));
  }
  return factor_and_switch(h);
}
int dhmc_get_metric_dense(dhmc_handle* h, double* minv) {
  if (!h || !minv) return DHMC_EARG;
  if (!h->dense) { h->err = "the current metric is diagonal"; return DHMC_EARG; }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpy(minv, h->minv_dense, sizeof(double) * (size_t)h->cfg.n_chains * h->cfg.dim * h->cfg.dim, cudaMemcpyDeviceToHost));
  return DHMC_OK;
}
int dhmc_metric_is_dense(dhmc_handle* h, int32_t* dense) { if (!h || !dense) return DHMC_EARG; *dense = h->dense ? 1 : 0; return DHMC_OK; }

int dhmc_set_metric(dhmc_handle* h, const double* minv, int broadcast) {
  if (!h) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const size_t B = (size_t)h->cfg.n_chains, D = (size_t)h->cfg.dim;
  if (!minv) {
    k_fill<<<1024, 256, 0, h->stream>>>(h->minv, 1.0, B * D);
  } else if (broadcast) {
    int rct = ensure_tmp(h, D);
    if (rct != DHMC_OK) return rct;
    CK(cudaMemcpyAsync(h->tmp_bd, minv, sizeof(double) * D, cudaMemcpyHostToDevice, h->stream));
    k_broadcast<<<1024, 256, 0, h->stream>>>(h->minv, h->tmp_bd, D, B);
    CK(cudaStreamSynchronize(h->stream));
  } else {
    CK(cudaMemcpyAsync(h->minv, minv, sizeof(double) * B * D, cudaMemcpyHostToDevice, h->stream));
  }
  h->launches += 1;
  CK(cudaStreamSynchronize(h->stream));
  h->pooled = false;
  if (h->dense) { h->dense = false; int rc = plan(h); if (rc != DHMC_OK) return rc; }
  return DHMC_OK;
}

int dhmc_set_stepsize(dhmc_handle* h, const double* eps, int broadcast) {
  if (!h || !eps) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const size_t B = (size_t)h->cfg.n_chains;
  if (broadcast) {
    if (!(eps[0] > 0)) { h->err = "ϵ > 0"; return DHMC_EARG; }    // stepsize.jl:135
    k_fill<<<256, 256, 0, h->stream>>>(h->eps, eps[0], B);
    h->launches += 1;
  } else {
    for (size_t i = 0; i < B; ++i) if (!(eps[i] > 0)) { h->err = "ϵ > 0"; return DHMC_EARG; }
    CK(cudaMemcpyAsync(h->eps, eps, sizeof(double) * B, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  h->has_eps = true;
  return DHMC_OK;
}

int dhmc_set_momentum(dhmc_handle* h, const double* p) {
  if (!h || !p) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpyAsync(h->p, p, sizeof(double) * (size_t)h->cfg.n_chains * h->cfg.dim, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DHMC_OK;
}

int dhmc_get_state(dhmc_handle* h, double* q, double* lq, double* grad, double* minv, double* eps, double* p) {
  if (!h) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const size_t B = (size_t)h->cfg.n_chains, D = (size_t)h->cfg.dim;
  if (q) CK(cudaMemcpyAsync(q, h->q, sizeof(double) * B * D, cudaMemcpyDeviceToHost, h->stream));
  if (grad) CK(cudaMemcpyAsync(grad, h->g, sizeof(double) * B * D, cudaMemcpyDeviceToHost, h->stream));
  if (minv) CK(cudaMemcpyAsync(minv, h->minv, sizeof(double) * B * D, cudaMemcpyDeviceToHost, h->stream));
  if (p) CK(cudaMemcpyAsync(p, h->p, sizeof(double) * B * D, cudaMemcpyDeviceToHost, h->stream));
  if (lq) CK(cudaMemcpyAsync(lq, h->lq, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
  if (eps) CK(cudaMemcpyAsync(eps, h->eps, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DHMC_OK;
}

int dhmc_chain_status(dhmc_handle* h, int32_t* status) {
  if (!h || !status) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemcpy(status, h->status, sizeof(int) * (size_t)h->cfg.n_chains, cudaMemcpyDeviceToHost));
  return DHMC_OK;
}
int dhmc_get_transition_count(dhmc_handle* h, uint32_t* t) { if (!h || !t) return DHMC_EARG; *t = h->t; return DHMC_OK; }
int dhmc_set_transition_count(dhmc_handle* h, uint32_t t) { if (!h) return DHMC_EARG; h->t = t; return DHMC_OK; }

int dhmc_leapfrog(dhmc_handle* h, int32_t n_steps, int32_t sign) {
  if (!h || n_steps < 0) return DHMC_EARG;
  if (!h->has_position || !h->has_eps) { h->err = "dhmc_leapfrog: set position and step size first"; return DHMC_EARG; }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)h->cfg.n_chains, h->stream));   // status words describe the current call
  KArgs a = base_args(h);
  a.lf_steps = n_steps; a.lf_sign = sign;
  int rc = launch(h, K_LEAPFROG, a, 1);
  if (rc != DHMC_OK) return rc;
  return sync_and_check_status(h, DHMC_CHAIN_NONFINITE_Q, "leapfrog: position vector has non-finite elements");
}

int dhmc_phase_logdensity(dhmc_handle* h, double* out) {
  if (!h || !out) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const size_t B = (size_t)h->cfg.n_chains;
  int rc = ensure_tmp(h, 0);
  if (rc != DHMC_OK) return rc;
  KArgs a = base_args(h);
  a.out_phase = h->tmp_b;
  rc = launch(h, K_PHASE, a, 0);
  if (rc != DHMC_OK) return rc;
  CK(cudaMemcpyAsync(out, h->tmp_b, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DHMC_OK;
}

int dhmc_find_initial_stepsize(dhmc_handle* h, double initial_eps, double log_threshold, int32_t maxiter) {
  if (!h) return DHMC_EARG;
  // InitialStepsizeSearch @argchecks — stepsize.jl:31-33
  if (!(std::isfinite(log_threshold) && log_threshold < 0)) { h->err = "isfinite(log_threshold) && log_threshold < 0"; return DHMC_EARG; }
  if (!(std::isfinite(initial_eps) && 0 < initial_eps)) { h->err = "isfinite(initial_ϵ) && 0 < initial_ϵ"; return DHMC_EARG; }
  if (!(maxiter >= 50)) { h->err = "maxiter_crossing ≥ 50"; return DHMC_EARG; }
  if (!h->has_position) { h->err = "set the position first"; return DHMC_EARG; }
  if (h->has_eps) { h->err = "stepsize ϵ manually specified, won't perform initial search"; return DHMC_EARG; }  // mcmc.jl:137
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaMemsetAsync(h->status, 0, sizeof(int) * (size_t)h->cfg.n_chains, h->stream));   // status words describe the current call
  KArgs a = base_args(h);
  a.s_init = initial_eps; a.s_thresh = log_threshold; a.s_maxiter = maxiter;
  int rc = launch(h, K_SEARCH, a, 1);
  if (rc != DHMC_OK) return rc;
  // the reference aborts when the search fails (stepsize.jl:58,78): ϵ counts as set only if every chain found one
  rc = sync_and_check_status(h, DHMC_CHAIN_SEARCH_FAILED | DHMC_CHAIN_NONFINITE_Q,
                             "initial stepsize search failed (no crossing, or non-finite starting density)");
  h->has_eps = (rc == DHMC_OK);
  return rc;
}

// common driver of sample_tree / warmup stage / mcmc.
// Host outputs of large runs are pipelined: the chains are cut into chunks, each
// chunk is one k_nuts launch on the compute stream, and its draws/statistics are
// copied D2H on the copy stream while the next chunk computes (pinned host
// buffers make the copies truly asynchronous; pageable ones still work).
static int ensure_stage(dhmc_handle* h, int i, size_t bytes) {
  if (h->stage_bytes[i] >= bytes) return DHMC_OK;
  cudaFree(h->stage[i]); h->stage[i] = nullptr; h->stage_bytes[i] = 0;
  CK(cudaMalloc(&h->stage[i], bytes));
  h->stage_bytes[i] = bytes;
  return DHMC_OK;
}
// Is `p` page-locked host memory that the device can address (cudaHostAlloc / cudaHostRegister)?  Then *dev is its device alias.
static bool host_mapped(const void* p, void** dev) {
  const char* evn = std::getenv("DHMC_NO_DIRECT");        // A/B switch: always stage
  if (evn && std::atoi(evn) == 1) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}
static int run_nuts(dhmc_handle* h, int N, const AdaptConfig& cfg, double lambda, const double* p_over_host,
                    const uint32_t* dir_over_host, double* posterior, dhmc_tree_stats* stats,
                    double* eps_used, double* logdens, bool outputs_on_device, bool advance_t,
                    const double* q_host = nullptr, int thin = 1, bool pool_metric = false) {
  if ((!h->has_position && !q_host) || !h->has_eps) { h->err = "set position and step size (or run the initial search) first"; return DHMC_EARG; }
  const auto tr0 = std::chrono::steady_clock::now();
  auto tr_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count(); };
  double tr_pt[6] = {0, 0, 0, 0, 0, 0};
  CK(cudaSetDevice(h->cfg.device));
  if (thin < 1 || N % thin != 0) { h->err = "thin >= 1 and N a multiple of thin"; return DHMC_EARG; }
  const size_t B = (size_t)h->cfg.n_chains, D = (size_t)h->cfg.dim, n = (size_t)(N / thin);   // n: kept draws per chain
  double *d_post = nullptr, *d_eps = nullptr, *d_ld = nullptr, *d_p = nullptr;
  dhmc_tree_stats* d_stats = nullptr;
  unsigned* d_dir = nullptr;
  int rc = DHMC_OK;
  auto cleanup = [&] {};              // (the override buffers are the handle's grow-only scratch)
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return e_ == cudaErrorMemoryAllocation ? DHMC_ENOMEM : DHMC_ECUDA; } } while (0)
#define CKS(call) do { int r_ = (call); if (r_ != DHMC_OK) { cleanup(); return r_; } } while (0)
  // Host outputs.  Page-locked (pinned) buffers are written by the kernel itself through their device alias — no staging
  // copy in HBM, no second hop, the PCIe writes overlap the sampling at cache-line granularity and N is not bounded by
  // device memory.  Pageable buffers are staged in HBM and copied chunk by chunk; if the draws would not fit, the buffer
  // is page-locked on the fly (cudaHostRegister) and written directly.
  bool direct[4] = {false, false, false, false};
  if (outputs_on_device) {
    d_post = posterior; d_stats = stats; d_eps = eps_used; d_ld = logdens;
  } else {
    void* dv = nullptr;
    if (posterior) {
      const size_t bytes = sizeof(double) * B * n * D;
      // Draws: staging + DMA copies overlapped chunk by chunk is the faster route when the draws fit in HBM (C2: 21.9 ms per
      // step against 24.1 ms with direct writes — kernel stores reach ~47 GB/s over PCIe, the copy engines 57 GB/s); the
      // kernel writes the host buffer directly when they do not fit (DHMC_DIRECT=1 forces it).
      size_t fr = 0, tot = 0;
      cudaMemGetInfo(&fr, &tot);
      const bool fits = bytes <= h->stage_bytes[0] || bytes + ((size_t)1 << 30) <= fr;
      const char* evd = std::getenv("DHMC_DIRECT");
      const bool force_direct = evd && std::atoi(evd) == 1;
      bool ok = false;
      if (!fits || force_direct) {
        ok = host_mapped(posterior, &dv);
        if (!ok && !fits) {                                                         // pageable and too large: page-lock the caller's buffer
          if (cudaHostRegister(posterior, bytes, cudaHostRegisterMapped) == cudaSuccess) {
            h->registered.push_back(posterior);
            ok = host_mapped(posterior, &dv);
          } else {
            cudaGetLastError();
          }
        }
      }
      if (ok) { d_post = (double*)dv; direct[0] = true; }
      else { CKS(ensure_stage(h, 0, bytes)); d_post = (double*)h->stage[0]; }
    }
    if (stats) {
      if (host_mapped(stats, &dv)) { d_stats = (dhmc_tree_stats*)dv; direct[1] = true; }
      else { CKS(ensure_stage(h, 1, sizeof(dhmc_tree_stats) * B * n)); d_stats = (dhmc_tree_stats*)h->stage[1]; }
    }
    if (eps_used) {
      if (host_mapped(eps_used, &dv)) { d_eps = (double*)dv; direct[2] = true; }
      else { CKS(ensure_stage(h, 2, sizeof(double) * B * n)); d_eps = (double*)h->stage[2]; }
    }
    if (logdens) {
      if (host_mapped(logdens, &dv)) { d_ld = (double*)dv; direct[3] = true; }
      else { CKS(ensure_stage(h, 3, sizeof(double) * B * n)); d_ld = (double*)h->stage[3]; }
    }
  }
  if (p_over_host || dir_over_host) CKS(ensure_tmp(h, p_over_host ? B * D : 0));
  if (p_over_host) {
    d_p = h->tmp_bd;
    CKR(cudaMemcpyAsync(d_p, p_over_host, sizeof(double) * B * D, cudaMemcpyHostToDevice, h->stream));
  }
  if (dir_over_host) {
    d_dir = h->tmp_dir;
    CKR(cudaMemcpyAsync(d_dir, dir_over_host, sizeof(unsigned) * B, cudaMemcpyHostToDevice, h->stream));
  }
  tr_pt[0] = tr_ms();
  KArgs a = base_args(h);
  a.N = N; a.thin = thin; a.N_keep = (int)n; a.cfg = cfg; a.p_override = d_p; a.dir_override = d_dir;
  a.out_q = d_post; a.out_stats = d_stats; a.out_eps = d_eps; a.out_lq = d_ld;
  if (cfg.metric == DHMC_METRIC_SYMMETRIC) a.covt = h->covt;
  if (pool_metric) a.mean_out = h->mean_pool;
  const size_t out_bytes = posterior ? sizeof(double) * B * n * D : 0;
  // chunks must stay many waves long, or the ragged tail of every chunk idles the SMs
  // chunks overlap the staged downloads (and the upload of q_host) with the sampling of the next chunk; with direct
  // host writes only an upload is left to overlap
  const bool staged_big = posterior && !direct[0] && out_bytes >= ((size_t)32 << 20);
  int nchunks = (!outputs_on_device && B >= 4096 && (staged_big || q_host)) ? 16 : 1;
  while (nchunks > 1 && B / (size_t)nchunks < (size_t)8 * (size_t)h->grid * (size_t)h->G) nchunks /= 2;   // >= 8 waves of chain slots per chunk
  if (const char* ev = std::getenv("DHMC_E2E_CHUNKS")) { const int v = std::atoi(ev); if (v >= 1 && v <= 16 && !outputs_on_device) nchunks = v; }
  CKR(cudaMemsetAsync(h->status, 0, sizeof(int) * B, h->stream));   // status words describe the current call
  for (int ci = 0; ci < nchunks; ++ci) {
    // (a pooled metric keeps its groups of 8 chains inside one chunk)
    const size_t unit = h->pooled ? 8 : 1;
    const size_t c0 = (B / unit) * ci / nchunks * unit, c1 = (B / unit) * (ci + 1) / nchunks * unit, nc = c1 - c0;
    if (nc == 0) continue;
    a.chain_begin = (int)c0; a.chain_end = (int)c1;
    if (q_host) {
      // positions of this chunk: H2D on its own stream, then evaluate_ℓ(strict) on the compute
      // stream — overlaps with the previous chunk's sampling and D2H
      if (ci == 0) {   // uploads start after everything already queued on the compute stream
        CKR(cudaEventRecord(h->h2d_ev[15], h->stream));
        CKR(cudaStreamWaitEvent(h->h2d_stream, h->h2d_ev[15], 0));
      }
      CKR(cudaMemcpyAsync(h->q + c0 * D, q_host + c0 * D, sizeof(double) * nc * D, cudaMemcpyHostToDevice, h->h2d_stream));
      CKR(cudaEventRecord(h->h2d_ev[ci], h->h2d_stream));
      CKR(cudaStreamWaitEvent(h->stream, h->h2d_ev[ci], 0));
      KArgs ea = a;
      ea.strict = 1; ea.randomize = 0;
      rc = launch(h, K_EVAL, ea, 0);
      if (rc != DHMC_OK) { cleanup(); return rc; }
    }
    const int timing = nchunks == 1 ? 1 : (ci == 0 ? 2 : (ci == nchunks - 1 ? 3 : 4));
    rc = launch(h, K_NUTS, a, timing == 1 ? 2 : timing, ci == 0);
    if (rc != DHMC_OK) { cleanup(); return rc; }
    if (nchunks == 1) CKR(cudaEventRecord(h->ev1, h->stream));
    if (!outputs_on_device) {
      CKR(cudaEventRecord(h->chunk_ev[ci], h->stream));
      CKR(cudaStreamWaitEvent(h->copy_stream, h->chunk_ev[ci], 0));
      if (posterior && !direct[0]) CKR(cudaMemcpyAsync(posterior + c0 * n * D, d_post + c0 * n * D, sizeof(double) * nc * n * D, cudaMemcpyDeviceToHost, h->copy_stream));
      if (stats && !direct[1]) CKR(cudaMemcpyAsync(stats + c0 * n, d_stats + c0 * n, sizeof(dhmc_tree_stats) * nc * n, cudaMemcpyDeviceToHost, h->copy_stream));
      if (eps_used && !direct[2]) CKR(cudaMemcpyAsync(eps_used + c0 * n, d_eps + c0 * n, sizeof(double) * nc * n, cudaMemcpyDeviceToHost, h->copy_stream));
      if (logdens && !direct[3]) CKR(cudaMemcpyAsync(logdens + c0 * n, d_ld + c0 * n, sizeof(double) * nc * n, cudaMemcpyDeviceToHost, h->copy_stream));
      if (h->trace) CKR(cudaEventRecord(h->copy_ev[ci], h->copy_stream));
    }
  }
  tr_pt[1] = tr_ms();
  unsigned long long steps = 0;
  CKR(cudaMemcpyAsync(&steps, h->total_steps, sizeof steps, cudaMemcpyDeviceToHost, h->stream));
  CKR(cudaStreamSynchronize(h->stream));
  tr_pt[2] = tr_ms();
  CKR(cudaStreamSynchronize(h->copy_stream));
  tr_pt[3] = tr_ms();
  CKS(read_timer(h));
  if (h->trace && nchunks > 1) {
    float t;
    std::fprintf(stderr, "[dhmc trace] chunks %d:", nchunks);
    for (int ci = 0; ci < nchunks; ++ci) {
      float a = -1, b2 = -1, c = -1;
      if (q_host) cudaEventElapsedTime(&a, h->ev0, h->h2d_ev[ci]);
      cudaEventElapsedTime(&b2, h->ev0, h->chunk_ev[ci]);
      if (!outputs_on_device) cudaEventElapsedTime(&c, h->ev0, h->copy_ev[ci]);
      std::fprintf(stderr, " [h2d %.2f kern %.2f d2h %.2f]", a, b2, c);
    }
    cudaEventElapsedTime(&t, h->ev0, h->ev1);
    std::fprintf(stderr, " total kernels %.2f ms | host: setup %.2f, enqueued %.2f, compute stream done %.2f, copy stream done %.2f\n", t,
                 tr_pt[0], tr_pt[1], tr_pt[2], tr_pt[3]);
    cudaGetLastError();
  }
#undef CKR
#undef CKS
  cleanup();
  h->last_steps = (int64_t)steps;
  if (advance_t) h->t += (uint32_t)N;
  if (q_host) h->has_position = true;
  if (h->trace) std::fprintf(stderr, "[dhmc trace] host: before status check %.2f ms\n", tr_ms());
  rc = sync_and_check_status(h, DHMC_CHAIN_NONFINITE_Q | DHMC_CHAIN_BAD_ACCEPTANCE | DHMC_CHAIN_BAD_STEPSIZE | (q_host ? DHMC_CHAIN_BAD_INITIAL : 0),
                             q_host ? "invalid initial position, or non-finite position / acceptance rate / step size while sampling"
                                    : "sampling: non-finite position, acceptance rate or step size");
  if (h->trace) std::fprintf(stderr, "[dhmc trace] host: after status check %.2f ms\n", tr_ms());
  if (rc != DHMC_OK) return rc;
  if (cfg.metric == DHMC_METRIC_DIAGONAL && h->dense) {   // κ ← Diagonal: back to the diagonal kernels
    h->dense = false; h->pooled = false;
    rc = plan(h);
    if (rc != DHMC_OK) return rc;
  }
  if (cfg.metric == DHMC_METRIC_SYMMETRIC) {
    // κ = GaussianKineticEnergy(regularize_M⁻¹(sample_M⁻¹(Symmetric, X), λ)) — mcmc.jl:282
    const int fgrid = (int)std::min<size_t>((size_t)h->sm_count * 4, (size_t)h->cfg.n_chains);
    if (pool_metric)
      k_cov_pool<<<fgrid, 128, sizeof(double) * h->cfg.dim, h->stream>>>(h->covt, h->mean_pool, h->minv_dense, N, lambda, (int)h->cfg.dim, (int)h->cfg.n_chains);
    else
      k_cov_finish<<<fgrid, 128, 0, h->stream>>>(h->covt, h->minv_dense, N, lambda, (int)h->cfg.dim, (int)h->cfg.n_chains);
    h->launches += 1;
    h->pooled = pool_metric;
    rc = factor_and_switch(h);
  }
  return rc;
}

int dhmc_sample_tree(dhmc_handle* h, const double* p, const uint32_t* directions, dhmc_tree_stats* stats) {
  if (!h) return DHMC_EARG;
  AdaptConfig cfg{};
  return run_nuts(h, 1, cfg, 0.0, p, directions, nullptr, stats, nullptr, nullptr, false, true);
}

int dhmc_warmup_stage(dhmc_handle* h, int32_t N, int32_t metric, const dhmc_dual_averaging* da,
                      double lambda, double* posterior, dhmc_tree_stats* stats, double* eps_used,
                      double* logdens) {
  if (!h) return DHMC_EARG;
  // TuningNUTS @argchecks — mcmc.jl:191-192
  if (!(N >= 20)) { h->err = "N ≥ 20"; return DHMC_EARG; }
  if (!(lambda >= 0)) { h->err = "λ ≥ 0"; return DHMC_EARG; }
  if (metric != DHMC_METRIC_NOTHING && metric != DHMC_METRIC_DIAGONAL && metric != DHMC_METRIC_SYMMETRIC &&
      metric != DHMC_METRIC_SYMMETRIC_POOLED) { h->err = "metric: Nothing, Diagonal, Symmetric (or the pooled Symmetric option)"; return DHMC_EARG; }
  const bool pool = metric == DHMC_METRIC_SYMMETRIC_POOLED;
  if (pool && (h->cfg.n_chains % 8 != 0 || h->cfg.chain_offset % 8 != 0)) { h->err = "pooled metric: n_chains and chain_offset must be multiples of 8"; return DHMC_EARG; }
  AdaptConfig cfg{};
  cfg.metric = pool ? DHMC_METRIC_SYMMETRIC : metric;
  if (da) {
    // DualAveraging @argchecks — stepsize.jl:108-111
    if (!(0 < da->delta && da->delta < 1) || !(da->gamma > 0) || !(0.5 < da->kappa && da->kappa <= 1) || !(da->t0 >= 0)) {
      h->err = "DualAveraging: 0 < δ < 1, γ > 0, 0.5 < κ ≤ 1, t₀ ≥ 0"; return DHMC_EARG;
    }
    cfg.adapt = 1; cfg.delta = da->delta; cfg.gamma = da->gamma; cfg.kappa = da->kappa; cfg.t0 = da->t0;
  }
  if (cfg.metric == DHMC_METRIC_SYMMETRIC) { int rcd = ensure_dense(h); if (rcd != DHMC_OK) return rcd; }
  if (pool && !h->mean_pool) CK(cudaMalloc(&h->mean_pool, sizeof(double) * (size_t)h->cfg.n_chains * (size_t)h->cfg.dim));
  return run_nuts(h, N, cfg, lambda, nullptr, nullptr, posterior, stats, eps_used, logdens, false, true, nullptr, 1, pool);
}

int dhmc_mcmc(dhmc_handle* h, int32_t N, double* posterior, dhmc_tree_stats* stats, double* logdens) {
  if (!h || N < 0) return DHMC_EARG;
  if (N == 0) return DHMC_OK;
  AdaptConfig cfg{};
  return run_nuts(h, N, cfg, 0.0, nullptr, nullptr, posterior, stats, nullptr, logdens, false, true);
}
int dhmc_mcmc_from(dhmc_handle* h, const double* q, int32_t N, double* posterior, dhmc_tree_stats* stats,
                   double* logdens) {
  if (!h || !q || N < 1) return DHMC_EARG;
  AdaptConfig cfg{};
  return run_nuts(h, N, cfg, 0.0, nullptr, nullptr, posterior, stats, nullptr, logdens, false, true, q);
}
int dhmc_mcmc_thinned(dhmc_handle* h, const double* q, int32_t N, int32_t thin, double* posterior,
                      dhmc_tree_stats* stats, double* logdens) {
  if (!h || N < 1 || thin < 1) return DHMC_EARG;
  AdaptConfig cfg{};
  return run_nuts(h, N, cfg, 0.0, nullptr, nullptr, posterior, stats, nullptr, logdens, false, true, q, thin);
}
// Page-locked host memory on the NUMA node of the handle's GPU: the calling thread is moved to that node's CPUs while the
// pages are allocated and pinned (first touch), so that the kernel's direct writes / the DMA engines cross one PCIe root
// complex and no inter-socket link.  *node receives the NUMA node (or -1 when unknown).
int dhmc_host_alloc(dhmc_handle* h, size_t bytes, void** out, int32_t* node) {
  if (!h || !out || bytes == 0) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  int numa = -1;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, h->cfg.device) == cudaSuccess) {
    for (char* c = bus; *c; ++c) *c = (char)std::tolower(*c);
    std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
    if (f) f >> numa;
  } else {
    cudaGetLastError();
  }
  cpu_set_t old_set, node_set;
  bool moved = false;
  if (numa >= 0 && sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(numa) + "/cpulist");
    std::string list;
    if (f && std::getline(f, list)) {
      CPU_ZERO(&node_set);
      size_t pos = 0;
      while (pos < list.size()) {            // "0-31,64-95"
        size_t end = list.find(',', pos);
        if (end == std::string::npos) end = list.size();
        const std::string tok = list.substr(pos, end - pos);
        const size_t dash = tok.find('-');
        const int a = std::atoi(tok.c_str()), b = dash == std::string::npos ? a : std::atoi(tok.c_str() + dash + 1);
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &old_set)) CPU_SET(c, &node_set);
        pos = end + 1;
      }
      if (CPU_COUNT(&node_set) > 0 && sched_setaffinity(0, sizeof node_set, &node_set) == 0) moved = true;
    }
  }
  void* p = nullptr;
  const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
  if (e != cudaSuccess) { h->err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); return DHMC_ENOMEM; }
  *out = p;
  if (node) *node = numa;
  return DHMC_OK;
}
int dhmc_host_free(dhmc_handle* h, void* p) {
  if (!h) return DHMC_EARG;
  if (p && cudaFreeHost(p) != cudaSuccess) { h->err = "cudaFreeHost failed"; cudaGetLastError(); return DHMC_ECUDA; }
  return DHMC_OK;
}
int dhmc_mcmc_dev(dhmc_handle* h, int32_t N, double* posterior, dhmc_tree_stats* stats, double* logdens) {
  if (!h || N < 0) return DHMC_EARG;
  if (N == 0) return DHMC_OK;
  AdaptConfig cfg{};
  return run_nuts(h, N, cfg, 0.0, nullptr, nullptr, posterior, stats, nullptr, logdens, true, true);
}

int dhmc_tree_summary_dev(dhmc_handle* h, const dhmc_tree_stats* stats_dev, int32_t N, int64_t* depth_counts,
                          int64_t* termination_counts, double* acceptance_sum, int64_t* steps_sum, double* ebfmi) {
  if (!h || !stats_dev || N < 1) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const int B = (int)h->cfg.n_chains;
  unsigned long long* d_cnt = nullptr;   // [33 depth][3 term][1 steps]
  double* d_f = nullptr;                 // [1 acc][B ebfmi]
  CK(cudaMalloc(&d_cnt, sizeof(unsigned long long) * 37));
  CK(cudaMalloc(&d_f, sizeof(double) * (1 + (size_t)B)));
  CK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 37, h->stream));
  CK(cudaMemsetAsync(d_f, 0, sizeof(double), h->stream));
  k_tree_summary<<<h->sm_count * 8, 256, 0, h->stream>>>(stats_dev, N, B, d_cnt, d_cnt + 33, d_f, d_cnt + 36, ebfmi ? d_f + 1 : nullptr);
  h->launches += 1;
  unsigned long long cnt[37];
  double acc = 0;
  cudaError_t e = cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&acc, d_f, sizeof acc, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && ebfmi) e = cudaMemcpyAsync(ebfmi, d_f + 1, sizeof(double) * B, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_cnt); cudaFree(d_f);
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return DHMC_ECUDA; }
  if (depth_counts) for (int i = 0; i < 33; ++i) depth_counts[i] = (int64_t)cnt[i];
  if (termination_counts) for (int i = 0; i < 3; ++i) termination_counts[i] = (int64_t)cnt[33 + i];
  if (steps_sum) *steps_sum = (int64_t)cnt[36];
  if (acceptance_sum) *acceptance_sum = acc;
  return DHMC_OK;
}

int dhmc_ess_rhat_dev(dhmc_handle* h, const double* draws_dev, int32_t N, int32_t max_lag, double* rhat, double* ess) {
  if (!h || !draws_dev || N < 4 || (!rhat && !ess)) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  const int D = (int)h->cfg.dim, B = (int)h->cfg.n_chains, n = N / 2;
  int L = max_lag > 0 ? max_lag : 64;
  if (L > n - 2) L = n - 2;
  if (L < 1) L = 1;
  double *d_pilot = nullptr, *d_acc = nullptr;
  CK(cudaMalloc(&d_pilot, sizeof(double) * D));
  CK(cudaMalloc(&d_acc, sizeof(double) * (size_t)D * (L + 3)));
  CK(cudaMemsetAsync(d_acc, 0, sizeof(double) * (size_t)D * (L + 3), h->stream));
  k_pilot_mean<<<(D + 127) / 128, 128, 0, h->stream>>>(draws_dev, n, N, D, d_pilot);
  k_ess_rhat<<<h->sm_count * 8, 256, 0, h->stream>>>(draws_dev, N, n, D, B, L, d_pilot, d_acc);
  h->launches += 2;
  std::vector<double> acc((size_t)D * (L + 3));
  cudaError_t e = cudaMemcpyAsync(acc.data(), d_acc, sizeof(double) * acc.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_pilot); cudaFree(d_acc);
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return DHMC_ECUDA; }
  const double m = 2.0 * B, dn = (double)n;
  for (int d = 0; d < D; ++d) {
    const double* a = &acc[(size_t)d * (L + 3)];
    const double mean_var = a[2] / m * dn / (dn - 1.0);                    // W: mean within-sequence variance (n − 1)
    const double var_means = m > 1 ? (a[1] - a[0] * a[0] / m) / (m - 1.0) : 0.0;
    const double var_plus = mean_var * (dn - 1.0) / dn + var_means;        // (n−1)/n·W + B/n
    if (rhat) rhat[d] = std::sqrt(var_plus / mean_var);
    if (ess) {
      // Geyer's initial monotone sequence on ρ̂_t = 1 − (W − mean acov_t) / var⁺, pairs (ρ̂_2k + ρ̂_2k+1)
      double tau = 0.0, prev = 1e300;
      for (int t = 0; t + 1 <= L; t += 2) {
        const double r0 = 1.0 - (mean_var - a[2 + t] / m) / var_plus, r1 = 1.0 - (mean_var - a[3 + t] / m) / var_plus;
        double pair = r0 + r1;
        if (!(pair > 0.0)) break;
        if (pair > prev) pair = prev;
        prev = pair;
        tau += 2.0 * pair;
      }
      tau -= 1.0;
      if (tau < 1.0 / std::log10(m * dn)) tau = 1.0 / std::log10(m * dn);   // cap of the super-efficient case (Stan: ESS ≤ S·log10 S)
      ess[d] = m * dn / tau;
    }
  }
  return DHMC_OK;
}
int dhmc_acceptance_quantiles_dev(dhmc_handle* h, const dhmc_tree_stats* stats_dev, int32_t N, const double* probs,
                                  int32_t nprobs, double* out) {
  if (!h || !stats_dev || N < 1 || !probs || nprobs < 1 || !out) return DHMC_EARG;
  CK(cudaSetDevice(h->cfg.device));
  constexpr int BINS = 4096;
  unsigned long long* d_hist = nullptr;
  CK(cudaMalloc(&d_hist, sizeof(unsigned long long) * (BINS + 1)));
  CK(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (BINS + 1), h->stream));
  const size_t n = (size_t)N * (size_t)h->cfg.n_chains;
  k_acceptance_hist<<<h->sm_count * 8, 256, 0, h->stream>>>(stats_dev, n, d_hist, BINS);
  h->launches += 1;
  std::vector<unsigned long long> hist(BINS + 1);
  cudaError_t e = cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * hist.size(), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_hist);
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return DHMC_ECUDA; }
  const unsigned long long tot = n - hist[BINS];
  for (int k = 0; k < nprobs; ++k) {
    if (!(probs[k] >= 0.0 && probs[k] <= 1.0)) { h->err = "0 ≤ p ≤ 1"; return DHMC_EARG; }
    if (tot == 0) { out[k] = dm_nan(); continue; }
    const double target = probs[k] * (double)(tot - 1);             // Julia's quantile (type 7): position in the sorted sample
    unsigned long long cum = 0;
    double q = 1.0;
    for (int b = 0; b < BINS; ++b) {
      if ((double)(cum + hist[b]) > target) {                       // the order statistic lies in bin b: interpolate inside it
        const double frac = hist[b] ? (target - (double)cum + 0.5) / (double)hist[b] : 0.5;
        q = ((double)b + std::min(1.0, std::max(0.0, frac))) / BINS;
        break;
      }
      cum += hist[b];
    }
    out[k] = q;
  }
  return DHMC_OK;
}

int dhmc_last_total_steps(dhmc_handle* h, int64_t* steps) { if (!h || !steps) return DHMC_EARG; *steps = h->last_steps; return DHMC_OK; }
int dhmc_last_kernel_ms(dhmc_handle* h, double* ms) { if (!h || !ms) return DHMC_EARG; *ms = h->last_ms; return DHMC_OK; }
int dhmc_kernel_launches(dhmc_handle* h, int64_t* n) { if (!h || !n) return DHMC_EARG; *n = h->launches; return DHMC_OK; }

// ---- multi-GPU: chains are sharded over ranks with no data-path collective; the one exchange is the
// all-gather of (thinned) draws / final positions at the end.  One process per GPU; rank 0 creates the id, the host
// program (Julia: MPI / Distributed; Python: torch.distributed) carries its 128 bytes to the other ranks.
// NCCL is bound with dlopen("libnccl.so.2") at the first dhmc_comm_* call instead of a DT_NEEDED entry: a host process
// that already carries an NCCL (e.g. the copy bundled with PyTorch) must end up with ONE libnccl, and a process that never
// shards pays nothing.  The soname lookup returns the copy that is already loaded, else the system library.
struct nccl_api {
  decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&::ncclCommInitRank) CommInitRank = nullptr;
  decltype(&::ncclCommDestroy) CommDestroy = nullptr;
  decltype(&::ncclAllGather) AllGather = nullptr;
  decltype(&::ncclGetErrorString) GetErrorString = nullptr;
  decltype(&::ncclGetVersion) GetVersion = nullptr;
  bool ok = false;
  std::string why;
};
static nccl_api& nccl() {
  static nccl_api api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { api.why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return api; }
#define DHMC_NCCL_SYM(name) api.name = reinterpret_cast<decltype(api.name)>(dlsym(lib, "nccl" #name)); if (!api.name) { api.why = "libnccl lacks nccl" #name; return api; }
    DHMC_NCCL_SYM(GetUniqueId) DHMC_NCCL_SYM(CommInitRank) DHMC_NCCL_SYM(CommDestroy) DHMC_NCCL_SYM(AllGather)
    DHMC_NCCL_SYM(GetErrorString) DHMC_NCCL_SYM(GetVersion)
#undef DHMC_NCCL_SYM
    api.ok = true;
  }
  return api;
}
int dhmc_comm_unique_id(void* id128) {
  if (!id128) return DHMC_EARG;
  static_assert(sizeof(ncclUniqueId) == DHMC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  if (!nccl().ok) { g_create_err = nccl().why; return DHMC_ENCCL; }
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) { g_create_err = "ncclGetUniqueId failed"; return DHMC_ENCCL; }
  std::memcpy(id128, &id, sizeof id);
  return DHMC_OK;
}
#define CKN(call)                                                                     \
  do {                                                                                \
    if (!nccl().ok) { h->err = nccl().why; return DHMC_ENCCL; }                       \
    ncclResult_t r_ = (call);                                                         \
    if (r_ != ncclSuccess) { h->err = std::string(#call) + ": " + nccl().GetErrorString(r_); return DHMC_ENCCL; } \
  } while (0)
int dhmc_comm_init(dhmc_handle* h, int32_t nranks, int32_t rank, const void* id128) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return DHMC_EARG;
  if (h->comm) { h->err = "communicator already initialised"; return DHMC_EARG; }
  CK(cudaSetDevice(h->cfg.device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  CKN(nccl().CommInitRank(&h->comm, nranks, id, rank));
  h->comm_nranks = nranks; h->comm_rank = rank;
  return DHMC_OK;
}
int dhmc_comm_destroy(dhmc_handle* h) {
  if (!h) return DHMC_EARG;
  if (h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; h->comm_nranks = 1; h->comm_rank = 0; }
  return DHMC_OK;
}
// ncclAllGather of `count` doubles per rank, DEVICE pointers (e.g. the draws buffer of dhmc_mcmc_dev): recv is
// [nranks][count]; rank order = global chain order, so recv is the [D, N, B·nranks] column-major draws array.
int dhmc_allgather_dev(dhmc_handle* h, const double* send_dev, double* recv_dev, size_t count) {
  if (!h || !send_dev || !recv_dev) return DHMC_EARG;
  if (!h->comm) { h->err = "dhmc_comm_init first"; return DHMC_EARG; }
  CK(cudaSetDevice(h->cfg.device));
  CK(cudaEventRecord(h->ev0, h->stream));
  CKN(nccl().AllGather(send_dev, recv_dev, count, ncclDouble, h->comm, h->stream));
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaEventSynchronize(h->ev1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_comm_ms = ms;
  return DHMC_OK;
}
// The current position of every chain of every rank: recv_dev [D, B·nranks] (DEVICE), one all-gather of D·B doubles per rank.
int dhmc_allgather_positions_dev(dhmc_handle* h, double* recv_dev) {
  if (!h || !recv_dev) return DHMC_EARG;
  return dhmc_allgather_dev(h, h->q, recv_dev, (size_t)h->cfg.n_chains * (size_t)h->cfg.dim);
}
int dhmc_last_comm_ms(dhmc_handle* h, double* ms) { if (!h || !ms) return DHMC_EARG; *ms = h->last_comm_ms; return DHMC_OK; }
#undef CKN

}  // extern "C"
