/* dhmc.h — C ABI of the B200-native many-chain NUTS engine (libdhmc_b200.so).
 *
 * Drop-in boundary for the sampler path of tpapp/DynamicHMC.jl:
 * host code (Julia via ccall, or the Python mirror in dynamichmc.jl_b200/)
 * keeps the mcmc_with_warmup / LogDensityProblems surface and calls these entry
 * points; everything numeric runs in hand-written sm_100a CUDA.  Plain pointers
 * and sizes only.  All functions return a status (0 = ok) unless noted.
 *
 * Conventions
 *  - B = n_chains on this handle, D = dim.  Every per-chain vector argument is
 *    [D, B] column-major (chain-major: element i of chain c at c*D + i), which
 *    is Julia's Matrix{Float64}(D, B); draws are [D, N, B] column-major so that
 *    results[k].posterior_matrix (mcmc.jl:230,275) is a zero-copy view.
 *  - dhmc_tree_stats is bit-compatible with TreeStatisticsNUTS (NUTS.jl:208-221).
 *  - Host pointers unless the function name ends in _dev.
 *  - One handle = one device + one stream; calls are synchronous and not
 *    thread-safe per handle (the reference call is single-threaded too).
 *  - Numerical per-chain failures never abort other chains: they set bits in the
 *    per-chain status word and the call returns DHMC_ENUMERIC (the Julia shim
 *    re-throws DynamicHMCError, utilities.jl:17-27, with the chain ids).
 */
#ifndef DHMC_H
#define DHMC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  DHMC_OK = 0,
  DHMC_EARG = 1,     /* -> ArgumentError (@argcheck sites: NUTS.jl:190-191,
                        stepsize.jl:31-33,108-111, mcmc.jl:191-192) */
  DHMC_ENUMERIC = 2, /* -> DynamicHMCError (hamiltonian.jl:203,213,215;
                        stepsize.jl:58,78) for at least one chain */
  DHMC_ECUDA = 3,
  DHMC_ENOMEM = 4,
  DHMC_ENCCL = 5
};

/* per-chain status bits (dhmc_chain_status) */
enum {
  DHMC_CHAIN_BAD_INITIAL = 1,   /* evaluate_ℓ(strict) failed, hamiltonian.jl:212-215 */
  DHMC_CHAIN_SEARCH_FAILED = 2, /* find_initial_stepsize, stepsize.jl:58 / :78 */
  DHMC_CHAIN_NONFINITE_Q = 4,   /* evaluate_ℓ: non-finite position, hamiltonian.jl:203 */
  DHMC_CHAIN_BAD_ACCEPTANCE = 8, /* adapt_stepsize @argcheck 0 ≤ a ≤ 1, stepsize.jl:148 */
  DHMC_CHAIN_NOT_POSDEF = 16,    /* cholesky(inv(M⁻¹)) failed, hamiltonian.jl:73 (PosDefException) */
  DHMC_CHAIN_BAD_STEPSIZE = 32   /* initial_adaptation_state @argcheck ϵ > 0, stepsize.jl:135 (chain left untouched) */
};

/* log-density family ids: see include/dhmc_models.h */
enum { DHMC_METRIC_NOTHING = 0, DHMC_METRIC_DIAGONAL = 1, DHMC_METRIC_SYMMETRIC = 2,
       /* NOT reference semantics (the reference adapts every chain on its own window, mcmc.jl:282): the optional exchange of
        * SURVEY.md §8e.  Every group of 8 consecutive GLOBAL chains ends the window with ONE dense metric estimated from the
        * pooled draws of the group (per-chain streaming moments merged in a fixed order, same shrinkage λ).  With a shared
        * metric the packed kernels run M⁻¹·[8 vectors] as a true FP64 tensor-core GEMM and read the metric once per group.
        * Needs n_chains and chain_offset to be multiples of 8. */
       DHMC_METRIC_SYMMETRIC_POOLED = 3 };

typedef struct dhmc_handle dhmc_handle;

/* Flattened NUTS(; max_depth, min_Δ) (NUTS.jl:178-195) + problem shape. */
typedef struct {
  int32_t device;            /* CUDA device ordinal */
  int32_t family;            /* DHMC_FAMILY_* (LogDensityProblems model id) */
  int64_t dim;               /* LogDensityProblems.dimension(ℓ) */
  int64_t n_chains;          /* chains resident on this handle (local shard) */
  int64_t chain_offset;      /* global id of local chain 0 (RNG key), §8e */
  uint64_t seed;             /* RNG seed (stands in for the rng argument) */
  int32_t max_depth;         /* NUTS.max_depth, 0 < . <= 32 (MAX_DIRECTIONS_DEPTH, trees.jl:10) */
  int32_t threads_per_chain; /* 0 = auto; 32..256, power of two.  LOGISTIC, dim <= 256: auto also packs 8 chains
                              * per CTA (shared passes over X, same results); an explicit value keeps one chain per CTA */
  double min_delta;          /* NUTS.min_Δ < 0 */
  int32_t ctas_per_sm;       /* 0 = auto */
  int32_t reserved;
} dhmc_config;

/* TreeStatisticsNUTS — NUTS.jl:208-221 (56 bytes) */
typedef struct {
  double pi;              /* π: logdensity(H, ζ) of the selected point */
  int64_t depth;
  int64_t left, right;    /* termination::InvalidTree, trees.jl:180-202 */
  double acceptance_rate;
  int64_t steps;
  uint32_t directions;    /* Directions.flags, trees.jl:19-21 */
  uint32_t pad;
} dhmc_tree_stats;

/* DualAveraging(; δ, γ, κ, t₀) — stepsize.jl:98-118 */
typedef struct { double delta, gamma, kappa; int32_t t0; int32_t pad; } dhmc_dual_averaging;

/* ---- lifecycle ------------------------------------------------------- */
int dhmc_create(const dhmc_config* cfg, dhmc_handle** out);
int dhmc_destroy(dhmc_handle* h);
/* Message of the last failing call on h (h == NULL: last dhmc_create failure).
 * Valid until the next call. */
const char* dhmc_last_error(dhmc_handle* h);
/* threads per chain (canonical reduction width T) and elements per thread */
int dhmc_get_layout(dhmc_handle* h, int32_t* threads_per_chain, int32_t* elems_per_thread);

/* ---- problem: replaces the ℓ argument (LogDensityProblems object) ------ */
/* params: DIAG_NORMAL [mu(D), prec(D)]; LOGISTIC [N, X row-major (N*D), y (N)]; STD_NORMAL / FUNNEL: n == 0;
 * USER: any block of doubles, handed to the user's formulas as `params`. */
int dhmc_set_problem(dhmc_handle* h, const double* params, size_t n);
/* The user's own ℓ (LogDensityProblems.logdensity_and_gradient, call site hamiltonian.jl:204) as device code: a library
 * built from a model header (include/dhmc_models.h "the model header contract"; `make -C dynamichmc.jl_b200/csrc user
 * USER_HEADER=… USER_LIB=…`) carries family DHMC_FAMILY_USER (and only that family).  Copies the model's DHMC_USER_NAME
 * (NUL-terminated, truncated to cap) and returns DHMC_OK, or DHMC_EARG in a library without a user model. */
int dhmc_user_family_name(char* name, size_t cap);
/* Whether this library carries the kernels of `family` (the stock library: the four shipped families; a user-model
 * library: DHMC_FAMILY_USER only).  dhmc_create refuses an absent family with DHMC_EARG. */
int dhmc_family_available(int32_t family, int32_t* available);

/* ---- state: initialization = (q, κ, ϵ), mcmc.jl:111-132 ---------------- */
/* q: [D,B]; evaluates ℓ, ∇ℓ strictly (initialize_warmup_state, mcmc.jl:129-132). */
int dhmc_set_position(dhmc_handle* h, const double* q);
/* random_position, mcmc.jl:108: q ~ U[-2,2]^D per chain from the handle's RNG. */
int dhmc_random_position(dhmc_handle* h);
/* κ = GaussianKineticEnergy(Diagonal(minv)) (hamiltonian.jl:80); minv [D,B],
 * or [D] broadcast to all chains when broadcast != 0; NULL = identity (:87). */
int dhmc_set_metric(dhmc_handle* h, const double* minv, int broadcast);
/* κ = GaussianKineticEnergy(Symmetric M⁻¹) (hamiltonian.jl:73): minv is [D,D,B], or
 * [D,D] for all chains when broadcast != 0.  W = cholesky(inv(M⁻¹)).L is computed on the
 * device; a matrix that is not positive definite sets DHMC_CHAIN_NOT_POSDEF. */
int dhmc_set_metric_dense(dhmc_handle* h, const double* minv, int broadcast);
int dhmc_get_metric_dense(dhmc_handle* h, double* minv /* [D,D,B] */);
int dhmc_metric_is_dense(dhmc_handle* h, int32_t* dense);
/* ϵ per chain [B], or one value for all chains when broadcast != 0. */
int dhmc_set_stepsize(dhmc_handle* h, const double* eps, int broadcast);
/* Momentum of the phase point used by dhmc_leapfrog / dhmc_phase_logdensity. */
int dhmc_set_momentum(dhmc_handle* h, const double* p);
/* Any output pointer may be NULL.  q, grad, minv, p: [D,B]; lq, eps: [B].  minv is the
 * DIAGONAL metric; with a Symmetric metric use dhmc_get_metric_dense. */
int dhmc_get_state(dhmc_handle* h, double* q, double* lq, double* grad, double* minv,
                   double* eps, double* p);
/* Status words of the most recent state-changing call (they are reset when a call starts). */
int dhmc_chain_status(dhmc_handle* h, int32_t* status /* [B] */);
/* Number of transitions already drawn per chain (RNG counter); get/set make the
 * (q, κ, ϵ, counter) tuple a checkpoint (mcmc_keep_warmup/mcmc_steps, §8f). */
int dhmc_get_transition_count(dhmc_handle* h, uint32_t* t);
int dhmc_set_transition_count(dhmc_handle* h, uint32_t t);

/* ---- fine-grained path (parity tests) --------------------------------- */
/* leapfrog(H, z, ±ϵ) n_steps times on every chain — hamiltonian.jl:273-282.
 * sign = +1 / -1 (NUTS.jl:28-31). */
int dhmc_leapfrog(dhmc_handle* h, int32_t n_steps, int32_t sign);
/* logdensity(H, z) per chain — hamiltonian.jl:251-256.  out: [B]. */
int dhmc_phase_logdensity(dhmc_handle* h, double* out);
/* One NUTS transition per chain at the current (q, κ, ϵ), no adaptation —
 * sample_tree, NUTS.jl:232-241.  p [D,B] and directions [B] override the RNG
 * draws when non-NULL (the reference's p= / directions= keywords). */
int dhmc_sample_tree(dhmc_handle* h, const double* p, const uint32_t* directions,
                     dhmc_tree_stats* stats /* [B] */);

/* ---- coarse path: warmup stages and inference -------------------------- */
/* warmup(::InitialStepsizeSearch) — mcmc.jl:134-148, stepsize.jl:46-85. */
int dhmc_find_initial_stepsize(dhmc_handle* h, double initial_eps, double log_threshold,
                               int32_t maxiter_crossing);
/* warmup(::TuningNUTS{M}) — mcmc.jl:258-286.  da == NULL: FixedStepsize
 * (stepsize.jl:181-189).  metric: DHMC_METRIC_NOTHING | _DIAGONAL | _SYMMETRIC.  lambda is
 * the shrinkage of regularize_M⁻¹ (mcmc.jl:218-221); identity for Diagonal (:223).
 * Outputs may be NULL: posterior [D,N,B], stats/eps_used/logdens [N,B]. */
int dhmc_warmup_stage(dhmc_handle* h, int32_t N, int32_t metric, const dhmc_dual_averaging* da,
                      double lambda, double* posterior, dhmc_tree_stats* stats,
                      double* eps_used, double* logdens);
/* mcmc — mcmc.jl:366-381: N transitions at the adapted (κ, ϵ). */
int dhmc_mcmc(dhmc_handle* h, int32_t N, double* posterior, dhmc_tree_stats* stats,
              double* logdens);
/* mcmc from host positions: WarmupState.Q given by the caller (q: [D,B], evaluated strictly
 * like dhmc_set_position) with the current (κ, ϵ) — mcmc_steps / mcmc_next_step,
 * mcmc.jl:335-351.  Upload, evaluation, sampling and download are pipelined by chain chunks. */
int dhmc_mcmc_from(dhmc_handle* h, const double* q, int32_t N, double* posterior,
                   dhmc_tree_stats* stats, double* logdens);
/* mcmc with a thinned output (§8f-3): N transitions, every thin-th one is kept, outputs are [D, N/thin, B] resp.
 * [N/thin, B].  q == NULL continues from the resident positions, else as dhmc_mcmc_from.
 * All host-output calls: while the kept draws fit in HBM they are staged there and copied chunk by chunk, overlapped with the
 * sampling of the next chunk (page-locked buffers — cudaHostAlloc / cudaHostRegister / dhmc_host_alloc — make the copies
 * asynchronous); when they do not fit, the sampling kernel writes the (page-locked, if need be on the fly) host buffer
 * directly, so N is not bounded by device memory. */
int dhmc_mcmc_thinned(dhmc_handle* h, const double* q, int32_t N, int32_t thin, double* posterior,
                      dhmc_tree_stats* stats, double* logdens);
/* Page-locked, device-mapped host memory on the NUMA node of the handle's GPU (for the output buffers above). */
int dhmc_host_alloc(dhmc_handle* h, size_t bytes, void** out, int32_t* numa_node);
int dhmc_host_free(dhmc_handle* h, void* p);
/* Same with DEVICE output pointers (draws stay in HBM for an NCCL all-gather). */
int dhmc_mcmc_dev(dhmc_handle* h, int32_t N, double* posterior, dhmc_tree_stats* stats,
                  double* logdens);

/* ---- diagnostics on device-resident statistics (§8f) ------------------------ */
/* Diagnostics.summarize_tree_statistics / EBFMI (diagnostics.jl:29-32, 65-106) reduced on the
 * GPU over stats_dev [N,B] (DEVICE pointer, e.g. the buffer given to dhmc_mcmc_dev): pooled
 * depth counts [33], termination counts [max_depth, divergence, turning], Σ acceptance rate,
 * Σ steps (host outputs, may be NULL) and the per-chain EBFMI [B] (host, may be NULL). */
int dhmc_tree_summary_dev(dhmc_handle* h, const dhmc_tree_stats* stats_dev, int32_t N,
                          int64_t* depth_counts, int64_t* termination_counts,
                          double* acceptance_sum, int64_t* steps_sum, double* ebfmi);

/* Cross-chain convergence diagnostics reduced on the GPU over device-resident draws [D, N, B] (the posterior buffer of
 * dhmc_mcmc_dev): per parameter the split-R̂ and the effective sample size of the pooled sequences (every chain split in two
 * halves; autocorrelations up to max_lag ≤ N/2 − 2, 0 = 64; Geyer's initial monotone sequence) — what the reference's
 * correctness tests compute with MCMCDiagnosticTools.ess_rhat (test/sample-correctness_utilities.jl:40-43).
 * rhat, ess: host [D], either may be NULL. */
int dhmc_ess_rhat_dev(dhmc_handle* h, const double* draws_dev, int32_t N, int32_t max_lag, double* rhat, double* ess);
/* Quantiles of the acceptance rates (Diagnostics.summarize_tree_statistics: a_quantiles at 0.05 … 0.95,
 * diagnostics.jl:35,100-106) of a device statistics buffer [N, B], from a 4096-bin histogram (resolution 2.4e-4). */
int dhmc_acceptance_quantiles_dev(dhmc_handle* h, const dhmc_tree_stats* stats_dev, int32_t N, const double* probs,
                                  int32_t nprobs, double* out);

/* ---- multi-GPU: chains sharded over ranks, ONE all-gather of draws at the end ------------
 * One process (rank) per GPU; a handle owns the chains [chain_offset, chain_offset + n_chains) and the RNG keys use
 * the global chain id, so results do not depend on the number of ranks.  Nothing is exchanged while sampling.
 * The reference has no counterpart (multi-chain = the user runs mcmc_with_warmup K times,
 * docs/src/worked_example.md:97-103); these entry points replace the user's own gather of the per-chain results
 * (stack_posterior_matrices / pool_posterior_matrices, mcmc.jl:602-617, over all ranks). */
#define DHMC_COMM_ID_BYTES 128
/* rank 0: a fresh ncclUniqueId (128 bytes), to be carried to the other ranks by the host program */
int dhmc_comm_unique_id(void* id128);
/* every rank: ncclCommInitRank on the handle's device */
int dhmc_comm_init(dhmc_handle* h, int32_t nranks, int32_t rank, const void* id128);
int dhmc_comm_destroy(dhmc_handle* h);
/* ncclAllGather of `count` doubles per rank between DEVICE buffers (recv: [nranks][count]); with the draws buffer of
 * dhmc_mcmc_dev as `send`, recv is the [D, N, B·nranks] column-major array of all ranks' draws. */
int dhmc_allgather_dev(dhmc_handle* h, const double* send_dev, double* recv_dev, size_t count);
/* the current position of every chain of every rank, recv_dev: [D, B·nranks] (DEVICE) */
int dhmc_allgather_positions_dev(dhmc_handle* h, double* recv_dev);
/* device time of the last all-gather (CUDA events on the handle's stream) */
int dhmc_last_comm_ms(dhmc_handle* h, double* ms);

/* ---- measurement hooks ------------------------------------------------- */
/* Σ tree_statistics.steps over all chains and draws of the last sampling call. */
int dhmc_last_total_steps(dhmc_handle* h, int64_t* steps);
/* Device time (CUDA events on the handle's stream) of the last sampling /
 * leapfrog kernel, and the number of kernels this handle has launched. */
int dhmc_last_kernel_ms(dhmc_handle* h, double* ms);
int dhmc_kernel_launches(dhmc_handle* h, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* DHMC_H */
