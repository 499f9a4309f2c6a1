# B200HMC.jl — the Julia side of the drop-in boundary (NOT executed in this image:
# Julia is not installed here; the Python mirror dynamichmc.jl_b200/api.py binds the
# very same C entry points through ctypes and is what the tests exercise).
#
# Usage next to the real package:
#     using DynamicHMC, B200HMC
#     ℓ = B200HMC.StandardNormal(1000)                  # a DeviceLogDensity
#     results = B200HMC.mcmc_with_warmup(2026, ℓ, 1000; chains = 65_536)
#     results[k].posterior_matrix, results[k].tree_statistics, results[k].κ, results[k].ϵ
# `results[k]` has the fields of DynamicHMC.mcmc_with_warmup's NamedTuple
# (src/mcmc.jl:575-584), so stack_posterior_matrices / pool_posterior_matrices
# (src/mcmc.jl:602-617) and DynamicHMC.Diagnostics work unchanged.
module B200HMC

using DynamicHMC: DynamicHMC, NUTS, DualAveraging, FixedStepsize, InitialStepsizeSearch,
                  TuningNUTS, GaussianKineticEnergy, TreeStatisticsNUTS, DynamicHMCError,
                  default_warmup_stages
using LinearAlgebra: Diagonal
import LogDensityProblems

const LIB = get(ENV, "DHMC_B200_LIB", "libdhmc_b200.so")

# ---- include/dhmc.h --------------------------------------------------------
struct Config                     # dhmc_config
    device::Int32; family::Int32; dim::Int64; n_chains::Int64; chain_offset::Int64
    seed::UInt64; max_depth::Int32; threads_per_chain::Int32; min_delta::Float64
    ctas_per_sm::Int32; reserved::Int32
end
struct DualAveragingC             # dhmc_dual_averaging
    delta::Float64; gamma::Float64; kappa::Float64; t0::Int32; pad::Int32
end
# TreeStatisticsNUTS is isbits with the layout of dhmc_tree_stats (56 bytes), so the
# output buffer is a Matrix{TreeStatisticsNUTS} passed as Ptr{Cvoid}.
@assert sizeof(TreeStatisticsNUTS) == 56

const OK, EARG, ENUMERIC = 0, 1, 2

mutable struct Handle
    ptr::Ptr{Cvoid}
    D::Int; K::Int
    function Handle(cfg::Config)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:dhmc_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, out)
        rc == OK || _throw(rc, C_NULL)
        h = new(out[], cfg.dim, cfg.n_chains)
        finalizer(h -> ccall((:dhmc_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr), h)
    end
end

function _throw(rc, ptr)
    msg = unsafe_string(ccall((:dhmc_last_error, LIB), Cstring, (Ptr{Cvoid},), ptr))
    rc == EARG && throw(ArgumentError(msg))
    if rc == ENUMERIC
        status = ptr == C_NULL ? Int32[] : chain_status(ptr)
        throw(DynamicHMCError(msg, (; failed_chains = findall(!iszero, status), status)))
    end
    error("libdhmc_b200 error [$rc]: $msg")
end
_ck(h::Handle, rc) = rc == OK ? nothing : _throw(rc, h.ptr)
function chain_status(ptr::Ptr{Cvoid}, K = 0)
    st = Vector{Int32}(undef, K)
    K > 0 && ccall((:dhmc_chain_status, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), ptr, st)
    st
end

# ---- device log densities (LogDensityProblems API on the CPU side too) ------
abstract type DeviceLogDensity end
struct StandardNormal <: DeviceLogDensity; D::Int; end
struct DiagNormal <: DeviceLogDensity; μ::Vector{Float64}; σ²::Vector{Float64}; end
struct Funnel <: DeviceLogDensity; D::Int; end
"Logistic regression with a N(0, I) prior: X is N×p, y ∈ {0,1}ᴺ (include/dhmc_models.h, LOGISTIC)."
struct LogisticRegression <: DeviceLogDensity; X::Matrix{Float64}; y::Vector{Float64}; end
family(::StandardNormal) = Int32(0); family(::DiagNormal) = Int32(1); family(::Funnel) = Int32(2)
family(::LogisticRegression) = Int32(3)
params(::DeviceLogDensity) = Float64[]
params(ℓ::DiagNormal) = vcat(ℓ.μ, 1 ./ ℓ.σ²)
params(ℓ::LogisticRegression) = vcat(Float64(size(ℓ.X, 1)), vec(permutedims(ℓ.X)), ℓ.y)   # [N, X row-major, y]
LogDensityProblems.capabilities(::Type{<:DeviceLogDensity}) = LogDensityProblems.LogDensityOrder{1}()
LogDensityProblems.dimension(ℓ::Union{StandardNormal,Funnel}) = ℓ.D
LogDensityProblems.dimension(ℓ::DiagNormal) = length(ℓ.μ)
LogDensityProblems.dimension(ℓ::LogisticRegression) = size(ℓ.X, 2)
function LogDensityProblems.logdensity_and_gradient(ℓ::LogisticRegression, β)
    η = ℓ.X * β
    sum(ℓ.y .* η .- log1p.(exp.(η))) - sum(abs2, β) / 2, ℓ.X' * (ℓ.y .- 1 ./ (1 .+ exp.(-η))) .- β
end
LogDensityProblems.logdensity_and_gradient(::StandardNormal, q) = (-sum(abs2, q) / 2, -q)
function LogDensityProblems.logdensity_and_gradient(ℓ::DiagNormal, q)
    t = (q .- ℓ.μ) ./ ℓ.σ²
    -sum((q .- ℓ.μ) .* t) / 2, -t
end
function LogDensityProblems.logdensity_and_gradient(ℓ::Funnel, q)
    v, x = q[1], @view q[2:end]; S = sum(abs2, x); ev = exp(-v); n = ℓ.D - 1
    -v^2 / 18 - ev * S / 2 - n * v / 2, vcat(-v / 9 + ev * S / 2 - n / 2, -ev .* x)
end

# ---- mcmc_with_warmup(rng, ℓ, N; …) for `chains` chains — src/mcmc.jl:575-584 --
function mcmc_with_warmup(seed::Integer, ℓ::DeviceLogDensity, N::Integer; chains::Integer = 1,
                          initialization = (), warmup_stages = default_warmup_stages(),
                          algorithm = NUTS(), device = 0, chain_offset = 0)
    D = LogDensityProblems.dimension(ℓ)
    h = Handle(Config(device, family(ℓ), D, chains, chain_offset, seed, algorithm.max_depth, 0,
                      algorithm.min_Δ, 0, 0))
    p = params(ℓ)
    _ck(h, ccall((:dhmc_set_problem, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t), h.ptr, p, length(p)))
    # initialize_warmup_state — src/mcmc.jl:129-132
    init = NamedTuple(initialization)
    if haskey(init, :κ)
        m = Matrix{Float64}(repeat(Vector(init.κ.M⁻¹.diag), 1, chains))
        _ck(h, ccall((:dhmc_set_metric, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), h.ptr, m, 0))
    end
    if haskey(init, :q)
        q = Matrix{Float64}(repeat(init.q, 1, chains))            # [D, K] column-major
        _ck(h, ccall((:dhmc_set_position, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, q))
    else
        _ck(h, ccall((:dhmc_random_position, LIB), Cint, (Ptr{Cvoid},), h.ptr))
    end
    haskey(init, :ϵ) &&
        _ck(h, ccall((:dhmc_set_stepsize, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Cint), h.ptr, Float64(init.ϵ), 1))
    for stage in warmup_stages                                   # _warmup fold — src/mcmc.jl:450-457
        warmup!(h, stage)
    end
    posterior = Array{Float64}(undef, D, N, chains)              # [D, N, K]: results[k] is a view
    stats = Matrix{TreeStatisticsNUTS}(undef, N, chains)
    logd = Matrix{Float64}(undef, N, chains)
    _ck(h, ccall((:dhmc_mcmc, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Cvoid}, Ptr{Float64}),
                 h.ptr, N, posterior, stats, logd))
    minv = Matrix{Float64}(undef, D, chains); ϵ = Vector{Float64}(undef, chains)
    _ck(h, ccall((:dhmc_get_state, LIB), Cint,
                 (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                 h.ptr, C_NULL, C_NULL, C_NULL, minv, ϵ, C_NULL))
    [(; posterior_matrix = view(posterior, :, :, k), tree_statistics = view(stats, :, k),
        logdensities = view(logd, :, k), κ = GaussianKineticEnergy(Diagonal(minv[:, k])), ϵ = ϵ[k])
     for k in 1:chains]
end

# GaussianKineticEnergy(Symmetric M⁻¹) — src/hamiltonian.jl:73 (W is computed on the device)
set_metric_dense!(h::Handle, M⁻¹::AbstractMatrix) =
    _ck(h, ccall((:dhmc_set_metric_dense, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint), h.ptr, Matrix{Float64}(M⁻¹), 1))

warmup!(h::Handle, ::Nothing) = nothing                          # src/mcmc.jl:99-101
warmup!(h::Handle, s::InitialStepsizeSearch) =                   # src/mcmc.jl:134-148
    _ck(h, ccall((:dhmc_find_initial_stepsize, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Int32),
                 h.ptr, s.initial_ϵ, s.log_threshold, s.maxiter_crossing))
function warmup!(h::Handle, t::TuningNUTS{M}) where {M}          # src/mcmc.jl:258-286
    metric = M === Nothing ? 0 : M <: Diagonal ? 1 : 2            # DHMC_METRIC_NOTHING/_DIAGONAL/_SYMMETRIC
    a = t.stepsize_adaptation
    da = a isa DualAveraging ? Ref(DualAveragingC(a.δ, a.γ, a.κ, a.t₀, 0)) : C_NULL
    _ck(h, ccall((:dhmc_warmup_stage, LIB), Cint,
                 (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
                 h.ptr, t.N, metric, da, t.λ, C_NULL, C_NULL, C_NULL, C_NULL))
end

end # module
