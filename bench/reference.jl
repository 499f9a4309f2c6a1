# reference.jl — times the GENUINE DynamicHMC.jl CPU path on the metric of BASELINE.json (leapfrog-steps/s), for the boxes
# where Julia exists.  NOT RUN in the build image (no Julia, no network): its row in BASELINE.md reads "not run here"; the CPU
# arm that bench.py times instead is the C++ oracle port (oracle/).
#
#   julia --project -t auto bench/reference.jl [C1|C2]
# C1: 100-dim standard MvNormal, 4 chains, default warm-up (900) + 1000 draws  (BASELINE.json configs[0])
# C2-shaped: 1000-dim standard MvNormal, one chain per thread, default warm-up + 200 draws
# Chains run as tasks, one per thread — the pattern the reference's own tests use (OhMyThreads.tcollect,
# test/sample-correctness_utilities.jl:17).  Output: one JSON line in bench.py's format (impl = "julia-reference").
using DynamicHMC, LogDensityProblems, Random, Statistics
import JSON

struct StdNormal; D::Int; end
LogDensityProblems.capabilities(::Type{StdNormal}) = LogDensityProblems.LogDensityOrder{1}()
LogDensityProblems.dimension(ℓ::StdNormal) = ℓ.D
LogDensityProblems.logdensity_and_gradient(::StdNormal, q) = (-sum(abs2, q) / 2, -q)

function run(config)
    D, chains, N = config == "C1" ? (100, 4, 1000) : (1000, Threads.nthreads(), 200)
    ℓ = StdNormal(D)
    mcmc_with_warmup(Random.Xoshiro(0), StdNormal(5), 50; reporter = NoProgressReport())      # compile
    t0 = time()
    results = fetch.([Threads.@spawn(mcmc_keep_warmup(Random.Xoshiro(k), ℓ, N; reporter = NoProgressReport())) for k in 1:chains])
    secs = time() - t0
    steps = sum(sum(s.steps for s in r.inference.tree_statistics) +
                sum(sum(s.steps for s in w.results.tree_statistics) for w in r.warmup if hasproperty(w.results, :tree_statistics))
                for r in results)
    draws = chains * N
    println(JSON.json(Dict("metric" => "leapfrog_steps_per_sec", "value" => steps / secs, "unit" => "leapfrog-steps/s",
                           "impl" => "julia-reference", "config" => Dict("workload" => "$config: $D-dim standard MvNormal, $chains chains, " *
                           "default warm-up + $N draws, DynamicHMC.jl $(pkgversion(DynamicHMC))", "threads" => Threads.nthreads()),
                           "draws_per_sec" => draws / secs, "seconds" => secs, "steps" => steps,
                           "posterior_mean_abs_max" => maximum(abs, mean(reduce(hcat, [r.inference.posterior_matrix for r in results]); dims = 2)))))
end
run(length(ARGS) ≥ 1 ? ARGS[1] : "C1")
