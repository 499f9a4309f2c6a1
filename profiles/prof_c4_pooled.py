"""C4-shaped timing of the pooled-metric option (DHMC_METRIC_SYMMETRIC_POOLED, NOT reference semantics) next to the per-chain
dense metric.  Usage: python profiles/prof_c4_pooled.py [chains] [draws]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
pkg = entry.load_package()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ℓ, _ = pkg.LogisticRegression.synthetic(N=10000, p=256, seed=7)
out = {}
for name, M in (("per_chain", pkg.Symmetric), ("pooled", pkg.SymmetricPooled)):
    eng = pkg.Engine(ℓ, chains=K, seed=2026)
    eng.random_position(); eng.find_initial_stepsize()
    for st in (pkg.TuningNUTS(20), pkg.TuningNUTS(20, M=M), pkg.TuningNUTS(20)):
        eng.warmup_stage(st)
    eng.mcmc(N)
    s, ms = eng.last_total_steps(), eng.last_kernel_ms()
    out[name] = {"steps": s, "ms": ms, "rate": s / (ms * 1e-3)}
    eng.close()
print(json.dumps({"chains": K, "draws": N, **out}))
