"""C4-shaped short run for timing/profiling the logistic family: N=10 000, p=256, dense metric.
Usage: python profiles/prof_c4.py [chains] [draws]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ℓ, _ = pkg.LogisticRegression.synthetic(N=10000, p=256, seed=7)
eng = pkg.Engine(ℓ, chains=K, seed=2026)
eng.random_position()
eng.find_initial_stepsize()
rows = []
for st in (pkg.TuningNUTS(20, M=None), pkg.TuningNUTS(20, M=pkg.Symmetric), pkg.TuningNUTS(20, M=None)):
    t0 = time.perf_counter()
    eng.warmup_stage(st)
    rows.append({"N": st.N, "steps": eng.last_total_steps(), "ms": eng.last_kernel_ms(),
                 "rate": eng.last_total_steps() / (eng.last_kernel_ms() * 1e-3), "dense": eng.metric_is_dense()})
eng.mcmc(N)
s, ms = eng.last_total_steps(), eng.last_kernel_ms()
print(json.dumps({"chains": K, "T": eng.layout()[0], "warmup": rows,
                  "sampling": {"draws": N, "steps": s, "ms": ms, "rate": s / (ms * 1e-3)}}))
eng.close()
