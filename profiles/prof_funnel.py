"""ncu driver for config C3 (Neal's funnel, D=10, one warp per chain): python profiles/prof_funnel.py [chains] [draws]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
eng = pkg.Engine(pkg.Funnel(10), chains=K, seed=2026)
eng.random_position()
eng.find_initial_stepsize()
eng.warmup_stage(pkg.TuningNUTS(100, pkg.DualAveraging(), pkg.Diagonal))
for it in range(2):
    out = eng.mcmc(N, keep_draws=False)
    print(it, "steps", eng.last_total_steps(), "ms", eng.last_kernel_ms(),
          "steps/s %.3e" % (eng.last_total_steps() / eng.last_kernel_ms() * 1e3),
          "mean depth %.2f" % out["tree_statistics"]["depth"].mean())
