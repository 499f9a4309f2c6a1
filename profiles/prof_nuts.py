"""Small driver for ncu captures of the NUTS transition kernel (k_nuts).
Usage: python profiles/prof_nuts.py [chains] [draws] [dim] [tpc] [ctas]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry

pkg = entry.load_package()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
D = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
tpc = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ctas = int(sys.argv[5]) if len(sys.argv) > 5 else 0
eng = pkg.Engine(pkg.StandardNormal(D), chains=K, seed=2026, threads_per_chain=tpc, ctas_per_sm=ctas)
eng.random_position()
eng.set_stepsize(0.28)
for it in range(3):
    eng.mcmc(N, keep_draws=False)
    print(it, "steps", eng.last_total_steps(), "ms", eng.last_kernel_ms(),
          "steps/s %.3e" % (eng.last_total_steps() / eng.last_kernel_ms() * 1e3))
