"""Where does the wall time of one e2e step go?  Variants: plain / torch context first / pause between calls."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
if "torch" in mode:
    import torch
    torch.cuda.set_device(0)
    x = torch.empty((65536, 2, 1000), dtype=torch.float64, device="cuda")
pkg = entry.load_package()
K, D, n = 65536, 1000, 2
eng = pkg.Engine(pkg.StandardNormal(D), chains=K, seed=2026)
eng.random_position(); eng.find_initial_stepsize(); eng.warmup_stage(pkg.TuningNUTS(60, pkg.DualAveraging()))
q_host = eng.host_alloc((K, D)); post = eng.host_alloc((K, n, D))
st = eng.host_alloc((K, n), dtype=pkg._lib.tree_stats_dtype); ld = eng.host_alloc((K, n))
q_host[...] = eng.get_state(("q",))["q"]
out = dict(posterior_matrix=post, tree_statistics=st, logdensities=ld)
for _ in range(3):
    eng.mcmc_from(q_host, n, out=out)
t0 = time.perf_counter()
ts = []
for it in range(10):
    ta = time.perf_counter()
    eng.mcmc_from(q_host, n, out=out)
    ts.append(time.perf_counter() - ta)
    if "pause" in mode:
        time.sleep(0.002)
tot = time.perf_counter() - t0
print(f"{mode}: per call {1e3*np.median(ts):.2f} ms (min {1e3*min(ts):.2f}, max {1e3*max(ts):.2f}); loop {1e3*tot/10:.2f} ms per step")
