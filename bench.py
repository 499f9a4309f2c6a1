#!/usr/bin/env python
"""bench.py — leapfrog-steps/s of the many-chain NUTS hot path (BASELINE.json metric).

Workload at N=1 (BASELINE.json configs[1]): 1000-dim standard MvNormal, 65 536
chains, diagonal M⁻¹, FP64, 1×B200.  Setup (untimed): random start, initial
step-size search, one dual-averaging stage so that ϵ is adapted per chain.
A timed "step" = one pass of the hot path over the batch: `draws_per_step`
NUTS transitions for every chain (dhmc_mcmc_dev, state and outputs in HBM).
`value` = Σ tree_statistics.steps ÷ device time; `e2e` repeats the same step
through the host-buffer C ABI call (pinned H2D of positions, D2H of draws+stats).

N>1 (torchrun): chains are sharded (rank r owns global chains r·B … (r+1)·B-1,
the Philox key is the global id), no data-path collective, weak scaling; the
draws of the last step are all-gathered once with NCCL after the timed region.

--impl reference: the reference's CPU path.  Julia is not in this image, so this
is the oracle port (oracle/, C++ restatement of DynamicHMC.jl) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "leapfrog_steps_per_sec"
UNIT = "leapfrog-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: oracle port, all host threads, same config/metric."""
    if rank != 0:
        return
    po = entry.load_oracle()
    D = args.dim
    cores = os.cpu_count() or 1
    threads = cores
    eps = args.ref_eps
    # calibrate a bounded sample: chains = threads, draws sized for ~8 s per step
    n_chains = threads
    st, sec = po.bench_mcmc(po.FAMILY_STD_NORMAL, D, n_chains, threads, 2, T=128, eps0=eps, seed=2026)
    rate = st / sec
    draws = max(2, int(args.ref_seconds * rate / max(st / 2, 1)))
    times, steps = [], []
    for it in range(args.warmup + args.steps):
        s, sec = po.bench_mcmc(po.FAMILY_STD_NORMAL, D, n_chains, threads, draws, T=128, eps0=eps,
                               seed=2026 + it)
        if it >= args.warmup:
            times.append(sec); steps.append(s)
    total_t, total_s = float(np.sum(times)), int(np.sum(steps))
    value = total_s / total_t
    sample = (f"{n_chains} chains x {draws} draws per step at D={D}, fixed eps={eps}, identity metric, "
              f"one chain per std::thread ({threads} threads)")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": f"{D}-dim standard MvNormal, NUTS, diagonal M^-1, FP64 (CPU arm: bounded sample)",
                       "dim": D, "chains": n_chains, "draws_per_step": draws},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dim", type=int, default=1000)
    ap.add_argument("--chains", type=int, default=65536, help="chains per GPU")
    ap.add_argument("--draws-per-step", type=int, default=2)
    ap.add_argument("--adapt-steps", type=int, default=60)
    ap.add_argument("--threads-per-chain", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--ref-eps", type=float, default=0.25)
    ap.add_argument("--ref-seconds", type=float, default=8.0)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--skip-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at any debug level
        if "DHMC_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["DHMC_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = entry.load_package()
    D, K, n = args.dim, args.chains, args.draws_per_step
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- setup (untimed) ----------------
    chain_offset, K = pkg.parallel.shard(world * K, world, rank)     # weak scaling: K chains per GPU
    eng = pkg.Engine(pkg.StandardNormal(D), chains=K, seed=2026, device=local_rank,
                     chain_offset=chain_offset, threads_per_chain=args.threads_per_chain,
                     ctas_per_sm=args.ctas_per_sm)
    T, EPL = eng.layout()
    eng.random_position()
    eng.find_initial_stepsize()
    eng.warmup_stage(pkg.TuningNUTS(args.adapt_steps, pkg.DualAveraging()))
    eps = eng.get_state(("eps",))["eps"]
    draws = torch.empty((K, n, D), dtype=torch.float64, device=dev)      # [D, n, K] column-major
    stats = torch.empty((K, n, 56), dtype=torch.uint8, device=dev)
    logd = torch.empty((K, n), dtype=torch.float64, device=dev)

    def step_dev():
        eng.mcmc_dev(n, draws.data_ptr(), stats.data_ptr(), logd.data_ptr())
        return eng.last_total_steps(), eng.last_kernel_ms()

    for _ in range(args.warmup):
        step_dev()
    # ---------------- timed: device-resident ----------------
    launches0 = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    tot_steps, dev_ms = 0, 0.0
    for _ in range(args.steps):
        s, ms = step_dev()
        tot_steps += s
        dev_ms += ms
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.kernel_launches() - launches0

    # ---------------- roofline legs ----------------
    hbm_peak, peak_kind = peaks()
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu --set full capture of this same command
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["k_nuts"]
        if (tr["dim"], tr["chains"], tr["draws_per_step"]) == (D, K, n):
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    algo_bytes_per_leapfrog = 48 * D                      # read q,p,∇ℓ; write q′,p′,∇ℓ′
    # standalone streaming leapfrog kernel (HBM-bound): per-chain metric => 56·D B per step
    lf_ms = []
    for _ in range(6):
        eng.leapfrog(1, 1)
        lf_ms.append(eng.last_kernel_ms())
    lf_ms = float(np.median(lf_ms[2:]))
    lf_bytes = 56 * D * K

    # ---------------- e2e: host buffers through the C ABI ----------------
    e2e = None
    if not args.skip_e2e:
        q_host = torch.empty((K, D), dtype=torch.float64).pin_memory()
        post_host = torch.empty((K, n, D), dtype=torch.float64).pin_memory()
        stats_host = torch.empty((K, n, 56), dtype=torch.uint8).pin_memory()
        logd_host = torch.empty((K, n), dtype=torch.float64).pin_memory()
        q_host.copy_(torch.from_numpy(eng.get_state(("q",))["q"]))
        import ctypes as C
        lib, h = eng._lib, eng._h

        def step_e2e():
            eng._ck(lib.dhmc_mcmc_from(h, C.c_void_p(q_host.data_ptr()), C.c_int32(n),
                                       C.c_void_p(post_host.data_ptr()), C.c_void_p(stats_host.data_ptr()),
                                       C.c_void_p(logd_host.data_ptr())))
            return eng.last_total_steps()

        for _ in range(max(1, args.warmup - 1)):
            step_e2e()
        barrier()
        t1 = time.perf_counter()
        e_steps = 0
        for _ in range(args.steps):
            e_steps += step_e2e()
        barrier()
        e_wall = time.perf_counter() - t1
        e2e = (e_steps, e_wall, K * D * 8, K * n * D * 8 + K * n * 56 + K * n * 8)

    # ---------------- multi-GPU: one NCCL all-gather of the draws, after timing ----------------
    gather_ms = None
    if world > 1:
        last = draws[:, n - 1, :].contiguous()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = pkg.parallel.gather_draws(last, world * K)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)

    # ---------------- reduce over ranks ----------------
    loc = torch.tensor([dev_ms, wall, float(tot_steps), float(launches), lf_ms,
                        e2e[1] if e2e else 0.0, float(e2e[0]) if e2e else 0.0,
                        gather_ms or 0.0], dtype=torch.float64, device=dev)
    mx, sm = loc.clone(), loc.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_max = mx[0].item(), mx[1].item()
    steps_all = sm[2].item()
    value = steps_all / (dev_ms_max * 1e-3)

    if rank == 0:
        nuts_ms_per_launch = dev_ms / args.steps
        achieved = (tot_steps / args.steps) * algo_bytes_per_leapfrog / (nuts_ms_per_launch * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{D}-dim standard MvNormal, {K} chains per GPU, NUTS (max_depth 10), "
                                   "diagonal per-chain M^-1, FP64, eps adapted by dual averaging",
                       "dim": D, "chains_per_gpu": K, "draws_per_step": n, "threads_per_chain": T,
                       "elems_per_thread": EPL, "parallelism": f"chains sharded x{world}, no data-path collective",
                       "l2": "state per step (q,grad,minv,draws) = %.1f GB > 126 MB L2" % ((3 + n) * K * D * 8 / 1e9),
                       "mean_eps": float(np.mean(eps)), "leapfrogs_per_transition": tot_steps / (args.steps * n * K)},
            "draws_per_sec": world * K * n * args.steps / (dev_ms_max * 1e-3),
            "wall_ms_per_step": 1e3 * wall_max / args.steps,
            "gpu_launches": int(sm[3].item()),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_nuts (whole NUTS transition, chain state resident on chip)",
                         "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "peak_kind": peak_kind, "traffic": traffic,
                         "algorithmic_bytes_per_launch": (tot_steps / args.steps) * algo_bytes_per_leapfrog,
                         "note": "achieved = leapfrog steps per launch x 48*D B / launch time; the kernel keeps q,p,grad "
                                 "in registers/shared memory across the tree, so actual DRAM traffic is far below the "
                                 "algorithmic bytes and frac may exceed 1 (SURVEY 8d)"},
            "roofline_leapfrog_stream": {"bound": "hbm", "kernel": "k_leapfrog (one leapfrog step per launch, HBM streaming)",
                                         "achieved": lf_bytes / (lf_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                         "frac": lf_bytes / (lf_ms * 1e-3) / 1e9 / hbm_peak, "peak_kind": peak_kind,
                                         "bytes_per_launch": lf_bytes, "ms": lf_ms},
        }
        if e2e:
            line["e2e"] = {"value": sm[6].item() / mx[5].item(), "unit": UNIT,
                           "h2d_bytes_per_step": e2e[2], "d2h_bytes_per_step": e2e[3]}
        if gather_ms is not None:
            line["allgather"] = {"ms": mx[7].item(), "bytes_per_rank": K * D * 8,
                                 "what": "last draw of every chain, one ncclAllGather after sampling"}
        if world == 1:
            po = entry.load_oracle()
            cores = os.cpu_count() or 1
            eps_med = float(np.median(eps))
            st, sec = po.bench_mcmc(po.FAMILY_STD_NORMAL, D, cores, cores, 2, T=T, eps0=eps_med, seed=2026)
            draws_cpu = max(2, int(args.cpu_baseline_seconds * (st / sec) / max(st / 2, 1)))
            st, sec = po.bench_mcmc(po.FAMILY_STD_NORMAL, D, cores, cores, draws_cpu, T=T, eps0=eps_med, seed=2026)
            line["cpu_baseline"] = {"value": st / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{cores} chains x {draws_cpu} draws at D={D}, eps={eps_med:.4f} (median adapted), "
                                              f"identity metric, oracle port, one chain per thread ({cores} threads)"}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
