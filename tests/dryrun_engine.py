"""TEST INFRASTRUCTURE — `pytest --dry-engine -m gpu -k <tests>`: run the PYTHON of a -m gpu test on a machine without a GPU.

The `Engine` of the host mirror is replaced by a stand-in that answers the engine calls with the CPU oracle, so that a GPU
test written here (where no device exists) can be checked for what CAN be checked without one: argument shapes, result
keys, index arithmetic, the oracle-side calls and their run time.  Device-vs-oracle assertions hold trivially under it, so
a dry run proves nothing about the CUDA path; the product never imports this file.

The stand-in replays: the oracle has no stepwise state, so every call recomputes the chain from its start
(`po.mcmc_with_warmup` over the stages seen so far) — fine for the short chains of the tests.  Supported: the calls
`mcmc_keep_warmup` / `mcmc_with_warmup` / `mcmc_steps` make (random_position, find_initial_stepsize, warmup_stage, mcmc,
mcmc_from continuing from the current positions, get_state, metric accessors, layout, close)."""
import numpy as np


class DryRunEngine:
    T = 32

    def __init__(self, po, pkg, ℓ, chains, seed=0, algorithm=None, device=0, chain_offset=0, **opts):
        self.po, self.pkg, self.ℓ = po, pkg, ℓ
        self.K, self.D, self.seed, self.off = int(chains), int(ℓ.dimension()), int(seed), int(chain_offset)
        self.algorithm = algorithm or pkg.NUTS()
        self.family = ℓ.family
        pr = np.asarray(ℓ.params(), float)
        self.params = pr if pr.size else None
        self.header = getattr(ℓ, "header", None)
        self.stages, self.search, self.da = [], (0.1, float(np.log(0.8)), 400), (0.8, 0.05, 0.75, 10)
        self.q0 = None
        self.n_drawn = 0
        self._last = None

    # -- plumbing
    def _ctx(self):
        import contextlib
        return self.po.user_model(self.header) if self.header else contextlib.nullcontext()

    def _run(self, N):
        """All chains from their start: the stages seen so far, then N draws."""
        out = []
        with self._ctx():
            for k in range(self.K):
                out.append(self.po.mcmc_with_warmup(
                    self.family, self.D, max(N, 1), self.seed, self.off + k, stages=list(self.stages), params=self.params, T=self.T,
                    max_depth=self.algorithm.max_depth, min_delta=self.algorithm.min_Δ, welford=True, da=self.da, search=self.search,
                    q0=None if self.q0 is None else self.q0[k], keep_warmup=True))
        self._last = out
        return out

    def layout(self):
        return self.T, (self.D + self.T - 1) // self.T

    def random_position(self):
        self.q0 = None

    def set_position(self, q):
        self.q0 = np.array(q, float).reshape(self.K, self.D)

    def set_kinetic_energy(self, κ):
        raise NotImplementedError("dry run: initial κ")

    def set_stepsize(self, e):
        raise NotImplementedError("dry run: initial ϵ")

    def find_initial_stepsize(self, s=None):
        s = s or self.pkg.InitialStepsizeSearch()
        self.search = (s.initial_ϵ, s.log_threshold, s.maxiter_crossing)
        self.stages.append((self.po.STAGE_SEARCH, 0, self.po.METRIC_NOTHING, 0))

    def warmup_stage(self, stage, keep=False):
        a = stage.stepsize_adaptation
        da_on = isinstance(a, self.pkg.DualAveraging)
        if da_on:
            self.da = (a.δ, a.γ, a.κ, a.t0)
        metric = {None: self.po.METRIC_NOTHING, "Diagonal": self.po.METRIC_DIAGONAL, "Symmetric": self.po.METRIC_SYMMETRIC}[stage.M]
        n0 = sum(s[1] for s in self.stages if s[0] == self.po.STAGE_TUNING)
        self.stages.append((self.po.STAGE_TUNING, stage.N, metric, int(da_on)))
        if not keep:
            return None
        out = self._run(0)
        sl = slice(n0, n0 + stage.N)
        return {"posterior_matrix": np.stack([o["warmup_posterior"][sl] for o in out]),
                "tree_statistics": np.stack([o["warmup_stats"][sl] for o in out]),
                "ϵs": np.stack([o["warmup_eps"][sl] for o in out]),
                "logdensities": np.zeros((self.K, stage.N))}

    def mcmc(self, N, keep_draws=True):
        if N == 0:
            return dict(posterior_matrix=np.empty((self.K, 0, self.D)),
                        tree_statistics=np.zeros((self.K, 0), dtype=self.po.tree_stats_dtype), logdensities=np.empty((self.K, 0)))
        out = self._run(self.n_drawn + N)
        sl = slice(self.n_drawn, self.n_drawn + N)
        self.n_drawn += N
        return dict(posterior_matrix=np.stack([o["posterior_matrix"][sl] for o in out]),
                    tree_statistics=np.stack([o["tree_statistics"][sl] for o in out]),
                    logdensities=np.stack([o["logdensities"][sl] for o in out]))

    def mcmc_from(self, q, N, out=None):
        cur = self.get_state(("q",))["q"]
        assert np.array_equal(np.asarray(q, float), cur), "dry run: mcmc_from only continues from the current positions"
        return self.mcmc(N)

    def get_state(self, fields=("q", "lq", "grad", "minv", "eps", "p")):
        if set(fields) <= {"q", "lq", "grad"} and not self.stages and self.q0 is not None:   # evaluate_ℓ at the set positions
            with self._ctx():
                ev = [self.po.logdensity_and_gradient(self.family, q, self.params, self.T) for q in self.q0]
            full = {"q": self.q0.copy(), "lq": np.array([e[0] for e in ev]), "grad": np.stack([e[1] for e in ev])}
            return {f: full[f] for f in fields}
        out = self._last if self._last is not None and self.n_drawn else self._run(self.n_drawn)
        st = {}
        for f in fields:
            if f == "q":
                st[f] = np.stack([o["posterior_matrix"][self.n_drawn - 1] if self.n_drawn else
                                  (o["warmup_posterior"][-1] if len(o["warmup_posterior"]) else o["q_final"]) for o in out])
            elif f == "eps":
                st[f] = np.array([o["eps"] for o in out])
            elif f == "minv":
                st[f] = np.stack([o["minv"] if o["minv"].ndim == 1 else np.ones(self.D) for o in out])
            else:
                raise NotImplementedError(f"dry run: get_state({f})")
        return st

    def metric_is_dense(self):
        dense = False
        for s in self.stages:                     # the kind of the LAST metric window
            if s[0] == self.po.STAGE_TUNING and s[2] != self.po.METRIC_NOTHING:
                dense = s[2] == self.po.METRIC_SYMMETRIC
        return dense

    def get_metric_dense(self):
        out = self._last or self._run(self.n_drawn)
        return np.stack([o["minv"] for o in out])

    def close(self):
        pass
