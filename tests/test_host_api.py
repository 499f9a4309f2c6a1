"""Host logic of the reference-facing mirror (dynamichmc.jl_b200/api.py) on CPU: argument checks of
the algorithm structs (the reference's @argcheck sites), the warm-up stage fold, `initialization`
handling, result containers and layout helpers.  The engine is replaced by a recording stand-in
that lives only in this file, so what is tested is the host orchestration — the numeric path is
tested through the C ABI in tests/test_gpu_parity.py."""
import sys

import numpy as np
import pytest

import __graft_entry__ as entry


@pytest.fixture(scope="module")
def pkg():
    return entry.load_package()


class RecordingEngine:
    """Answers the calls mcmc_keep_warmup makes and records them."""
    calls = None

    def __init__(self, ℓ, chains, seed=0, algorithm=None, device=0, chain_offset=0, **opts):
        self.ℓ, self.K, self.D = ℓ, int(chains), int(ℓ.dimension())
        self.log = [("create", dict(seed=seed, algorithm=algorithm, device=device, chain_offset=chain_offset, **opts))]
        self.eps = np.full(self.K, 0.25)
        self.minv = np.ones((self.K, self.D))
        self.dense = False
        self.closed = False
        RecordingEngine.calls = self.log

    def set_kinetic_energy(self, κ):
        self.log.append(("set_kinetic_energy", κ))
        self.dense = κ.dense
        self.minv = np.broadcast_to(np.asarray(κ.minv, float), (self.K,) + np.asarray(κ.minv).shape[-(2 if κ.dense else 1):]).copy()

    def set_position(self, q):
        self.log.append(("set_position", np.array(q)))

    def random_position(self):
        self.log.append(("random_position",))

    def set_stepsize(self, e):
        self.log.append(("set_stepsize", e))
        self.eps = np.broadcast_to(np.asarray(e, float), (self.K,)).copy()

    def find_initial_stepsize(self, s):
        self.log.append(("find_initial_stepsize", s))

    def warmup_stage(self, stage, keep=False):
        self.log.append(("warmup_stage", stage, keep))
        if not keep:
            return None
        K, N, D = self.K, stage.N, self.D
        return {"posterior_matrix": np.zeros((K, N, D)), "tree_statistics": np.zeros((K, N)),
                "ϵs": np.zeros((K, N)), "logdensities": np.zeros((K, N))}

    def mcmc(self, N):
        self.log.append(("mcmc", N))
        K, D = self.K, self.D
        post = np.arange(K * N * D, dtype=float).reshape(K, N, D)
        return dict(posterior_matrix=post, tree_statistics=np.zeros((K, N)), logdensities=-post.sum(-1))

    def get_state(self, fields):
        return {"minv": self.minv if not self.dense else np.ones((self.K, self.D)), "eps": self.eps}

    def metric_is_dense(self):
        return self.dense

    def get_metric_dense(self):
        return self.minv

    def close(self):
        self.closed = True


@pytest.fixture()
def api(pkg, monkeypatch):
    mod = sys.modules[pkg.__name__ + ".api"]
    monkeypatch.setattr(mod, "Engine", RecordingEngine)
    return mod


# ---------------------------------------------------------------- argument checks (the reference's @argcheck sites)
def test_algorithm_struct_argchecks(pkg):
    E = pkg.ArgumentError
    pkg.NUTS(max_depth=32)
    for bad in (dict(max_depth=0), dict(max_depth=33), dict(min_Δ=0.0)):          # NUTS.jl:188-190
        with pytest.raises(E):
            pkg.NUTS(**bad)
    for bad in (dict(δ=0.0), dict(δ=1.0), dict(γ=0.0), dict(κ=0.5), dict(κ=1.01), dict(t0=-1)):   # stepsize.jl:112-116
        with pytest.raises(E):
            pkg.DualAveraging(**bad)
    for bad in (dict(initial_ϵ=0.0), dict(initial_ϵ=float("inf")), dict(log_threshold=0.0),
                dict(log_threshold=float("-inf")), dict(maxiter_crossing=49)):       # stepsize.jl:30-34
        with pytest.raises(E):
            pkg.InitialStepsizeSearch(**bad)
    for bad in (dict(N=19), dict(N=20, λ=-0.1), dict(N=20, M="Dense")):              # mcmc.jl:190-193
        with pytest.raises(E):
            pkg.TuningNUTS(**bad)
    assert pkg.TuningNUTS(50).λ == 5.0 / 50                                          # mcmc.jl:187


def test_default_and_fixed_stepsize_warmup_stages(pkg):
    st = pkg.default_warmup_stages()                                                 # mcmc.jl:415-425
    assert isinstance(st[0], pkg.InitialStepsizeSearch)
    assert [s.N for s in st[1:]] == [75, 25, 50, 100, 200, 400, 50] and sum(s.N for s in st[1:]) == 900
    assert [s.M for s in st[1:]] == [None] + [pkg.Diagonal] * 5 + [None]
    st = pkg.default_warmup_stages(M=pkg.Symmetric, doubling_stages=2, stepsize_search=None)
    assert st[0] is None and [s.M for s in st[1:]] == [None, pkg.Symmetric, pkg.Symmetric, None]
    fx = pkg.fixed_stepsize_warmup_stages(middle_steps=30, doubling_stages=3)        # mcmc.jl:436-440
    assert [s.N for s in fx] == [30, 60, 120]
    assert all(isinstance(s.stepsize_adaptation, pkg.FixedStepsize) and s.M == pkg.Diagonal for s in fx)


# ---------------------------------------------------------------- the warm-up fold and initialization
def test_warmup_fold_calls_the_engine_in_stage_order(pkg, api):
    ℓ = pkg.StandardNormal(4)
    stages = (None, pkg.InitialStepsizeSearch(initial_ϵ=0.3), pkg.TuningNUTS(20), pkg.TuningNUTS(25, M=pkg.Diagonal))
    r = api.mcmc_keep_warmup(11, ℓ, 7, chains=3, warmup_stages=stages, algorithm=pkg.NUTS(max_depth=6),
                             chain_offset=40, engine_opts=dict(threads_per_chain=32))
    log = RecordingEngine.calls
    assert log[0] == ("create", dict(seed=11, algorithm=pkg.NUTS(max_depth=6), device=0, chain_offset=40,
                                     threads_per_chain=32))
    assert [c[0] for c in log[1:]] == ["random_position", "find_initial_stepsize", "warmup_stage", "warmup_stage", "mcmc"]
    assert log[2][1].initial_ϵ == 0.3 and log[3][1].N == 20 and log[4][1].M == pkg.Diagonal and log[5] == ("mcmc", 7)
    assert [w["stage"] for w in r["warmup"]] == list(stages)                         # no-op stage kept, mcmc.jl:99-101
    assert r["warmup"][0]["results"] is None and r["warmup"][1]["results"] is None
    assert r["warmup"][2]["results"]["posterior_matrix"].shape == (3, 20, 4)
    assert not r["engine"].closed                                                    # keep_warmup hands the engine back
    with pytest.raises(pkg.ArgumentError):
        api.mcmc_keep_warmup(1, ℓ, 5, warmup_stages=("bogus",))


def test_reporter_hook_and_stepwise_sampling(pkg, api):
    """`reporter` (mcmc.jl:279,378 call sites; one report per batch here) and mcmc_steps / mcmc_next_step (mcmc.jl:335-351)."""
    ℓ = pkg.StandardNormal(4)
    seen = []
    stages = (pkg.InitialStepsizeSearch(), pkg.TuningNUTS(20), None)
    api.mcmc_with_warmup(3, ℓ, 9, chains=2, warmup_stages=stages, reporter=lambda msg, **kw: seen.append((msg, kw)))
    assert [m for m, _ in seen] == ["warmup stage finished"] * 3 + ["inference finished"]
    assert seen[1][1] == dict(stage=2, of=3, kind="TuningNUTS", transitions=20, chains=2)
    assert seen[3][1] == dict(transitions=9, chains=2)
    # a reporter that raises never aborts sampling
    api.mcmc_with_warmup(3, ℓ, 2, chains=2, warmup_stages=(), reporter=lambda *a, **k: 1 / 0)

    class StepEngine(RecordingEngine):
        def mcmc_from(self, q, N, out=None):
            self.log.append(("mcmc_from", np.array(q), N))
            K, D = self.K, self.D
            st = np.zeros((K, N), dtype=[("depth", "<i8")]); st["depth"] = 3
            return dict(posterior_matrix=np.repeat(np.asarray(q)[:, None, :] + 1.0, N, axis=1), tree_statistics=st,
                        logdensities=np.zeros((K, N)))

        def get_state(self, fields):
            return {"q": np.full((self.K, self.D), 0.5)} if tuple(fields) == ("q",) else super().get_state(fields)

    eng = StepEngine(ℓ, 3)
    steps = api.mcmc_steps(eng)
    assert isinstance(steps, api.MCMCSteps) and np.array_equal(steps.Q, np.full((3, 4), 0.5))
    Q1, stats = api.mcmc_next_step(steps, steps.Q)
    assert Q1.shape == (3, 4) and np.all(Q1 == 1.5) and stats.shape == (3,) and np.all(stats["depth"] == 3)
    assert eng.log[-1][0] == "mcmc_from" and eng.log[-1][2] == 1
    Q2, _ = api.mcmc_next_step(steps, Q1)
    assert np.all(Q2 == 2.5)


def test_user_model_from_source_and_gradient_check_host_logic(pkg, api, monkeypatch, tmp_path):
    """UserLogDensity.from_source writes the header once (content-addressed); diagnostics.check_gradient lays out q, q ± h eᵢ
    as 2·D + 1 chains and differences the returned ℓ values (engine replaced by a closed-form stand-in)."""
    monkeypatch.setattr(api, "_CSRC", str(tmp_path))
    src = "#define DHMC_USER_NSUMS 0\n/* ... */\n"
    a = api.UserLogDensity.from_source(src, "toy", 3, params=[1.0], library="prebuilt.so")
    b = api.UserLogDensity.from_source(src, "toy", 3, library="prebuilt.so")
    assert a.header == b.header and open(a.header).read() == src and a.library_path == "prebuilt.so"
    assert a.dimension() == 3 and list(a.params()) == [1.0] and a.family == pkg._lib.FAMILY_USER
    with pytest.raises(pkg.ArgumentError):
        api.UserLogDensity.from_source(src, "not a name", 3, library="x")

    class QuadEngine:                                       # ℓ(q) = -½ Σ c_i q_i², ∇ℓ = -c q  (c = 1, 2, 3)
        def __init__(self, ℓ, chains, **kw):
            self.K = chains
        def set_position(self, Q):
            c = np.array([1.0, 2.0, 3.0])
            self.lq = -0.5 * (np.asarray(Q) ** 2 @ c)
            self.g = -np.asarray(Q) * c
        def get_state(self, fields):
            return {"lq": self.lq, "grad": self.g}
        def close(self):
            self.closed = True
    monkeypatch.setattr(api, "Engine", QuadEngine)
    r = pkg.diagnostics.check_gradient(a, [0.5, -1.0, 2.0])
    assert np.allclose(r["grad"], [-0.5, 2.0, -6.0]) and np.allclose(r["fd"], r["grad"], rtol=1e-8)
    assert r["max_abs_err"] < 1e-8 and r["lq"] == pytest.approx(-0.5 * (0.25 + 2.0 + 12.0))


def test_checkpoint_and_restore_host_logic(pkg, api, tmp_path):
    """Engine.checkpoint / restore / save_checkpoint / load_checkpoint: (Q, κ, ϵ, RNG counter), diagonal and Symmetric κ,
    through a file; the primitives are replaced by recorders (the device round trip: test_gpu_parity.py)."""
    class FakeEngine(pkg.Engine):                           # the real class (the `api` fixture swaps api.Engine for a recorder)
        def __init__(self, K, D, dense):
            self.K, self.D, self._dense, self.log, self._t = K, D, dense, [], 7
        def get_state(self, fields):
            return {"q": np.arange(self.K * self.D, dtype=float).reshape(self.K, self.D), "minv": np.full((self.K, self.D), 2.0),
                    "eps": np.linspace(0.1, 0.2, self.K)}
        def metric_is_dense(self): return self._dense
        def get_metric_dense(self): return np.tile(np.eye(self.D) * 3.0, (self.K, 1, 1))
        def set_metric(self, m): self.log.append(("set_metric", np.array(m)))
        def set_metric_dense(self, m): self.log.append(("set_metric_dense", np.array(m)))
        def set_position(self, q): self.log.append(("set_position", np.array(q)))
        def set_stepsize(self, e): self.log.append(("set_stepsize", np.array(e)))
        transition_count = property(lambda self: self._t, lambda self, t: self.log.append(("transition_count", t)))
        def close(self): pass
    for dense in (False, True):
        a = FakeEngine(3, 4, dense)
        path = str(tmp_path / f"ck{int(dense)}.npz")
        a.save_checkpoint(path)
        b = FakeEngine(3, 4, dense)
        b.load_checkpoint(path)
        assert [c[0] for c in b.log] == ["set_metric_dense" if dense else "set_metric", "set_position", "set_stepsize", "transition_count"]
        assert b.log[0][1].shape == ((3, 4, 4) if dense else (3, 4)) and np.array_equal(b.log[1][1], a.get_state(("q",))["q"])
        assert np.array_equal(b.log[2][1], np.linspace(0.1, 0.2, 3)) and b.log[3][1] == 7
        with pytest.raises(pkg.ArgumentError):
            FakeEngine(5, 4, dense).load_checkpoint(path)


def test_initialization_fields(pkg, api):
    ℓ = pkg.StandardNormal(3)
    κ = pkg.GaussianKineticEnergy(np.array([1.0, 2.0, 3.0]))
    api.mcmc_with_warmup(5, ℓ, 4, chains=2, warmup_stages=(), initialization={"q": [0.1, 0.2, 0.3], "κ": κ, "ϵ": 0.5})
    log = RecordingEngine.calls
    assert [c[0] for c in log[1:]] == ["set_kinetic_energy", "set_position", "set_stepsize", "mcmc"]
    assert log[2][1].shape == (2, 3) and np.array_equal(log[2][1][1], [0.1, 0.2, 0.3])   # one q for all chains
    # keyword spelling: Python NFKC-normalises ϵ to ε; "eps" is accepted as well
    for init in (dict(ϵ=0.5), dict(eps=0.5), {"ε": 0.5}):
        api.mcmc_with_warmup(5, ℓ, 4, chains=2, warmup_stages=(), initialization=init)
        assert ("set_stepsize", 0.5) in RecordingEngine.calls and ("random_position",) in RecordingEngine.calls
    with pytest.raises(pkg.ArgumentError):
        api.mcmc_with_warmup(5, ℓ, 4, initialization={"momentum": 1})


# ---------------------------------------------------------------- results and layout helpers
def test_results_views_and_posterior_layouts(pkg, api):
    ℓ = pkg.StandardNormal(3)
    K, N, D = 4, 5, 3
    res = api.mcmc_with_warmup(5, ℓ, N, chains=K, warmup_stages=(), initialization={"ϵ": [0.1, 0.2, 0.3, 0.4]})
    assert len(res) == K
    r2 = res[2]
    assert r2["posterior_matrix"].shape == (D, N)                                    # [parameter, draw], mcmc.jl:230
    assert np.shares_memory(r2["posterior_matrix"], res._post)                       # a view, not a copy
    assert r2["ϵ"] == 0.3 and r2["eps"] == 0.3 and not r2["κ"].dense and r2["κ"].minv.shape == (D,)
    assert np.array_equal(r2["logdensities"], -res._post[2].sum(-1))
    stacked = pkg.stack_posterior_matrices(res)                                      # [draw, chain, parameter], mcmc.jl:602-604
    assert stacked.shape == (N, K, D) and stacked[1, 2, 0] == r2["posterior_matrix"][0, 1]
    pooled = pkg.pool_posterior_matrices(res)                                        # [parameter, draw ⊗ chain], mcmc.jl:614-616
    assert pooled.shape == (D, K * N) and np.array_equal(pooled[:, N * 2 + 1], r2["posterior_matrix"][:, 1])
    # dense metric comes back as one [D, D] matrix per chain
    κ = pkg.GaussianKineticEnergy.symmetric(np.eye(D) * 2.0)
    res = api.mcmc_with_warmup(5, ℓ, N, chains=2, warmup_stages=(), initialization={"κ": κ})
    assert res[0]["κ"].dense and res[0]["κ"].minv.shape == (D, D)


def test_model_types_mirror_logdensityproblems(pkg):
    rng = np.random.default_rng(0)
    for ℓ in (pkg.StandardNormal(5), pkg.DiagNormal(rng.normal(size=5), rng.uniform(0.5, 2, 5)), pkg.Funnel(5),
              pkg.LogisticRegression(rng.normal(size=(30, 5)), (rng.uniform(size=30) < 0.5).astype(float))):
        assert ℓ.dimension() == 5 and ℓ.capabilities() >= 1                          # hamiltonian.jl:146
        q = rng.normal(size=5) * 0.3
        _, g = ℓ.logdensity_and_gradient(q)
        h = 1e-6
        num = np.array([(ℓ.logdensity_and_gradient(q + h * e)[0] - ℓ.logdensity_and_gradient(q - h * e)[0]) / (2 * h)
                        for e in np.eye(5)])
        np.testing.assert_allclose(g, num, rtol=1e-5, atol=1e-7)
    with pytest.raises(pkg.ArgumentError):
        pkg.DiagNormal(np.zeros(3), np.ones(4))
