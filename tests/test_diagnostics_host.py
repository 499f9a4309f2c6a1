"""Host logic of the trajectory diagnostics (dynamichmc.jl_b200/diagnostics.py, mirroring
src/diagnostics.jl:139-216) exercised on CPU against an oracle-backed stand-in for the engine.
The stand-in exists only in this test: it answers the handful of engine calls the diagnostics
make (set_position / set_momentum / set_stepsize / leapfrog / phase_logdensity / get_state) with
the oracle, so chain bookkeeping, ordering of the (ϵ, p) grid, stopping rules and argument
checks are covered without a GPU.  The same assertions run against the CUDA path in
tests/test_gpu_parity.py::test_trajectory_diagnostics_match_oracle."""
import sys

import numpy as np
import pytest

import __graft_entry__ as entry


class OracleEngine:
    def __init__(self, po, ℓ, chains, seed=0, device=0, **kw):
        self.po, self.ℓ, self.K, self.D = po, ℓ, chains, ℓ.dimension()
        self.params, self.minv, self.T = ℓ.params(), np.ones(self.D), 32

    def set_kinetic_energy(self, κ):
        self.minv = np.asarray(κ.minv, float)

    def set_position(self, q):
        self.q = np.array(q, float).reshape(self.K, self.D).copy()
        self.lq, self.g = np.empty(self.K), np.empty((self.K, self.D))
        for k in range(self.K):
            self.lq[k], self.g[k] = self.po.logdensity_and_gradient(self.ℓ.family, self.q[k], self.params, self.T)

    def set_momentum(self, p):
        self.p = np.array(p, float).reshape(self.K, self.D).copy()

    def set_stepsize(self, e):
        e = np.asarray(e, float).reshape(-1)
        self.eps = np.full(self.K, e[0]) if e.size == 1 else e.copy()

    def get_state(self, fields):
        return {f: getattr(self, f).copy() for f in fields}

    def phase_logdensity(self):
        return np.array([self.po.phase_logdensity(self.minv, self.lq[k], self.p[k], self.T) for k in range(self.K)])

    def leapfrog(self, n, sign):
        for k in range(self.K):
            self.q[k], self.p[k], self.g[k], self.lq[k] = self.po.leapfrog(
                self.ℓ.family, self.q[k], self.p[k], sign * self.eps[k], minv=self.minv, params=self.params,
                T=self.T, n_steps=n)

    def close(self):
        pass


@pytest.fixture()
def pkg_with_oracle_engine(monkeypatch):
    pkg = entry.load_package()
    po = entry.load_oracle()
    api = sys.modules[pkg.__name__ + ".api"]
    monkeypatch.setattr(api, "Engine", lambda ℓ, chains, **kw: OracleEngine(po, ℓ, chains, **kw))
    return pkg, po


def test_explore_log_acceptance_ratios_grid(pkg_with_oracle_engine):
    pkg, po = pkg_with_oracle_engine
    D = 11
    rng = np.random.default_rng(3)
    ℓ = pkg.DiagNormal(rng.normal(size=D), rng.uniform(0.3, 4, D))
    minv = rng.uniform(0.5, 2, D)
    κ = pkg.GaussianKineticEnergy(minv)
    q, ps = rng.normal(size=D), rng.normal(size=(3, D))
    log2eps = [-5, -2, 0, 1]
    A = pkg.diagnostics.explore_log_acceptance_ratios(ℓ, q, log2eps, κ=κ, ps=ps)
    assert A.shape == (4, 3)                       # [ϵ, p] as in the reference's comprehension
    for i, l2 in enumerate(log2eps):
        for j in range(3):
            assert A[i, j] == po.local_log_acceptance_ratio(ℓ.family, q, ps[j], 2.0 ** l2, minv=minv,
                                                            params=ℓ.params(), T=32)
    # default momenta: N of them, scaled by the metric
    assert pkg.diagnostics.explore_log_acceptance_ratios(ℓ, q, [-2.0], κ=κ, N=7, seed=3).shape == (1, 7)
    Minv = np.diag(minv) + 0.05
    assert pkg.diagnostics._rand_ps(pkg.GaussianKineticEnergy.symmetric(Minv), D, 5, 1).shape == (5, D)


def test_leapfrog_trajectory_positions_and_stopping(pkg_with_oracle_engine):
    pkg, po = pkg_with_oracle_engine
    D = 6
    rng = np.random.default_rng(4)
    ℓ = pkg.StandardNormal(D)
    q, p = rng.normal(size=D), rng.normal(size=D)
    traj = pkg.diagnostics.leapfrog_trajectory(ℓ, q, 0.2, range(-2, 4), p=p)
    assert [t["position"] for t in traj] == [-2, -1, 0, 1, 2, 3]
    assert traj[2]["Δ"] == 0.0 and np.array_equal(traj[2]["z"]["q"], q)
    lq0, _ = po.logdensity_and_gradient(ℓ.family, q, None, 32)
    π0 = po.phase_logdensity(None, lq0, p, 32)
    for t in traj:
        i = t["position"]
        if i:
            qo, p_o, _, lqo = po.leapfrog(ℓ.family, q, p, 0.2 if i > 0 else -0.2, n_steps=abs(i))
            assert np.array_equal(t["z"]["q"], qo) and np.array_equal(t["z"]["p"], p_o)
            assert t["Δ"] == po.phase_logdensity(None, lqo, p_o, 32) - π0
    # a huge step on the funnel overflows: the walk stops after the first non-finite log density
    f = pkg.Funnel(5)
    far = pkg.diagnostics.leapfrog_trajectory(f, np.array([-30.0, 5, 5, 5, 5]), 50.0, range(0, 6),
                                              p=np.ones(5))
    assert far[-1]["position"] < 5 and not np.isfinite(far[-1]["z"]["lq"])
    with pytest.raises(pkg.ArgumentError):
        pkg.diagnostics.leapfrog_trajectory(ℓ, q, 0.1, range(1, 4))


# ---------------------------------------------------------------- the reference's own diagnostics tests
def test_reference_kat_summarize_tree_statistics(pkg_with_oracle_engine):
    """test/test_diagnostics.jl:5-40 ("summarize tree statistics")."""
    pkg, _ = pkg_with_oracle_engine
    rng = np.random.default_rng(1)
    N = 1000
    ts = np.zeros(N, dtype=pkg._lib.tree_stats_dtype)
    ts["pi"] = rng.normal(size=N)
    ts["depth"] = rng.integers(0, 6, N)
    maxd = rng.uniform(size=N) < 0.1                          # REACHED_MAX_DEPTH = InvalidTree(1, 0)
    left = rng.integers(-5, 6, N)
    right = left + rng.integers(0, 6, N)
    ts["left"] = np.where(maxd, 1, left)
    ts["right"] = np.where(maxd, 0, right)
    ts["acceptance_rate"] = rng.uniform(size=N)
    ts["steps"] = rng.integers(1, 31, N)
    s = pkg.diagnostics.summarize_tree_statistics(ts)
    assert s.N == N and np.isclose(s.a_mean, ts["acceptance_rate"].mean())
    assert s.a_quantiles == [float(v) for v in np.quantile(ts["acceptance_rate"], pkg.diagnostics.ACCEPTANCE_QUANTILES)]
    div = int(((ts["left"] == ts["right"])).sum())            # is_divergent: left == right, trees.jl:197
    assert s.termination_counts["divergence"] == div and s.termination_counts["max_depth"] == int(maxd.sum())
    assert s.termination_counts["turning"] == N - div - int(maxd.sum())
    for i, c in enumerate(s.depth_counts):
        assert int((ts["depth"] == i).sum()) == c
    assert sum(s.depth_counts) == N
    assert 1.8 <= pkg.diagnostics.EBFMI(ts) <= 2.2            # "nonsensical value, just checking calculation"


def test_reference_kat_log_acceptance_ratios(pkg_with_oracle_engine):
    """test/test_diagnostics.jl:42-49."""
    pkg, _ = pkg_with_oracle_engine
    ℓ = pkg.DiagNormal(np.ones(5), np.ones(5))               # multivariate_normal(ones(5))
    log2eps = list(range(-5, 6))
    A = pkg.diagnostics.explore_log_acceptance_ratios(ℓ, np.zeros(5), log2eps, N=13)
    assert np.all(np.isfinite(A)) and A.shape == (len(log2eps), 13)


def test_reference_kat_leapfrog_trajectory(pkg_with_oracle_engine):
    """test/test_diagnostics.jl:51-80: a manually stepped trajectory against leapfrog_trajectory
    started from its fifth point."""
    pkg, po = pkg_with_oracle_engine
    K = 2
    ℓ = pkg.DiagNormal(np.ones(K), np.ones(K))
    params = ℓ.params()
    κ = pkg.GaussianKineticEnergy.identity(K)
    q, p, eps, ix0 = np.zeros(K), np.full(K, 0.98), 0.1, 5
    zs, πs = [], []
    for _ in range(15):
        lq, _ = po.logdensity_and_gradient(ℓ.family, q, params, 32)
        zs.append((q.copy(), p.copy()))
        πs.append(po.phase_logdensity(None, lq, p, 32))
        q, p, _, _ = po.leapfrog(ℓ.family, q, p, eps, params=params)
    Δs = np.array(πs) - πs[ix0 - 1]
    traj = pkg.diagnostics.leapfrog_trajectory(ℓ, zs[ix0 - 1][0], eps, range(1 - ix0, 15 - ix0 + 1), κ=κ,
                                               p=zs[ix0 - 1][1])
    assert len(traj) == 15
    np.testing.assert_allclose([t["Δ"] for t in traj], Δs, atol=1e-5)
    for t, (qq, pp) in zip(traj, zs):
        np.testing.assert_allclose(t["z"]["q"], qq, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(t["z"]["p"], pp, rtol=1e-8, atol=1e-12)


def test_ess_rhat_mirror_on_ar1_chains(pkg):
    """The numpy mirror of the device diagnostic on AR(1) chains with a known answer: ESS/N = (1 − φ)/(1 + φ); R̂ ≈ 1 for
    chains from the same distribution and clearly above 1 when half of the chains are shifted."""
    rng = np.random.default_rng(5)
    K, N, D = 64, 400, 3
    phis = np.array([0.0, 0.5, 0.8])
    x = np.empty((K, N, D))
    x[:, 0] = rng.normal(size=(K, D))
    for i in range(1, N):
        x[:, i] = phis * x[:, i - 1] + np.sqrt(1 - phis ** 2) * rng.normal(size=(K, D))
    r = pkg.diagnostics.ess_rhat(x, max_lag=60)
    assert np.all(np.abs(r["rhat"] - 1) < 0.03)
    expect = K * (N // 2 * 2) * (1 - phis) / (1 + phis)
    assert np.all(np.abs(r["ess"] / expect - 1) < 0.15), (r["ess"], expect)
    x[: K // 2] += 1.0
    assert np.all(pkg.diagnostics.ess_rhat(x, max_lag=60)["rhat"] > 1.1)
