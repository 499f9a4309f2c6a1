"""Integration checks of the oracle's mcmc_with_warmup (test/test_mcmc.jl)."""
import numpy as np
import pytest


def test_mcmc_normal_moments(po):  # test_mcmc.jl:18-26 (5-dim N(1, I), default warmup)
    D, N = 5, 10000
    params = np.concatenate([np.ones(D), np.ones(D)])
    r = po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, N, seed=1, chain=0, params=params)
    Z = r["posterior_matrix"]
    assert np.linalg.norm(Z.mean(0) - 1, np.inf) < 0.04
    assert np.linalg.norm(Z.std(0, ddof=1) - 1, np.inf) < 0.04
    assert r["tree_statistics"]["acceptance_rate"].mean() >= 0.8
    assert 0.5 <= r["eps"] <= 2
    for q, ld in zip(Z[:50], r["logdensities"][:50]):
        assert po.logdensity_and_gradient(po.FAMILY_DIAG_NORMAL, q, params)[0] == ld


def test_fixed_stepsize_and_no_search(po):  # :28-48
    D, N = 5, 2000
    params = np.concatenate([np.ones(D), np.ones(D)])
    st = po.default_warmup_stages(search=False, dual_averaging=False)
    r = po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, N, seed=2, chain=0, params=params,
                            stages=st, eps0=1.0)
    assert r["eps"] == 1.0
    assert np.linalg.norm(r["posterior_matrix"].mean(0) - 1, np.inf) < 0.15
    st = po.default_warmup_stages(search=False)
    r = po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, N, seed=3, chain=0, params=params,
                            stages=st, eps0=1.0)
    assert 0.5 <= r["eps"] <= 2
    # InitialStepsizeSearch refuses to run when ϵ was supplied (mcmc.jl:137)
    with pytest.raises(po.OracleError) as e:
        po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, 10, seed=3, chain=0, params=params, eps0=1.0)
    assert e.value.status == 1


def test_default_warmup_is_900_transitions(po):  # mcmc.jl:415-425
    st = po.default_warmup_stages()
    assert sum(s[1] for s in st) == 900 and st[0][0] == po.STAGE_SEARCH
    assert [s[1] for s in st[1:]] == [75, 25, 50, 100, 200, 400, 50]


def test_no_max_depth_on_200_dim_normal(po):  # :60-72 (shortened: 2 chains × 300 draws)
    D = 200
    for chain in range(2):
        r = po.mcmc_with_warmup(po.FAMILY_STD_NORMAL, D, 300, seed=4, chain=chain, max_depth=12, T=64)
        ts = r["tree_statistics"]
        assert not np.any((ts["left"] == 1) & (ts["right"] == 0))     # REACHED_MAX_DEPTH
        assert ts["depth"].max() < 12


def test_welford_matches_two_pass_metric(po):
    # The device accumulates the window variance by Welford; the reference is
    # two-pass var (mcmc.jl:209).  Same draws ⇒ metrics agree to rounding.
    D = 20
    prec = 1 / np.linspace(0.1, 10, D)
    params = np.concatenate([np.zeros(D), prec])
    st = [(po.STAGE_SEARCH, 0, 0, 0), (po.STAGE_TUNING, 75, 0, 1), (po.STAGE_TUNING, 100, 1, 1)]
    a = po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, 5, seed=5, chain=1, params=params, stages=st,
                            welford=False)
    b = po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, 5, seed=5, chain=1, params=params, stages=st,
                            welford=True)
    assert np.allclose(a["minv"], b["minv"], rtol=1e-12)
    assert a["eps"] == b["eps"]          # ϵ adaptation precedes the metric switch


def test_adapted_metric_tracks_posterior_variance(po):
    D = 10
    var = np.logspace(-2, 2, D)
    params = np.concatenate([np.zeros(D), 1 / var])
    r = po.mcmc_with_warmup(po.FAMILY_DIAG_NORMAL, D, 200, seed=8, chain=0, params=params)
    assert np.all(np.abs(np.log(r["minv"] / var)) < 0.5)


def test_pooled_metric_is_the_shrunk_covariance_of_the_pooled_window(po):
    """The optional pooled-per-group metric (not reference semantics; include/dhmc.h DHMC_METRIC_SYMMETRIC_POOLED): the shared
    M⁻¹ must be regularize_M⁻¹(cov(pooled draws of the 8 chains), λ) (mcmc.jl:211,218-221 applied to the pooled window) up to
    rounding — here checked against numpy on the draws of a run whose LAST stage is the pooled window."""
    D, Nw, seed, chain0 = 5, 40, 9, 24
    stages = [(po.STAGE_SEARCH, 0, 0, 0), (po.STAGE_TUNING, 25, 0, 1), (po.STAGE_TUNING, Nw, po.METRIC_SYMMETRIC_POOLED, 1)]
    r = po.mcmc_with_warmup_pooled(po.FAMILY_STD_NORMAL, D, 3, seed, chain0, stages, T=32)
    # the window draws of every chain: rerun each chain alone with a per-chain Symmetric stage (same draws: the metric only
    # changes AFTER the window) and collect the warm-up draws of that stage
    st1 = [(po.STAGE_SEARCH, 0, 0, 0), (po.STAGE_TUNING, 25, 0, 1), (po.STAGE_TUNING, Nw, po.METRIC_SYMMETRIC, 1)]
    window = []
    for c in range(8):
        o = po.mcmc_with_warmup(po.FAMILY_STD_NORMAL, D, 1, seed, chain0 + c, stages=st1, T=32, welford=True, keep_warmup=True)
        window.append(o["warmup_posterior"][-Nw:])
    X = np.concatenate(window)                                   # [8·Nw, D]
    S = np.cov(X.T)
    lam = 5.0 / Nw
    expect = (1 - lam) * S + lam * np.diag(np.diag(S))
    np.testing.assert_allclose(r["minv"], expect, rtol=1e-10, atol=1e-13)
    assert np.all(r["eps"] > 0) and r["posterior_matrix"].shape == (8, 3, D)
