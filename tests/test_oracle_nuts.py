"""Known answers of test/test_NUTS.jl, test/test_hamiltonian.jl and
test/test_stepsize.jl against the oracle."""
import math

import numpy as np
import pytest


# --------------------------------------------------------------- test_NUTS.jl
def test_random_booleans(po):  # :10-21
    for prob in np.arange(1, 10) / 10:
        lp = math.log(prob)
        hits = [po.rand_bool_logprob(11, 0, 0, j, lp) for j in range(10000)]
        assert all(c == 1 for _, c in hits)
        assert abs(np.mean([b for b, _ in hits]) - prob) <= 0.02
    # logprob ≥ 0 never touches the RNG
    for lp in (0.0, 10.0):
        for j in range(100):
            assert po.rand_bool_logprob(11, 0, 0, j, lp) == (True, 0)


def test_low_level_turn_statistics(po):  # :27-42
    p = np.ones(3)
    c = 0.1
    t1 = np.stack([p, p - c, p, p - c, p])
    t2 = np.stack([3 * p, 3 * p + c, 3 * p, 3 * p + c, 3 * p])
    t3 = np.stack([2 * p, 2 * p + c, 2 * p, 2 * p + c, -2 * p])
    turning, rho = po.combine_turn_statistics(t1, t2)
    assert not turning and np.array_equal(rho, t1[4] + t2[4])
    turning, _ = po.combine_turn_statistics(t1, t3)
    assert turning


def test_turn_statistic_nan_is_not_turning(po):
    # NaN dots compare false ⇒ not turning
    p = np.ones(3)
    t1 = np.stack([p, p, p, p, p])
    t2 = np.stack([p, p, p, p, np.full(3, np.nan)])
    turning, _ = po.combine_turn_statistics(t1, t2)
    assert not turning


def test_low_level_visited_statistics(po):  # :44-55
    assert po.acceptance_rate([math.log(0.3)], [0]) == pytest.approx(0.3)
    assert po.acceptance_rate([math.log(0.6)], [0]) == pytest.approx(0.6)
    d = [math.log(0.3), math.log(0.3), math.log(0.6), math.log(10)]
    assert po.acceptance_rate(d, [0, 0, 0, 1]) == pytest.approx(0.4)


def test_unconditional_divergence(po):  # :58-85
    K = 3
    r = po.sample_tree(po.FAMILY_STD_NORMAL, np.zeros(K), 1.0, seed=5, chain=0, t=0,
                       always_divergent=True)
    s = r["stats"]
    assert s["left"] == s["right"]            # is_divergent
    assert s["acceptance_rate"] == 0
    assert s["depth"] == 0 and s["steps"] == 1
    assert np.array_equal(r["q"], np.zeros(K))


def test_sample_tree_mean_and_cov(po):  # :87-111 (diagonal Σ; κ perfectly adapted)
    rng = np.random.default_rng(4)
    for rep in range(3):
        K = int(rng.integers(2, 9))
        N = 10000
        mu = rng.normal(size=K)
        var = rng.uniform(0.2, 3.0, size=K)
        params = np.concatenate([mu, 1 / var])
        q = rng.normal(size=K)
        qs = np.empty((N, K))
        for i in range(N):
            r = po.sample_tree(po.FAMILY_DIAG_NORMAL, q, 0.5, seed=100 + rep, chain=0, t=i,
                               minv=var, params=params)
            q = r["q"]
            qs[i] = q
        C = np.cov(qs.T)
        tol = np.max(np.diag(C)) / 50
        assert np.sum(np.abs(qs.mean(0) - mu)) <= tol + tol * np.sum(np.abs(mu))
        assert np.allclose(C, np.diag(var), atol=0.1, rtol=0.1)


# -------------------------------------------------------- test_hamiltonian.jl
def test_leapfrog_calculation(po):  # :69-110
    def leapfrog_gaussian(q, p, grad, eps, m):
        u = np.sqrt(1 / m)
        ph = p + eps / 2 * grad(q)
        q1 = q + eps * u * (u * ph)
        p1 = ph + eps / 2 * grad(q1)
        return q1, p1

    rng = np.random.default_rng(0)
    n = 3
    m = rng.uniform(0.5, 2.0, n)          # diag of M; κ = GaussianKineticEnergy(inv(M))
    mu = rng.normal(size=n)
    prec = rng.uniform(0.5, 2.0, n)
    params = np.concatenate([mu, prec])
    grad = lambda x: -(x - mu) * prec
    q, p = rng.normal(size=n), rng.normal(size=n)
    eps = 0.1
    q0, p0 = q.copy(), p.copy()
    qo, po_ = q.copy(), p.copy()
    for i in range(100):
        q, p = leapfrog_gaussian(q, p, grad, eps, m)
        qo, po_, g, lq = po.leapfrog(po.FAMILY_DIAG_NORMAL, qo, po_, eps, minv=1 / m, params=params)
        assert np.allclose(qo, q, rtol=1e-9, atol=1e-12)
        assert np.allclose(po_, p, rtol=1e-9, atol=1e-12)
        lq2, g2 = po.logdensity_and_gradient(po.FAMILY_DIAG_NORMAL, qo, params)
        assert lq2 == lq and np.array_equal(g2, g)        # cache consistency :49-67
    assert np.array_equal(q0, q0) and np.array_equal(p0, p0)


def test_invalid_position_throws(po):  # :111-115
    with pytest.raises(po.OracleError) as e:
        po.evaluate_l(po.FAMILY_STD_NORMAL, np.full(3, np.nan))
    assert e.value.status == 2


def test_hamiltonian_invariance_and_reversibility(po):  # :118-177
    rng = np.random.default_rng(2)
    for _ in range(50):
        n = int(rng.integers(2, 6))
        mu = rng.normal(size=n)
        prec = rng.uniform(0.3, 3.0, n)
        params = np.concatenate([mu, prec])
        minv = rng.uniform(0.3, 3.0, n)
        q, p = rng.normal(size=n), rng.normal(size=n)
        eps = po.find_initial_stepsize(po.FAMILY_DIAG_NORMAL, q, p, minv=minv, params=params)
        lq0, _ = po.logdensity_and_gradient(po.FAMILY_DIAG_NORMAL, q, params)
        pi0 = po.phase_logdensity(minv, lq0, p)
        qq, pp = q, p
        for i in range(10):
            qq, pp, g, lq = po.leapfrog(po.FAMILY_DIAG_NORMAL, qq, pp, eps / 100, minv=minv, params=params)
            assert abs(po.phase_logdensity(minv, lq, pp) - pi0) < 0.5
        # back and forth
        q1, p1, _, _ = po.leapfrog(po.FAMILY_DIAG_NORMAL, q, p, 0.1, minv=minv, params=params)
        q2, p2, _, _ = po.leapfrog(po.FAMILY_DIAG_NORMAL, q1, p1, -0.1, minv=minv, params=params)
        assert np.max(np.abs(p2 - p)) < 1e-5 and np.max(np.abs(q2 - q)) < 1e-6


def test_infinity_fallbacks(po):  # :196-200
    one = np.ones(1)
    assert po.phase_logdensity(one, -np.inf, one) == -np.inf
    assert po.phase_logdensity(one, np.nan, one) == -np.inf
    assert po.phase_logdensity(one, 9.0, np.array([np.nan])) == -np.inf
    assert po.phase_logdensity(one, 9.0, np.array([2.0])) == 7.0


def test_funnel_gradient_matches_finite_differences(po):
    rng = np.random.default_rng(6)
    q = rng.normal(size=10)
    lq, g = po.logdensity_and_gradient(po.FAMILY_FUNNEL, q)
    ref = -q[0] ** 2 / 18 - 0.5 * np.exp(-q[0]) * np.sum(q[1:] ** 2) - 4.5 * q[0]
    assert lq == pytest.approx(ref, rel=1e-14)
    for i in range(10):
        h = 1e-6
        qp, qm = q.copy(), q.copy()
        qp[i] += h
        qm[i] -= h
        fd = (po.logdensity_and_gradient(po.FAMILY_FUNNEL, qp)[0] -
              po.logdensity_and_gradient(po.FAMILY_FUNNEL, qm)[0]) / (2 * h)
        assert g[i] == pytest.approx(fd, rel=1e-6, abs=1e-7)


# ----------------------------------------------------------- test_stepsize.jl
def test_stepsize_general_rootfinding(po):  # :9-25
    lt = math.log(0.8)
    for bad in (dict(log_threshold=float("nan")), dict(log_threshold=1.0),
                dict(initial_eps=-0.5), dict(maxiter=2)):
        with pytest.raises(po.OracleError) as e:
            po.search_params_check(**bad)
        assert e.value.status == 1          # ArgumentError
    A = lambda e: -3.0 * e
    eps = po.find_initial_stepsize_affine(-3.0, 0.0)
    assert A(eps) > lt > A(0.1)
    eps = po.find_initial_stepsize_affine(-3.0, 0.0, initial_eps=0.01)
    assert A(eps) < lt < A(0.01)
    with pytest.raises(po.OracleError) as e:
        po.find_initial_stepsize_affine(0.0, 1.0)   # constant ⇒ DynamicHMCError
    assert e.value.status == 2


def _dummy_acceptance_rate(rng, eps, sigma=0.05):  # :33
    return min(1 / eps * math.exp(rng.normal() * sigma - sigma ** 2 / 2), 1)


@pytest.mark.parametrize("eps0,n,sigma,atol", [(100.0, 500, 0.05, 0.02), (2.0, 2000, 0.05, 0.01),
                                               (20.0, 10000, 2.0, 0.04)])
def test_dual_averaging(po, eps0, n, sigma, atol):  # :37-71
    rng = np.random.default_rng(9)
    delta = 0.65
    A = po.da_init(eps0)
    assert A[4] == 0 and A[1] == 1 and A[2] == 0          # logϵ̄ = 0, m = 1, H̄ = 0  (:41-44)
    assert A[3] == pytest.approx(math.log(eps0)) and A[0] == pytest.approx(math.log(10) + math.log(eps0))
    for _ in range(n):
        A = po.da_adapt(A, _dummy_acceptance_rate(rng, math.exp(A[3]), sigma), delta=delta)
    final = math.exp(A[4])
    mean_rate = np.mean([_dummy_acceptance_rate(rng, final, sigma) for _ in range(10000)])
    assert abs(mean_rate - delta) < atol


def test_dual_averaging_recurrence_exact(po):
    # adapt_stepsize — src/stepsize.jl:147-156, restated in Python
    A = po.da_init(0.3)
    mu, m, Hb, le, leb = A
    rng = np.random.default_rng(1)
    for _ in range(50):
        a = rng.uniform()
        A = po.da_adapt(A, a)
        m += 1
        Hb += (0.8 - a - Hb) / (m + 10)
        le = mu - math.sqrt(m) / 0.05 * Hb
        leb += m ** (-0.75) * (le - leb)
        assert A[1] == m and A[2] == Hb and A[3] == le
        assert A[4] == pytest.approx(leb, rel=1e-13)
    with pytest.raises(po.OracleError):
        po.da_adapt(A, 1.5)


def test_find_reasonable_stepsize_brackets(po):  # :82-91
    rng = np.random.default_rng(10)
    lt = math.log(0.8)
    for _ in range(100):
        n = int(rng.integers(3, 6))
        mu, prec = rng.normal(size=n), rng.uniform(0.3, 3, n)
        params = np.concatenate([mu, prec])
        minv = rng.uniform(0.3, 3, n)
        q, p = rng.normal(size=n), rng.normal(size=n) / np.sqrt(minv)
        eps = po.find_initial_stepsize(po.FAMILY_DIAG_NORMAL, q, p, minv=minv, params=params)
        A = lambda e: po.local_log_acceptance_ratio(po.FAMILY_DIAG_NORMAL, q, p, e, minv=minv, params=params)
        bkt = lambda C: (A(eps) - lt) * (A(eps * C) - lt) <= 0
        assert bkt(0.5) or bkt(2.0)


def test_error_for_nonfinite_initial_density(po):  # :93-98
    with pytest.raises(po.OracleError) as e:
        po.find_initial_stepsize(po.FAMILY_STD_NORMAL, np.zeros(2), np.full(2, np.nan))
    assert e.value.status == 2


def test_logistic_regression_model(po):
    """LOGISTIC family: ℓ and ∇ℓ against numpy, and finite differences."""
    rng = np.random.default_rng(12)
    N, p = 300, 7
    X = rng.normal(size=(N, p)) / np.sqrt(p)
    y = (rng.uniform(size=N) < 0.5).astype(float)
    beta = rng.normal(size=p)
    params = po.logistic_params(X, y)
    lq, g = po.logdensity_and_gradient(po.FAMILY_LOGISTIC, beta, params)
    eta = X @ beta
    ref = np.sum(y * eta - np.logaddexp(0, eta)) - 0.5 * beta @ beta
    gref = X.T @ (y - 1 / (1 + np.exp(-eta))) - beta
    assert lq == pytest.approx(ref, rel=1e-13) and np.allclose(g, gref, rtol=1e-12, atol=1e-13)
    r = po.mcmc_with_warmup(po.FAMILY_LOGISTIC, p, 300, seed=3, chain=0, params=params)
    assert r["tree_statistics"]["acceptance_rate"].mean() > 0.6
    assert np.all(np.abs(r["posterior_matrix"].mean(0)) < 3.0)
