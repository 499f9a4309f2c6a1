"""Frozen golden vectors (tests/golden/oracle_vectors.json): oracle on CPU, CUDA path on GPU."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))["cases"]


def _f(h):
    return np.array([float.fromhex(x) for x in h]) if isinstance(h, list) else float.fromhex(h)


def test_oracle_reproduces_golden_vectors(po):
    for c in CASES:
        params = None if c["params"] is None else _f(c["params"])
        r = po.sample_tree(c["family"], _f(c["q"]), _f(c["eps"]), c["seed"], c["chain"], c["t"],
                           minv=_f(c["minv"]), params=params, T=c["T"])
        g = c["tree"]
        for f in ("depth", "left", "right", "steps", "directions"):
            assert int(r["stats"][f]) == g[f]
        assert r["stats"]["pi"] == _f(g["pi"]) and r["stats"]["acceptance_rate"] == _f(g["acceptance_rate"])
        assert np.array_equal(r["q"], _f(g["q"])) and r["lq"] == _f(g["lq"])
        p0 = po.normals(c["seed"], c["chain"], 2, c["t"], c["D"])
        ql, pl, _, lql = po.leapfrog(c["family"], _f(c["q"]), p0, _f(c["eps"]), minv=_f(c["minv"]), params=params,
                                     T=c["T"], n_steps=2)
        assert np.array_equal(ql, _f(c["leapfrog2"]["q"])) and np.array_equal(pl, _f(c["leapfrog2"]["p"]))


JULIA_FIXTURES = sorted(f for f in os.listdir(os.path.join(HERE, "golden")) if f.startswith("julia_") and f.endswith(".json"))
FAMILY_IDS = {"std_normal": 0, "diag_normal": 1, "funnel": 2, "logistic": 3}


@pytest.mark.skipif(not JULIA_FIXTURES, reason="no tests/golden/julia_*.json: julia/parity.jl has not been run (no Julia in the build image)")
@pytest.mark.parametrize("fname", JULIA_FIXTURES or ["-"])
def test_julia_fixtures_when_present(po, fname):
    """Vectors written by julia/parity.jl from the REAL DynamicHMC.jl (momenta and direction words injected through
    sample_tree's p= / directions= keywords, randexp from the engine's Philox stream): the oracle must reproduce every
    integer of TreeStatisticsNUTS exactly and the new position within 1e-10 relative (BASELINE.json north_star)."""
    fx = json.load(open(os.path.join(HERE, "golden", fname)))
    fam = FAMILY_IDS[fx["family"]]
    params = np.array(fx["params"], float) if fx["params"] else None
    for c in fx["cases"]:
        r = po.sample_tree(fam, np.array(c["q"], float), c["eps"], fx["seed"], c["chain"], c["t"], params=params,
                           p=np.array(c["p"], float), directions=c["directions"], T=32)
        for f in ("depth", "left", "right", "steps", "directions"):
            assert int(r["stats"][f]) == int(c["stats"][f]), (fname, c["chain"], c["t"], f)
        np.testing.assert_allclose(r["q"], np.array(c["q_new"], float), rtol=1e-10, atol=0)
        np.testing.assert_allclose(r["stats"]["pi"], c["stats"]["pi"], rtol=1e-10)
        np.testing.assert_allclose(r["stats"]["acceptance_rate"], c["stats"]["acceptance_rate"], rtol=1e-10, atol=1e-300)


@pytest.mark.gpu
def test_cuda_reproduces_golden_vectors(pkg, po):
    for c in CASES:
        D = c["D"]
        params = None if c["params"] is None else _f(c["params"])
        ℓ = [pkg.StandardNormal(D), None, pkg.Funnel(D)][c["family"]] if c["family"] != 1 else \
            pkg.DiagNormal(params[:D], 1.0 / params[D:])
        if c["family"] == 1:   # exact precisions, not 1/(1/x)
            ℓ.params = lambda p=params: p
        K = 3
        eng = pkg.Engine(ℓ, chains=K, seed=c["seed"], chain_offset=c["chain"] - 1)   # local chain 1 = golden chain
        if eng.layout()[0] != c["T"]:
            eng.close()
            continue
        q = np.tile(_f(c["q"]), (K, 1))
        eng.set_metric(np.tile(_f(c["minv"]), (K, 1)))
        eng.set_position(q)
        eng.set_stepsize(_f(c["eps"]))
        eng.transition_count = c["t"]
        p0 = np.stack([po.normals(c["seed"], c["chain"] - 1 + k, 2, c["t"], D) for k in range(K)])
        eng.set_momentum(p0)
        eng.leapfrog(2, 1)
        st = eng.get_state(("q", "p", "lq"))
        assert np.array_equal(st["q"][1], _f(c["leapfrog2"]["q"])) and np.array_equal(st["p"][1], _f(c["leapfrog2"]["p"]))
        assert st["lq"][1] == _f(c["leapfrog2"]["lq"])
        eng.set_position(q)
        stats = eng.sample_tree()
        g = c["tree"]
        for f in ("depth", "left", "right", "steps", "directions"):
            assert int(stats[1][f]) == g[f]
        assert stats[1]["pi"] == _f(g["pi"]) and stats[1]["acceptance_rate"] == _f(g["acceptance_rate"])
        assert np.array_equal(eng.get_state(("q",))["q"][1], _f(g["q"]))
        eng.close()
