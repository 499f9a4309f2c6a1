"""Known answers of the reference's tree tests (test/test_trees.jl), run against
the recursive oracle (oracle/oracle.hpp: adjacent_tree, sample_trajectory)."""
import itertools
import math

import numpy as np
import pytest


def _testl(z):  # test_trees.jl:106
    return -abs(z - 3) ** 2 * 0.1


def _testA(zs):  # test_trees.jl:109
    return sum(min(math.exp(_testl(z)), 1) for z in zs)


def test_directions(po):  # test_trees.jl:8-17
    assert po.next_directions(0b110101, 6) == [True, False, True, False, True, True]
    assert 0 <= po.directions(1, 2, 3) < 2 ** 32


def test_dummy_adjacent_tree_full(po):  # :114-124
    r = po.dummy_adjacent_tree(0, 0, 2, True)
    assert r["valid"] and r["adjacency_ok"]
    assert r["zeta"][0] == range(1, 5)
    assert np.exp(r["zeta"][1]).sum() == pytest.approx(1)
    assert r["visited"] == [1, 2, 3, 4]
    assert not r["tau"][0]
    assert r["v"][0] == pytest.approx(_testA(r["visited"])) and r["v"][1] == 4
    assert r["z_last"] == r["i_last"] == 4


def test_dummy_adjacent_tree_turning(po):  # :126-133
    r = po.dummy_adjacent_tree(0, 0, 3, True, turning=range(5, 8))
    assert r["visited"] == [1, 2, 3, 4, 5, 6]
    assert not r["valid"] and r["invalid"] == (5, 6)
    assert r["v"][0] == pytest.approx(_testA(r["visited"]), rel=1e-15) and r["v"][1] == 6


def test_dummy_adjacent_tree_divergent(po):  # :135-142
    r = po.dummy_adjacent_tree(0, 0, 3, True, divergent=range(5, 8))
    assert r["visited"] == [1, 2, 3, 4, 5]
    assert not r["valid"] and r["invalid"] == (5, 5)
    assert r["v"][0] == pytest.approx(_testA(range(1, 6))) and r["v"][1] == 5


def test_dummy_adjacent_tree_full_backward(po):  # :144-154
    r = po.dummy_adjacent_tree(0, 0, 3, False)
    assert r["valid"] and r["adjacency_ok"]
    assert r["zeta"][0] == range(-8, 0)
    assert np.exp(r["zeta"][1]).sum() == pytest.approx(1)
    assert r["visited"] == [-k for k in range(1, 9)]
    assert not r["tau"][0]
    assert r["v"][0] == pytest.approx(_testA([-k for k in range(1, 9)])) and r["v"][1] == 8
    assert r["z_last"] == r["i_last"] == -8


def test_dummy_backward_turning_is_unsorted(po):
    # sub-tree turning returns InvalidTree(i′, i₊) in build order
    r = po.dummy_adjacent_tree(0, 0, 2, False, turning=[-1, -2])
    assert not r["valid"] and r["invalid"] == (-1, -2)


def test_dummy_sampled_tree(po):  # :156-165
    r = po.dummy_sample_trajectory(0, 3, 0b101)
    assert r["visited"] == [1, -1, -2, 2, 3, 4, 5]
    assert r["zeta"][0] == range(-2, 6)
    assert np.exp(r["zeta"][1]).sum() == pytest.approx(1)
    assert r["termination"] == (1, 0)  # REACHED_MAX_DEPTH
    assert r["v"][0] == pytest.approx(_testA(r["visited"])) and r["v"][1] == 7
    assert r["depth"] == 3 and r["adjacency_ok"]


# ---- detailed balance by exhaustive enumeration — test_trees.jl:167-262
def _logaddexp(a, b):
    return float(np.logaddexp(a, b))


def visited_log_probabilities(po, z, depth, **kw):
    acc = {}
    for flags in range(2 ** depth):
        r = po.dummy_sample_trajectory(z, depth, flags, **kw)
        assert r["adjacency_ok"]
        for zz, p in zip(r["zeta"][0], r["zeta"][1]):
            acc[zz] = _logaddexp(acc[zz], p) if zz in acc else p
    D = math.log(0.5) * depth
    return {k: v + D for k, v in acc.items()}


def transition_log_probability(po, z, z1, depth, **kw):
    p = -math.inf
    for flags in range(2 ** depth):
        r = po.dummy_sample_trajectory(z, depth, flags, **kw)
        zs = list(r["zeta"][0])
        if z1 in zs:
            p = _logaddexp(p, r["zeta"][1][zs.index(z1)])
    return p + depth * math.log(0.5)


def test_transition_consistency(po):  # :229-236
    for z1, pi in visited_log_probabilities(po, 9, 5).items():
        assert pi == pytest.approx(transition_log_probability(po, 9, z1, 5), abs=1e-12)


@pytest.mark.parametrize("kw,z,depths", [
    (dict(), 0, range(1, 6)),
    (dict(turning=[1, 2]), 3, range(1, 6)),
    (dict(divergent=[10, 11]), 3, range(1, 7)),
    (dict(divergent=[10, 11, 12], turning=[-3, -2]), 3, range(1, 7)),
])
def test_detailed_balance(po, kw, z, depths):  # :238-262
    atol = math.sqrt(np.finfo(float).eps)
    for depth in depths:
        lz = _testl(z)
        for z1, pi in visited_log_probabilities(po, z, depth, **kw).items():
            pi1 = transition_log_probability(po, z1, z, depth, **kw)
            assert pi + lz == pytest.approx(pi1 + _testl(z1), abs=atol)
