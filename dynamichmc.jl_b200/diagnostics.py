"""Host-side mirror of DynamicHMC.Diagnostics (src/diagnostics.jl) over the returned
tree-statistics arrays, plus access to the device-side reduction (dhmc_tree_summary_dev)
that avoids shipping 56 B × N × chains of statistics to the host."""
from dataclasses import dataclass

import numpy as np

ACCEPTANCE_QUANTILES = [0.05, 0.25, 0.5, 0.75, 0.95]      # diagnostics.jl:35


def EBFMI(tree_statistics):
    """diagnostics.jl:29-32: mean(abs2, diff(πs)) / var(πs) for one chain's statistics."""
    pis = np.asarray(tree_statistics["pi"], float)
    return float(np.mean(np.diff(pis) ** 2) / np.var(pis, ddof=1))


def count_terminations(tree_statistics):
    """diagnostics.jl:65-81"""
    l, r = np.asarray(tree_statistics["left"]), np.asarray(tree_statistics["right"])
    maxd = (l == 1) & (r == 0)
    div = (l == r)
    return dict(max_depth=int(maxd.sum()), divergence=int(div.sum()), turning=int((~maxd & ~div).sum()))


def count_depths(tree_statistics):
    """diagnostics.jl:88-94 (first element is depth 0, trailing zeros dropped)"""
    c = np.bincount(np.asarray(tree_statistics["depth"]).ravel(), minlength=33)
    nz = np.nonzero(c)[0]
    return c[: nz[-1] + 1].tolist() if nz.size else []


@dataclass
class TreeStatisticsSummary:
    """diagnostics.jl:44-57"""
    N: int
    a_mean: float
    a_quantiles: list
    termination_counts: dict
    depth_counts: list


def summarize_tree_statistics(tree_statistics):
    """diagnostics.jl:100-106"""
    a = np.asarray(tree_statistics["acceptance_rate"], float).ravel()
    return TreeStatisticsSummary(int(a.size), float(a.mean()),
                                 [float(v) for v in np.quantile(a, ACCEPTANCE_QUANTILES)],
                                 count_terminations(tree_statistics), count_depths(tree_statistics))
