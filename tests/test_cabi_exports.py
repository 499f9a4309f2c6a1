"""The C-ABI library loads without a GPU and exports every symbol include/dhmc.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dhmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dhmc_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported(pkg):
    lib = pkg._lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dhmc.h but not exported"
    assert sorted(pkg._lib.EXPORTS) == names


def test_struct_layouts_match_header(pkg):
    assert pkg._lib.tree_stats_dtype.itemsize == 56          # TreeStatisticsNUTS, NUTS.jl:208-221
    assert ctypes.sizeof(pkg._lib.Config) == 64
    assert ctypes.sizeof(pkg._lib.DualAveragingC) == 32


def test_argument_checks_without_gpu(pkg):
    # @argcheck mirrors run before any CUDA call
    with pytest.raises(pkg.ArgumentError):
        pkg.NUTS(max_depth=0)
    with pytest.raises(pkg.ArgumentError):
        pkg.NUTS(min_Δ=1.0)
    with pytest.raises(pkg.ArgumentError):
        pkg.DualAveraging(δ=1.5)
    with pytest.raises(pkg.ArgumentError):
        pkg.InitialStepsizeSearch(maxiter_crossing=2)
    with pytest.raises(pkg.ArgumentError):
        pkg.TuningNUTS(10)
    st = pkg.default_warmup_stages()
    assert isinstance(st[0], pkg.InitialStepsizeSearch)
    assert [s.N for s in st[1:]] == [75, 25, 50, 100, 200, 400, 50]
    assert [s.M for s in st[1:]] == [None] + [pkg.Diagonal] * 5 + [None]
    assert [s.N for s in pkg.fixed_stepsize_warmup_stages()] == [25, 50, 100, 200, 400]


def test_fails_loudly_without_cuda(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Engine(pkg.StandardNormal(10), chains=4)


def test_device_models_match_oracle_models(pkg, po):
    import numpy as np
    rng = np.random.default_rng(0)
    q = rng.normal(size=10)
    for ℓ, fam, params in ((pkg.StandardNormal(10), 0, None),
                           (pkg.DiagNormal(rng.normal(size=10), rng.uniform(0.5, 2, 10)), 1, "p"),
                           (pkg.Funnel(10), 2, None),
                           (pkg.LogisticRegression(rng.normal(size=(30, 10)), (rng.uniform(size=30) < 0.5) * 1.0), 3, "p")):
        pr = ℓ.params() if params else None
        lq, g = po.logdensity_and_gradient(fam, q, pr)
        lq2, g2 = ℓ.logdensity_and_gradient(q)
        assert lq == pytest.approx(lq2, rel=1e-13) and np.allclose(g, g2, rtol=1e-13)
