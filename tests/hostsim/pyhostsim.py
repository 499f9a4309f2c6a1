"""ctypes wrapper over tests/hostsim/libhostsim.so: the product's flattened NUTS
state machine (nuts_machine.cuh) compiled for the host with a plain-loop
backend.  Test harness only — see hostsim.cpp."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

tree_stats_dtype = np.dtype(
    [("pi", "<f8"), ("depth", "<i8"), ("left", "<i8"), ("right", "<i8"),
     ("acceptance_rate", "<f8"), ("steps", "<i8"), ("directions", "<u4"), ("pad", "<u4")])


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-C", _HERE, "libhostsim.so"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(os.path.join(_HERE, "libhostsim.so"))
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def run(family, q, eps, seed, chain, t0=0, N=1, minv=None, params=None, T=32, max_depth=10,
        min_delta=-1000.0, adapt=None, metric=0, p=None, directions=None):
    q = _d(q).copy()
    D = q.size
    minv = np.ones(D) if minv is None else _d(minv).copy()
    pr = _d(params) if params is not None else np.zeros(1)
    post = np.empty((N, D))
    stats = np.zeros(N, dtype=tree_stats_dtype)
    logd = np.empty(N)
    eps_used = np.empty(N)
    eps_out = C.c_double()
    lq = C.c_double()
    g = np.empty(D)
    ad = None if adapt is None else _d(adapt)
    pp = None if p is None else _d(p)
    dd = None if directions is None else np.array([directions], dtype=np.uint32)
    status = lib().hs_run(C.c_int(family), C.c_int(D), _p(pr), C.c_int(T), _p(minv),
                          C.c_int(max_depth), C.c_double(min_delta), C.c_uint64(seed),
                          C.c_uint64(chain), C.c_uint32(t0), C.c_int(N), _p(q), C.c_double(eps),
                          _p(ad), C.c_int(metric), _p(pp), _p(dd), _p(post), _p(stats), _p(logd),
                          _p(eps_used), C.byref(eps_out), C.byref(lq), _p(g))
    return dict(q=q, lq=lq.value, g=g, minv=minv, eps=eps_out.value, status=status,
                posterior_matrix=post, tree_statistics=stats, logdensities=logd, eps_used=eps_used)


def find_initial_stepsize(family, q, seed, chain, minv=None, params=None, T=32, p=None,
                          initial_eps=0.1, log_threshold=float(np.log(0.8)), maxiter=400):
    q = _d(q)
    D = q.size
    minv = np.ones(D) if minv is None else _d(minv)
    pr = _d(params) if params is not None else np.zeros(1)
    pp = None if p is None else _d(p)
    eps = C.c_double()
    status = lib().hs_find_initial_stepsize(C.c_int(family), C.c_int(D), _p(pr), C.c_int(T),
                                            _p(minv), C.c_uint64(seed), C.c_uint64(chain), _p(q),
                                            _p(pp), C.c_double(initial_eps),
                                            C.c_double(log_threshold), C.c_int(maxiter),
                                            C.byref(eps))
    return eps.value, status


def dummy_sample_trajectory(z, max_depth, flags, turning=(), divergent=()):
    """The flattened machine on the reference's DummyTrajectory (test/test_trees.jl:28-103)."""
    t = np.ascontiguousarray(list(turning), dtype=np.int64)
    d = np.ascontiguousarray(list(divergent), dtype=np.int64)
    vis = np.zeros(4096, dtype=np.int64)
    nv, depth = C.c_int(), C.c_int()
    left, right, steps = C.c_long(), C.c_long(), C.c_long()
    lib().hs_dummy_sample_trajectory(C.c_long(z), C.c_int(max_depth), C.c_uint32(flags), _p(t), C.c_int(t.size),
                                     _p(d), C.c_int(d.size), _p(vis), C.byref(nv), C.byref(depth),
                                     C.byref(left), C.byref(right), C.byref(steps))
    return dict(visited=vis[:nv.value].tolist(), depth=depth.value, termination=(left.value, right.value),
                steps=steps.value)
