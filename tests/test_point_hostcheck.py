"""The per-point arithmetic of the CUDA material kernel (exaconstit_b200/csrc/material_point.hpp), compiled for
the host by tests/hostcheck, against the oracle point by point -- no GPU needed.  This pins the kernel's
mathematics (including the solver branches a GPU parity run rarely visits: dogleg steps, rejected trials,
the pivoted-LU fallback, the exp/log power-law path) in the CPU suite; the GPU parity tests then only have to
show that the device build of the same source agrees."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import hotpath_cases as hc

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        d = os.path.join(HERE, "hostcheck")
        subprocess.check_call(["make", "-s", "-C", d])
        _lib = C.CDLL(os.path.join(d, "build", "libhostcheck.so"))
        _lib.hostcheck_model_setup.restype = C.c_long
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def run_host_point_update(case, force_pivot=0, disable_powi=0, layout=1):
    from oracle import orc
    c = case
    xend = c["xbeg"] + c["dt"] * c["vel"]
    jac = orc.jacobians(c["G"], orc.gather(c["e2n"], xend))
    velE = orc.gather(c["e2n"], c["vel"])
    ne, nsv = c["ne"], c["nsv"]
    s1, h1, mg = np.zeros(ne * 48), np.zeros(ne * 8 * nsv), np.zeros(ne * 8 * 36)
    props = np.ascontiguousarray(c["props"], dtype=np.float64)
    G = np.ascontiguousarray(c["G"])
    nfev = C.c_long(0)
    st = (C.c_long * 8)()
    lib().hostcheck_stats(st)
    nfail = lib().hostcheck_model_setup(c["xtal"], c["kin"], _p(props), props.size, force_pivot, disable_powi, C.c_long(ne),
                                        C.c_double(c["dt"]), _p(jac), _p(G), _p(velE), _p(c["stress0"]), _p(c["hist0"]),
                                        _p(s1), _p(h1), _p(mg), C.byref(nfev), layout)
    if layout == 2:   # compact records -> Voigt 6x6 for the comparison
        full = np.zeros_like(mg)
        lib().hostcheck_compact_expand(C.c_long(ne * 8), _p(mg), _p(full))
        mg = full
    lib().hostcheck_stats(st)
    names = ("trials", "unused", "rejac", "rejected", "pivot_loop", "dogleg", "pivot_tangent")
    return dict(nfail=nfail, stress1=s1, hist1=h1, matgrad=mg, nfev=nfev.value, stats=dict(zip(names, list(st))))


def _check(case, out, cpu, tol=1e-11):
    assert out["nfail"] == 0 and cpu["nfail"] == 0
    assert hc.rel_err(out["stress1"], cpu["stress1"]) < tol
    nsv = case["nsv"]
    hg, hcpu = out["hist1"].reshape(-1, nsv), cpu["hist1"].reshape(-1, nsv)
    for c in range(nsv):
        scale = max(np.abs(hcpu[:, c]).max(), 1e-12)
        assert np.abs(hg[:, c] - hcpu[:, c]).max() / scale < max(tol, 1e-10), c
    assert hc.rel_err(out["matgrad"], cpu["matgrad"]) < 1e-9
    # identical iterates: the evaluation counts agree point by point
    assert np.array_equal(hg[:, 3], hcpu[:, 3])


@pytest.mark.parametrize("n,xtal,kin", [(3, 0, 0), (4, 0, 0), (4, 1, 0), (3, 0, 1), (4, 1, 2), (3, 0, 2), (3, 2, 2)])
def test_point_update_matches_oracle(n, xtal, kin):
    case = hc.make_case(n=n, seed=10 + n, ngrains=5, xtal=xtal, kin=kin)
    cpu = hc.run_oracle_hot_path(case)
    out = run_host_point_update(case)
    _check(case, out, cpu)
    assert out["stats"]["pivot_loop"] == 0 and out["stats"]["pivot_tangent"] == 0
    if kin != 2:
        props = np.ascontiguousarray(case["props"], dtype=np.float64)
        assert lib().hostcheck_pl_n(xtal, kin, _p(props), props.size) == 49  # 1/m - 1 with m = 0.02


@pytest.mark.parametrize("force_pivot,disable_powi", [(1, 0), (0, 1)])
def test_point_update_alternate_paths(force_pivot, disable_powi):
    """pivoted-LU fallback everywhere / exp-log power law instead of repeated squaring"""
    case = hc.make_case(n=3, seed=77, ngrains=4)
    cpu = hc.run_oracle_hot_path(case)
    out = run_host_point_update(case, force_pivot=force_pivot, disable_powi=disable_powi)
    _check(case, out, cpu)
    if force_pivot:
        assert out["stats"]["pivot_loop"] > 0 and out["stats"]["pivot_tangent"] == case["ne"] * 8


@pytest.mark.parametrize("xtal,kin,rate,dt", [(0, 0, 5e-2, 0.5), (0, 0, 1.0, 0.1), (2, 2, 2e-2, 1.0)])
def test_point_update_dogleg_paths(xtal, kin, rate, dt):
    """large increments: the trust region is active (dogleg / Cauchy steps, rejected trials)"""
    case = hc.make_case(n=3, seed=5, ngrains=4, xtal=xtal, kin=kin, rate=rate, dt=dt)
    cpu = hc.run_oracle_hot_path(case)
    out = run_host_point_update(case)
    _check(case, out, cpu, tol=1e-10)
    st = out["stats"]
    assert st["rejected"] > 0 and st["dogleg"] > 0 and st["trials"] == out["nfev"]


@pytest.mark.parametrize("xtal,kin", [(0, 0), (1, 0), (0, 1), (1, 2), (0, 2)])
def test_compact_tangent_record_expands_to_the_voigt_tangent(xtal, kin):
    """layout 2 (the private 32-double record the fused PA gradient apply streams: 5x5 deviatoric operator,
    -dp/dlnV, deviatoric stress) expands to exactly the 6x6 the reference layout holds"""
    case = hc.make_case(n=3, seed=21, ngrains=4, xtal=xtal, kin=kin)
    cpu = hc.run_oracle_hot_path(case)
    out = run_host_point_update(case, layout=2)
    _check(case, out, cpu)
    ref = run_host_point_update(case, layout=1)
    assert hc.rel_err(out["matgrad"], ref["matgrad"]) < 1e-13
