"""The device-resident CG of SURVEY 8f / N1 is a restatement of ALGLIB's mincg
(super-resolution_b200/csrc/srb_cg.h) over a vector backend.  Here its control flow is pinned on
the CPU: instantiated over host arrays with ALGLIB's summation orders (oracle/cg_host_harness.cpp),
it must reproduce the reference's own ALGLIB bit for bit -- every objective value the line searches
ask for, the iterate, the iteration / evaluation counts and the termination type -- against the
committed fixture (tests/golden/cg_golden.npz, made by tools/make_cg_golden.py) and, where
oracle/_ref is present, against ALGLIB run live."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cg_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = cg_cases.cases()


@pytest.fixture(scope="module")
def host_cg():
    path = os.path.join(ROOT, "oracle", "_build", "libsrb_cg_host.so")
    src = os.path.join(ROOT, "oracle", "cg_host_harness.cpp")
    hdr = os.path.join(ROOT, "super-resolution_b200", "csrc", "srb_cg.h")
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/libsrb_cg_host.so"])
    return C.CDLL(path)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "cg_golden.npz"))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_restatement_reproduces_alglib_fixture(host_cg, golden, case):
    name, fg, x0, kw = case
    x, rep, trace = cg_cases.run(host_cg.srbcg_host_minimize, x0, fg, **kw)
    np.testing.assert_array_equal(rep[:4], golden[name + "_report"])
    np.testing.assert_array_equal(trace, golden[name + "_trace"])
    np.testing.assert_array_equal(x, golden[name + "_x"])
    assert rep[5] == len(trace)


def test_cases_cover_the_termination_rules_and_restarts(host_cg, golden):
    kinds = {int(golden[c[0] + "_report"][2]) for c in CASES}
    assert {1, 2, 4, 5} <= kinds
    restarts = 0
    for name, fg, x0, kw in CASES:
        restarts += cg_cases.run(host_cg.srbcg_host_minimize, x0, fg, **kw)[1][4]
    assert restarts > 0


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_restatement_reproduces_alglib_live(host_cg, case):
    from oracle import sr_ref
    if not sr_ref.available():
        pytest.skip("oracle/_ref (the reference's ALGLIB) is not built here")
    name, fg, x0, kw = case
    xr, rr, tr = cg_cases.run(sr_ref.lib().ref_mincg, x0, fg, **kw)
    xh, rh, th = cg_cases.run(host_cg.srbcg_host_minimize, x0, fg, **kw)
    np.testing.assert_array_equal(rh[:4], rr[:4])
    np.testing.assert_array_equal(th, tr)
    np.testing.assert_array_equal(xh, xr)
