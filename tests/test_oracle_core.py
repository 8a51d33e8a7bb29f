"""Pins for the CPU oracle of letkf_core / mtx_eigen / rs.

The reference holds no golden vectors for this path, so the oracle is pinned by
(1) analytic known answers, (2) algebraic invariants, (3) an independent LAPACK solver.
"""
import numpy as np
import pytest
import scipy.linalg

from scale_letkf_b200 import synth


def test_pythag(oracle):
    for a, b in [(3.0, 4.0), (0.0, 0.0), (1e200, 1e200), (1e-200, 3e-200), (-5.0, 12.0)]:
        assert oracle.lib().oracle_pythag(a, b) == pytest.approx(np.hypot(a, b), rel=1e-15)


@pytest.mark.parametrize("n", [1, 2, 3, 20, 50, 100])
def test_rs_vs_lapack(oracle, n):
    g = synth.rng(90, n)
    y = g.standard_normal((3 * n, n))
    a = y.T @ y + (n - 1) * np.eye(n)
    ierr, w, z = oracle.rs(a)
    assert ierr == 0
    w_ref = scipy.linalg.eigh(a, eigvals_only=True)
    assert np.allclose(w, w_ref, rtol=1e-13, atol=0)
    assert np.allclose(z @ np.diag(w) @ z.T, a, rtol=0, atol=1e-12 * np.abs(a).max())
    assert np.allclose(z.T @ z, np.eye(n), atol=1e-13)


def test_mtx_eigen_descending(oracle):
    g = synth.rng(91)
    y = g.standard_normal((30, 12))
    a = y.T @ y + 11 * np.eye(12)
    nrank, w, z = oracle.mtx_eigen(a)
    assert nrank == 12
    assert np.all(np.diff(w) <= 0)
    assert np.allclose(a @ z, z * w[None, :], atol=1e-11)


def test_kat_p0(oracle):
    """nobsl == 0: trans = sqrt(infl) I, transm = 0, pao = infl/(k-1) I (common_letkf.f90:89-107)."""
    k = 7
    r = oracle.letkf_core(np.zeros((4, k)), np.ones(4), np.ones(4), np.zeros(4), 1.21, nobsl=0,
                          rdiag_wloc=True, depd=np.zeros(4))
    assert np.array_equal(r["trans"], np.sqrt(1.21) * np.eye(k))
    assert np.array_equal(r["transm"], np.zeros(k))
    assert np.array_equal(r["transmd"], np.zeros(k))
    assert np.allclose(r["pao"], 1.21 / (k - 1) * np.eye(k), rtol=1e-16)


def test_kat_p1_sherman_morrison(oracle):
    """p = 1: Pa = rho/(k-1) [I - y y^T / (r (k-1)/rho + y^T y)]."""
    g = synth.rng(92)
    k, rho, r_ = 9, 1.1, 0.7
    y = g.standard_normal(k)
    y -= y.mean()
    d = 0.3
    res = oracle.letkf_core(y[None, :], np.array([r_]), np.array([1.0]), np.array([d]), rho,
                            rdiag_wloc=True)
    c = (k - 1) / rho
    pa = (1.0 / c) * (np.eye(k) - np.outer(y, y) / (r_ * c + y @ y))
    assert np.allclose(res["pao"], pa, rtol=1e-12, atol=1e-15)
    assert np.allclose(res["transm"], pa @ (y / r_) * d, rtol=1e-12, atol=1e-15)
    w = scipy.linalg.sqrtm((k - 1) * pa).real
    assert np.allclose(res["trans"], w, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("k,p", [(20, 100), (20, 7), (50, 200), (10, 3)])
def test_invariants_random(oracle, k, p):
    g = synth.rng(93, k * 1000 + p)
    y = g.standard_normal((p, k))
    y -= y.mean(axis=1, keepdims=True)
    rloc = np.exp(-0.5 * g.uniform(0, 13.3, p))
    rdiag = 1.0 / rloc
    dep = g.standard_normal(p)
    infl = 1.05
    res = oracle.letkf_core(y, rdiag, rloc, dep, infl, rdiag_wloc=True)
    a = y.T @ (y / rdiag[:, None]) + (k - 1) / infl * np.eye(k)
    pa, w, wm = res["pao"], res["trans"], res["transm"]
    assert np.allclose(pa @ a, np.eye(k), atol=1e-11)
    assert np.allclose(w, w.T, atol=1e-12)
    assert np.allclose(w @ w, (k - 1) * pa, rtol=1e-10, atol=1e-13)
    # zero-mean perturbations => A 1 = (k-1)/rho 1 => W 1 = sqrt(rho) 1
    assert np.allclose(w @ np.ones(k), np.sqrt(infl) * np.ones(k), rtol=1e-11)
    assert np.allclose(wm, pa @ (y.T @ (dep / rdiag)), rtol=1e-10, atol=1e-13)
    # independent solver: LAPACK eigh
    lam, v = scipy.linalg.eigh(a)
    assert np.allclose(pa, (v / lam) @ v.T, rtol=1e-11, atol=1e-14)
    assert np.allclose(w, (v * np.sqrt((k - 1) / lam)) @ v.T, rtol=1e-11, atol=1e-14)


def test_rdiag_wloc_false_applies_rloc(oracle):
    g = synth.rng(94)
    k, p = 8, 15
    y = g.standard_normal((p, k))
    rloc = g.uniform(0.1, 1.0, p)
    err2 = g.uniform(0.5, 2.0, p)
    dep = g.standard_normal(p)
    a = oracle.letkf_core(y, err2, rloc, dep, 1.0, rdiag_wloc=False)
    b = oracle.letkf_core(y, err2 / rloc, rloc, dep, 1.0, rdiag_wloc=True)
    assert np.allclose(a["trans"], b["trans"], rtol=1e-12, atol=1e-14)
    assert np.allclose(a["transm"], b["transm"], rtol=1e-12, atol=1e-14)


def test_transm_absent_adds_mean_weight(oracle):
    g = synth.rng(95)
    k, p = 6, 9
    y = g.standard_normal((p, k))
    rd = g.uniform(0.5, 2.0, p)
    dep = g.standard_normal(p)
    a = oracle.letkf_core(y, rd, np.ones(p), dep, 1.0, rdiag_wloc=True, want_transm=True)
    b = oracle.letkf_core(y, rd, np.ones(p), dep, 1.0, rdiag_wloc=True, want_transm=False)
    assert np.allclose(b["trans"], a["trans"] + a["transm"][:, None], rtol=1e-14, atol=1e-15)


def test_adaptive_inflation_formula(oracle):
    g = synth.rng(96)
    k, p = 10, 40
    y = g.standard_normal((p, k))
    rloc = g.uniform(0.2, 1.0, p)
    rdiag = 1.0 / rloc
    dep = 1.5 * g.standard_normal(p)
    infl = 1.1
    r = oracle.letkf_core(y, rdiag, rloc, dep, infl, rdiag_wloc=True, infl_update=True)
    p1 = np.sum(dep * dep / rdiag)
    p2 = np.sum((y / rdiag[:, None]) * y) / (k - 1)
    p3 = rloc.sum()
    p4 = (p1 - p3) / p2 - infl
    sigma_o = 2.0 / p3 * ((infl * p2 + p3) / p2) ** 2
    gain = 0.04 ** 2 / (sigma_o + 0.04 ** 2)
    assert r["parm_infl"] == pytest.approx(infl + gain * p4, rel=1e-12)


def test_core_batch_matches_single(oracle):
    c = synth.make_core_batch(ne=8, npts=40, nobs=12, seed_no=97, det=True)
    r = oracle.core_batch(c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"],
                          c["parm_infl"], depd=c["depd"])
    assert r["status"] == 0
    for i in (0, 1, 5, 39):
        s = oracle.letkf_core(c["hdxb"][i].T, c["rdiag"][i], c["rloc"][i], c["dep"][i], 1.0,
                              nobsl=int(c["nobsl"][i]), rdiag_wloc=True, depd=c["depd"][i])
        assert np.array_equal(r["trans"][i].T, s["trans"])
        assert np.array_equal(r["transm"][i], s["transm"])
        assert np.array_equal(r["transmd"][i], s["transmd"])


def test_quickselect_vs_sort(oracle):
    g = synth.rng(98)
    for n, K in [(10, 3), (1000, 100), (1000, 999), (57, 1), (500, 250), (5000, 100)]:
        A = g.uniform(size=n)
        for desc in (False, True):
            X = (g.permutation(n) + 1).astype(np.int32)
            oracle.quickselect_arg(A, X, 1, n, K, desc=desc)
            assert sorted(X.tolist()) == list(range(1, n + 1))   # still a permutation
            sel = np.sort(A[X[:K] - 1])
            ref = np.sort(A)[::-1][:K][::-1] if desc else np.sort(A)[:K]
            assert np.array_equal(sel, ref)
