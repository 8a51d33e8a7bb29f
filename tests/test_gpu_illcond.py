"""GPU parity of das_letkf outside the benign conditioning regime.

The tensor-core solver computes A^-1/2 by a Newton-Schulz iteration; the reference diagonalises A with
EISPACK (common/common_mtx.f90:41-99, common/common_letkf.f90:147-206), which is insensitive to the
conditioning of A = (k-1)/rho I + Yr^T Y.  Real radar volumes are hundreds of observations of the same
storm: lambda_max(A)/c0 = 10^3..10^5.  These tests hand das_letkf (not core_batch) observation ensembles
with PRESCRIBED spectra, rank-deficient and full rank, for the three solver size classes used by the
BASELINE configs (k = 20, 50, 100), a "storm" case, and the H(x)-consistent correlated workload of bench.py.
Bar: anal3d, RTPS factor and adaptive inflation within 1e-10 of the oracle, or LETKF_B200_EEIGEN -- never a
silent wrong answer.  Above lambda_max/c0 = 10^6 the problem itself is conditioned worse than 1e-10 (the
eigenvalue c0 is known to eps * lambda_max at best, in the reference too), so the bar there is eps * ratio."""
import numpy as np
import pytest

import scale_letkf_b200 as sl
from scale_letkf_b200 import synth, capi
from helpers import TOL, relerr, host_logp, truth_analysis_point

pytestmark = pytest.mark.gpu

REPORT = []   # one line per case; written to gpurun_out/illcond_report.txt when the directory exists


@pytest.fixture(scope="module", autouse=True)
def _write_report():
    yield
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "illcond_report.txt"), "w") as f:
            f.write("\n".join(REPORT) + "\n")


def spectrum_case(k, ratio, rank, p, seed, storm=False):
    cfg = synth.config_c2(nlon=6, nlat=6, nlev=2, member=k)
    for t in range(24):
        cfg.HORI_LOCAL[t] = 5000.0e3    # every point sees every observation with a weight close to 1
        cfg.VERT_LOCAL[t] = 50.0
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=seed + 1)
    g = synth.rng(seed, 23)
    err = 1.0
    if storm:   # p rows = 5 shared member patterns + 5 % noise, sigma_b / sigma_o = 3
        pat = g.standard_normal((5, k))
        Y = 3.0 * err * (pat[g.integers(0, 5, p)] + 0.05 * g.standard_normal((p, k)))
        Y -= Y.mean(axis=1, keepdims=True)
    else:
        r = min(rank, k - 1)
        Q, _ = np.linalg.qr(np.column_stack([np.ones(k), g.standard_normal((k, k - 1))]))
        V = Q[:, 1:1 + r]                                   # orthonormal, orthogonal to ones
        U, _ = np.linalg.qr(g.standard_normal((p, r)))
        ev = ratio * (k - 1) * err * err * np.logspace(0.0, -np.log10(max(ratio, 10.0)), r)
        Y = (U * np.sqrt(ev)) @ V.T
    lo_i, hi_i = cfg.IHALO + 0.5, cfg.nlon + cfg.IHALO + 0.5
    obs = dict(elm=np.full(p, capi.ID_U, np.int32), typ=np.full(p, capi.TYP_ADPUPA, np.int32),
               ri=g.uniform(lo_i, hi_i, p), rj=g.uniform(lo_i, hi_i, p),
               lev=85000.0 * np.exp(g.uniform(-0.1, 0.1, p)), dat=np.zeros(p), err=np.full(p, err),
               val=2.0 * g.standard_normal(p), ensval=np.ascontiguousarray(Y))
    return cfg, rig1, rjg1, hgt1, obs, gues


def run_both(cfg, rig1, rjg1, hgt1, obs, gues, oracle, adaptive=False, truth=False):
    k = cfg.MEMBER
    if adaptive:
        cfg.INFL_MUL_ADAPTIVE = 1
        cfg.INFL_MUL = 1.05
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    e = sl.LETKF(cfg, device=0)
    e.set_letkf_obs(obs)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    i1 = i2 = None
    if adaptive:
        i1 = np.full((gues.shape[0], gues.shape[1], gues.shape[3]), cfg.INFL_MUL, order="F")
        i2 = i1.copy(order="F")
    ref = o.das_letkf(gues.copy(order="F"), infl3d=i1, want_rtps=True, want_nobsl=True)
    out = e.das_letkf(gues.copy(order="F"), infl3d=i2, want_rtps=True, want_nobsl=True, logp=host_logp(cfg, gues),
                      allow_eigen_fail=True)
    e.close()
    res = dict(status=out["status"], iters=out["solver_iterations"] / max(out["nsolved"], 1))
    if out["status"] == 0:
        assert np.array_equal(out["nobsl"], ref["nobsl"])
        res["anal"] = relerr(out["anal3d"][:, :, :k, :], ref["anal3d"][:, :, :k, :], axis=(0, 1, 2))
        res["rtps"] = relerr(out["rtps"], ref["rtps"])
        if adaptive:
            res["infl"] = relerr(i2, i1)
        if truth and not adaptive:
            # both implementations against an 80-bit evaluation of the same formulas at a few points: where the
            # problem itself is conditioned worse than 1e-10 the GPU must be as accurate as the reference algorithm
            s2o = o.sorted_index()
            ens_s, val_s = obs["ensval"][s2o], obs["val"][s2o]
            sc = np.abs(ref["anal3d"][:, :, :k, :]).max(axis=(0, 1, 2))
            eg = eo = 0.0
            for ij, il in ((0, 0), (17, 1), (35, 0)):
                pm = gues[ij, il, k, cfg.iv3d_p - 1]
                n, idx, rd, _ = o.obs_local([rig1[ij]], [rjg1[ij]], [pm], [hgt1[ij, il]], 1, len(s2o))
                pl = int(n[0])
                dx = gues[ij, il, :k, :] - gues[ij, il, k, :][None, :]
                xa, _ = truth_analysis_point(cfg, ens_s[idx[0, :pl], :k], rd[0, :pl], val_s[idx[0, :pl]], dx,
                                             gues[ij, il, k, :], infl=cfg.INFL_MUL)
                eg = max(eg, float((np.abs(out["anal3d"][ij, il, :k, :] - xa) / sc).max()))
                eo = max(eo, float((np.abs(ref["anal3d"][ij, il, :k, :] - xa) / sc).max()))
            res["gpu_vs_truth"], res["oracle_vs_truth"] = eg, eo
    return res


@pytest.mark.parametrize("k", [20, 50, 100])
@pytest.mark.parametrize("ratio", [1e2, 1e3, 1e4, 1e5, 1e6, 1e7])
@pytest.mark.parametrize("rank", ["low", "full"])
def test_prescribed_spectrum(oracle, k, ratio, rank, record_property):
    c = spectrum_case(k, ratio, 5 if rank == "low" else k - 1, 2 * k, seed=int(np.log10(ratio)) * 100 + k)
    adaptive = (k == 50 and ratio <= 1e5)
    r = run_both(*c, oracle, adaptive=adaptive, truth=ratio >= 1e5)
    record_property("iterations_per_solve", r["iters"])
    REPORT.append(f"k={k} ratio={ratio:.0e} rank={rank}: " + ", ".join(f"{a}={b:.3g}" for a, b in r.items()))
    if r["status"] == capi.EEIGEN:
        assert ratio >= 1e7, "EEIGEN is only acceptable where the reference's own truncation rule is near"
        return
    if ratio <= 1e5:
        assert r["anal"] <= TOL and r["rtps"] <= TOL
        if "infl" in r:
            assert r["infl"] <= TOL
    if "gpu_vs_truth" in r:   # never less accurate than the reference algorithm (x3 for the rounding lottery)
        assert r["gpu_vs_truth"] <= max(TOL, 3.0 * r["oracle_vs_truth"])
        assert r["anal"] <= max(TOL, 4.0 * (r["gpu_vs_truth"] + r["oracle_vs_truth"]))


@pytest.mark.parametrize("k", [20, 50, 100])
def test_storm(oracle, k):
    r = run_both(*spectrum_case(k, 0.0, 0, 1000, seed=900 + k, storm=True), oracle)
    REPORT.append(f"storm k={k}: " + ", ".join(f"{a}={b:.3g}" for a, b in r.items()))
    assert r["status"] == 0 and r["anal"] <= TOL and r["rtps"] <= TOL


@pytest.mark.parametrize("kind,member", [("sonde", 20), ("sonde", 50), ("radar", 24), ("radar", 100)])
def test_hx_consistent_workload(oracle, kind, member):
    """the bench workload generator at test size: members = smooth random fields + noise, ensval = H(x_m)"""
    if kind == "sonde":
        cfg = synth.config_c2(nlon=24, nlat=24, nlev=6, member=member)
        for t in range(24):
            cfg.HORI_LOCAL[t] = 60.0e3
        rig1, rjg1, hgt1 = synth.make_grid(cfg, topo_amp=300.0)
        se = synth.SmoothEnsemble(cfg, seed_no=31, topo_amp=300.0, hlen=4.0)
        obs = se.attach(synth.make_sonde_obs(cfg, 30, 100, nlevobs=8, seed_no=41))
    else:
        cfg = synth.config_c3(nlon=24, nlat=24, nlev=6, member=member, max_nobs=200 if member == 100 else 60)
        rig1, rjg1, hgt1 = synth.make_grid(cfg)
        se = synth.SmoothEnsemble(cfg, seed_no=32, hlen=8.0, vlen=2000.0)
        obs = se.attach(synth.make_radar_obs(cfg, radius_m=5.0e3, zmin=500.0, zmax=6000.0, dz=1000.0, seed_no=43))
    gues = se.state(rig1, rjg1, as_numpy=True)
    r = run_both(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    REPORT.append(f"hx {kind} k={member}: " + ", ".join(f"{a}={b:.3g}" for a, b in r.items()))
    assert r["status"] == 0 and r["anal"] <= TOL and r["rtps"] <= TOL
