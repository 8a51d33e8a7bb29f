"""NOBS_OUT fields of das_letkf (scale/letkf/letkf_tools.f90:440-447, 767-778; obs_local's nobsl_t / cutd_t :1380-1390,
:1473-1475, :1653-1660): the oracle's restatement against an independent brute-force numpy evaluation (CPU), and
letkf_b200_nobs_out against the oracle (GPU, host and device buffers)."""
import os
import sys

import numpy as np
import pytest

from helpers import radar_case, sonde_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

REF, RE0, VR = 4001, 4004, 4002


def _beta(cfg, ri, rj, rz, radar_only):
    zcut = cfg.RADAR_ZMAX + max(cfg.VERT_LOCAL[21], cfg.VERT_LOCAL_RADAR_VR) * cfg.dist_zero_fac
    if radar_only and rz > zcut:
        return 0.0
    if cfg.BOUNDARY_BUFFER_WIDTH > 0.0:
        d = min(min(ri - cfg.IHALO, cfg.nlon + cfg.IHALO + 1 - ri) * cfg.DX,
                min(rj - cfg.JHALO, cfg.nlat + cfg.JHALO + 1 - rj) * cfg.DY) / cfg.BOUNDARY_BUFFER_WIDTH
        if d < 1.0:
            return max(d, 0.0)
    return 1.0


def _ndist(cfg, obs, n, ri, rj, rz):
    e = obs["elm"][n]
    hl = cfg.HORI_LOCAL_RADAR_OBSNOREF if e == RE0 else cfg.HORI_LOCAL_RADAR_VR if e == VR else cfg.HORI_LOCAL[21]
    vl = cfg.VERT_LOCAL_RADAR_VR if e == VR else cfg.VERT_LOCAL[21]
    nd_v = abs(obs["lev"][n] - rz) / vl
    nd_h = np.sqrt(((ri - obs["ri"][n]) * cfg.DX) ** 2 + ((rj - obs["rj"][n]) * cfg.DY) ** 2) / hl
    return nd_h * nd_h + nd_v * nd_v


@pytest.mark.parametrize("max_nobs", [12, 0])
def test_oracle_nobs_out_matches_bruteforce(oracle, max_nobs):
    from make_golden import select_bruteforce
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=6, nlon=24, nlat=24, nlev=5, max_nobs=max_nobs, seed=902, radius=5.0e3)
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    k = cfg.MEMBER
    pmean = np.asfortranarray(gues[:, :, k, cfg.iv3d_p - 1])
    out, hits = o.nobs_out(4, pmean)
    nij1, nlev = hgt1.shape
    dzf = cfg.dist_zero_fac
    checked = full = 0
    for il in range(nlev):
        for ij in range(0, nij1, 7):
            ri, rj, rz = rig1[ij], rjg1[ij], hgt1[ij, il]
            got = out[ij, il, :]
            if _beta(cfg, ri, rj, rz, True) == 0.0:
                assert np.array_equal(got, np.zeros(11))
                checked += 1
                continue
            ids = select_bruteforce(cfg, obs, (np.array([ri]), np.array([rj]), np.array([pmean[ij, il]]), np.array([rz])), 4)[0]
            e = obs["elm"][ids]
            nref, nre0, nvr = int((e == REF).sum()), int((e == RE0).sum()), int((e == VR).sum())
            hl_ref, hl_vr = cfg.HORI_LOCAL[21], cfg.HORI_LOCAL_RADAR_VR
            want = np.zeros(11)
            if max_nobs > 0:      # merged budget: the count of the group at the master type, nothing at the merged one
                want[4] = nref + nre0 + nvr
                want[5:8] = [nref + nre0, 0, nvr]
                want[8:11] = [hl_ref * dzf, 0.0, hl_vr * dzf]
                if nref + nre0 == max_nobs:
                    want[8] = hl_ref * np.sqrt(max(_ndist(cfg, obs, n, ri, rj, rz) for n in ids if obs["elm"][n] in (REF, RE0)))
                    full += 1
                if nvr == max_nobs:
                    want[10] = hl_vr * np.sqrt(max(_ndist(cfg, obs, n, ri, rj, rz) for n in ids if obs["elm"][n] == VR))
            else:                 # no limit: the reference's counts are cumulative over the merged types (:1444, :1473-1475)
                want[4] = nref + (nref + nre0) + nvr
                want[5:8] = [nref, nref + nre0, nvr]
                want[8:11] = [hl_ref * dzf, 0.0, hl_vr * dzf]
            assert np.array_equal(got[:8], want[:8]), (ij, il)
            if hits[ij, il] == 0:      # (exact hit: the reference reports the last SCANNED observation instead)
                assert np.allclose(got[8:], want[8:], rtol=1e-13, atol=0.0), (ij, il, got[8:], want[8:])
            checked += 1
    assert checked > 60 and (max_nobs == 0 or full > 5)


@pytest.mark.parametrize("criterion", [2, 3])
def test_oracle_nobs_out_criterion_2_3_matches_bruteforce(oracle, criterion):
    """MAX_NOBS_PER_GRID_CRITERION 2 (largest localisation weights) / 3 (smallest localised error variances): counts and the
    cut-off value (rloc / rdiag of the worst selected observation) against a brute-force numpy selection over ALL observations"""
    max_nobs = 10
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=6, nlon=24, nlat=24, nlev=5, max_nobs=max_nobs, seed=905, radius=5.0e3)
    cfg.MAX_NOBS_PER_GRID_CRITERION = criterion
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    pmean = np.asfortranarray(gues[:, :, cfg.MEMBER, cfg.iv3d_p - 1])
    out, hits = o.nobs_out(4, pmean)
    nij1, nlev = hgt1.shape
    dzf, dzf2 = cfg.dist_zero_fac, cfg.dist_zero_fac_square
    checked = full = 0
    for il in range(nlev):
        for ij in range(0, nij1, 11):
            ri, rj, rz = rig1[ij], rjg1[ij], hgt1[ij, il]
            got = out[ij, il, :]
            if _beta(cfg, ri, rj, rz, True) == 0.0:
                assert np.array_equal(got, np.zeros(11))
                continue
            want = np.zeros(11)
            for grp, slot in (((REF, RE0), 0), ((VR,), 2)):
                keys = []
                for n in np.nonzero(np.isin(obs["elm"], grp) & (obs["typ"] == 22))[0]:
                    e = obs["elm"][n]
                    hl = cfg.HORI_LOCAL_RADAR_OBSNOREF if e == RE0 else cfg.HORI_LOCAL_RADAR_VR if e == VR else cfg.HORI_LOCAL[21]
                    vl = cfg.VERT_LOCAL_RADAR_VR if e == VR else cfg.VERT_LOCAL[21]
                    nd_v = abs(obs["lev"][n] - rz) / vl
                    nd_h = np.sqrt(((ri - obs["ri"][n]) * cfg.DX) ** 2 + ((rj - obs["rj"][n]) * cfg.DY) ** 2) / hl
                    nd = nd_h * nd_h + nd_v * nd_v
                    if nd_v > dzf or nd_h > dzf or nd > dzf2:
                        continue
                    rloc = np.exp(-0.5 * nd)
                    keys.append(-rloc if criterion == 2 else obs["err"][n] ** 2 / rloc)     # ascending = best first
                keys.sort()
                cnt = min(len(keys), max_nobs)
                want[4] += cnt
                want[5 + slot] = cnt
                if cnt == max_nobs:
                    want[8 + slot] = -keys[cnt - 1] if criterion == 2 else keys[cnt - 1]
                    full += 1
            assert np.array_equal(got[:8], want[:8]), (ij, il, got, want)
            if hits[ij, il] == 0:
                assert np.allclose(got[8:], want[8:], rtol=1e-13, atol=0.0), (ij, il, got[8:], want[8:])
            checked += 1
    assert checked > 30 and full > 5


def test_oracle_nobs_out_sonde_types(oracle):
    """conventional report types: out(:,:,1) counts ADPUPA (type 1), nothing at the radar slots"""
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=7, nlon=10, nlat=9, nlev=4)
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    pmean = np.asfortranarray(gues[:, :, cfg.MEMBER, cfg.iv3d_p - 1])
    out, _ = o.nobs_out(4, pmean)
    r = o.das_letkf(gues.copy(order="F"), want_nobsl=True)
    ntyp = {t: out[:, :, i] for i, t in enumerate((1, 3, 4, 8, 22))}
    others = r["nobsl"] - sum(ntyp.values())
    assert (others >= 0).all() and np.array_equal(out[:, :, 5:], np.zeros_like(out[:, :, 5:]))
    assert set(np.unique(obs["typ"])) <= {1, 3, 4, 8, 22} and (others == 0).all() or (others > 0).any()
    assert ntyp[1].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("max_nobs,criterion", [(12, 1), (0, 1), (10, 2), (10, 3)])
def test_gpu_nobs_out_matches_oracle(oracle, space, max_nobs, criterion):
    import torch
    import scale_letkf_b200 as sl
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=8, nlon=32, nlat=32, nlev=6, max_nobs=max_nobs, seed=77, radius=6.0e3)
    cfg.MAX_NOBS_PER_GRID_CRITERION = criterion
    k = cfg.MEMBER
    pmean = np.asfortranarray(gues[:, :, k, cfg.iv3d_p - 1])
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    want, hits = o.nobs_out(4, pmean)
    eng = sl.LETKF(cfg, device=0)
    eng.set_letkf_obs(obs)
    eng.set_common_mpi_grid(rig1, rjg1, hgt1)
    if space == "host":
        got = eng.nobs_out(4, pmean)
    else:
        d_p = torch.from_numpy(np.ascontiguousarray(pmean.T)).cuda()
        d_lp = torch.from_numpy(np.ascontiguousarray(np.log(pmean).T)).cuda()
        got = eng.nobs_out(4, d_p, logp=d_lp).cpu().numpy().T
    eng.close()
    assert np.array_equal(got[:, :, :8], want[:, :, :8])                   # counts: bit-exact
    ok = hits == 0
    assert ok.mean() > 0.8
    if criterion == 1:
        assert np.array_equal(got[:, :, 8:][ok], want[:, :, 8:][ok])        # cut-off distance: same arithmetic, bit-exact
    else:                                                                   # rloc / rdiag carry exp(): CUDA vs libm, <= 1 ulp
        assert np.allclose(got[:, :, 8:][ok], want[:, :, 8:][ok], rtol=1e-14, atol=0.0)
    # at exact hits the reference reports the last scanned observation: never farther / worse than the worst selected one
    if criterion in (1, 3):
        assert (got[:, :, 8:][~ok] >= want[:, :, 8:][~ok] * (1 - 1e-14)).all()
    else:
        assert (got[:, :, 8:][~ok] <= want[:, :, 8:][~ok] * (1 + 1e-14)).all()
    assert (want[:, :, 5] > 0).any()
