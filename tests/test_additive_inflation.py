"""Additive inflation block of das_letkf (scale/letkf/letkf_tools.f90:804-929): the oracle's restatement against an
independent numpy evaluation (CPU), and letkf_b200_additive_inflation against the oracle (GPU, device and host buffers)."""
import numpy as np
import pytest

from helpers import radar_case, relerr, sonde_case


def _addi(gues, seed):
    rng = np.random.default_rng(seed)
    a = np.asfortranarray(gues + 0.1 * np.abs(gues).mean(axis=(0, 1, 2), keepdims=True) * rng.standard_normal(gues.shape))
    return a


def _numpy_additive(cfg, addi, anal, infl_add, gues=None, q_ratio=False, w=None, ishuf=None):
    """independent of oracle/: whole-array numpy, (nij, nlev, nens, nv3d)"""
    k = cfg.MEMBER
    mean = addi[:, :, 0, :].copy()
    for m in range(1, k):
        mean = mean + addi[:, :, m, :]          # member order, like ensmean_grd
    mean = mean / k
    out = anal.copy(order="F")
    wi = np.ones(addi.shape[0]) if w is None else w
    for m in range(k):
        ms = m if ishuf is None else ishuf[m] - 1
        t = (addi[:, :, ms, :] - mean) * infl_add * wi[:, None, None]
        if q_ratio:
            q = slice(cfg.iv3d_q - 1, cfg.iv3d_qg)
            t[:, :, q] = t[:, :, q] * gues[:, :, k, q]
        out[:, :, m, :] = anal[:, :, m, :] + t
    return out


@pytest.mark.parametrize("q_ratio,shuffle", [(False, False), (True, False), (True, True)])
def test_oracle_additive_inflation_matches_numpy(oracle, q_ratio, shuffle):
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=7, nlon=10, nlat=9, nlev=4)
    k = cfg.MEMBER
    addi, anal = _addi(gues, 3), _addi(gues, 4)
    ishuf = (np.random.default_rng(5).permutation(k) + 1).astype(np.int32) if shuffle else None
    want = _numpy_additive(cfg, addi, anal, 0.3, gues=gues, q_ratio=q_ratio, ishuf=ishuf)
    a2, an2 = addi.copy(order="F"), anal.copy(order="F")
    w = oracle.additive_inflation(cfg, a2, an2, 0.3, gues3d=gues, q_ratio=q_ratio, ishuf=ishuf)
    assert np.array_equal(w, np.ones(len(rig1)))
    assert np.array_equal(an2[:, :, :k, :], want[:, :, :k, :])          # same operations in the same order: bit-identical
    assert np.array_equal(an2[:, :, k:, :], anal[:, :, k:, :])          # mean / det slots untouched
    # the additive ensemble became perturbations around its own mean (letkf_tools.f90:869-877)
    assert np.abs(a2[:, :, :k, :].sum(axis=2)).max() <= 1e-9 * np.abs(addi).max()


def test_oracle_addinfl_weight_ref_only(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=6, nlon=24, nlat=24, nlev=4, radius=3.0e3)
    ref = (obs["elm"] == 4001)
    hloc = 400.0
    addi, anal = _addi(gues, 6), _addi(gues, 7)
    w = oracle.additive_inflation(cfg, addi.copy(order="F"), anal.copy(order="F"), 0.2, ref_only=True, rig1=rig1, rjg1=rjg1,
                                  ref_ri=obs["ri"][ref], ref_rj=obs["rj"][ref], hloc=hloc)
    d2 = ((rig1[:, None] - obs["ri"][ref][None, :]) * cfg.DX) ** 2 + ((rjg1[:, None] - obs["rj"][ref][None, :]) * cfg.DY) ** 2
    nd = d2.min(axis=1) / hloc ** 2
    want = np.where(nd <= cfg.dist_zero_fac_square, np.exp(-0.5 * nd), 0.0)
    assert relerr(w, want) <= 1e-15
    assert (w == 0).any() and (w > 0.5).any()      # the case covers both sides of the cutoff


@pytest.mark.gpu
@pytest.mark.parametrize("space", ["device", "host"])
@pytest.mark.parametrize("q_ratio,shuffle,ref_only", [(False, False, False), (True, True, False), (True, False, True)])
def test_gpu_additive_inflation_matches_oracle(oracle, space, q_ratio, shuffle, ref_only):
    import torch
    import scale_letkf_b200 as sl
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=9, nlon=24, nlat=24, nlev=5, radius=3.0e3)
    k = cfg.MEMBER
    addi, anal = _addi(gues, 8), _addi(gues, 9)
    ishuf = (np.random.default_rng(10).permutation(k) + 1).astype(np.int32) if shuffle else None
    eng = sl.LETKF(cfg, device=0)
    eng.set_letkf_obs(obs)
    eng.set_common_mpi_grid(rig1, rjg1, hgt1)
    # the oracle takes the (REF, PHARAD) observations explicitly; their localisation scale comes from the combined-type table
    ref = (obs["elm"] == 4001) & (obs["typ"] == 22)
    hloc = None
    for ic in range(eng.obs_info()[1]):
        ct = eng.ctype(ic)
        if ct.elm == 4001 and ct.typ == 22:
            hloc = ct.hori_loc
    assert hloc is not None and ref.any()
    want_a, want_an = addi.copy(order="F"), anal.copy(order="F")
    w_ref = oracle.additive_inflation(cfg, want_a, want_an, 0.25, gues3d=gues, q_ratio=q_ratio, ref_only=ref_only, ishuf=ishuf,
                                      rig1=rig1, rjg1=rjg1, ref_ri=obs["ri"][ref], ref_rj=obs["rj"][ref], hloc=hloc)
    if space == "device":
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda()     # Fortran (nij,nlev,nens,nv) memory order
        d_addi, d_anal, d_gues = T(addi), T(anal), T(gues)
        w = eng.additive_inflation(d_addi, d_anal, 0.25, gues3d=d_gues, q_ratio=q_ratio, ref_only=ref_only, ishuf=ishuf, want_weight=True)
        got, w = d_anal.cpu().numpy().T, w.cpu().numpy()
        assert torch.equal(d_addi, T(addi))            # the additive ensemble is read only
    else:
        got = anal.copy(order="F")
        w = eng.additive_inflation(addi.copy(order="F"), got, 0.25, gues3d=gues, q_ratio=q_ratio, ref_only=ref_only, ishuf=ishuf,
                                   want_weight=True)
    eng.close()
    assert relerr(w, w_ref) <= 1e-14                   # exp() of CUDA against libm
    if ref_only:
        assert relerr(got[:, :, :k, :], want_an[:, :, :k, :], axis=(0, 1, 2)) <= 1e-13
    else:
        assert np.array_equal(got[:, :, :k, :], want_an[:, :, :k, :])      # no transcendental involved: bit-identical
    assert np.array_equal(got[:, :, k:, :], anal[:, :, k:, :])
