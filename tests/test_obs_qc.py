"""Departure + QC half of set_letkf_obs (scale/letkf/letkf_obs.f90:355-560).

PARITY UNPINNED BY THE REFERENCE (no fixtures in gylien/scale-letkf, Fortran not buildable here): the
C++ oracle is pinned against an INDEPENDENT numpy restatement written from the reference's text (and the
fixture tests/golden/obsqc.npz frozen from it); the CUDA path must match the oracle bit for bit
(integer QC codes, IEEE add / divide / subtract in the reference's order)."""
import ctypes as C
import os

import numpy as np
import pytest

import scale_letkf_b200 as sl
from scale_letkf_b200 import capi, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obsqc.npz")
CASE = dict(member=12, nobs=600, det=True, seed_no=7)


def qc_defaults(**kw):
    q = capi.QcConfig()
    q.GROSS_ERROR = 5.0
    for f in ("GROSS_ERROR_RAIN", "GROSS_ERROR_RADAR_REF", "GROSS_ERROR_RADAR_VR", "GROSS_ERROR_RADAR_PRH",
              "GROSS_ERROR_TCX", "GROSS_ERROR_TCY", "GROSS_ERROR_TCP"):
        setattr(q, f, -1.0)
    q.RADAR_REF_THRES_DBZ = 15.0
    q.USE_RADAR_REF = q.USE_RADAR_VR = 1
    q.MIN_RADAR_REF_MEMBER = q.MIN_RADAR_REF_MEMBER_OBSREF = 1
    for k, v in kw.items():
        setattr(q, k, v)
    return q


def numpy_departure_qc(q, member, det, elm, dat, err, qc, ensval):
    """Vectorised restatement of letkf_obs.f90:362-549 (independent of oracle/)."""
    qc, ens, val = qc.copy(), ensval.copy(), np.zeros(len(elm))
    act = qc <= 0
    ref = (elm == 4001) | (elm == 4004)
    thr = q.RADAR_REF_THRES_DBZ + 1.0e-6
    if not q.USE_RADAR_REF:
        qc[act & ref] = 90
    act = qc <= 0
    bad = act & ref & (dat == -9.99e33)
    qc[bad] = 50
    act = qc <= 0
    mem_ref = (ens[:, :member] > thr).sum(axis=1)
    need = np.where(dat > thr, q.MIN_RADAR_REF_MEMBER_OBSREF, q.MIN_RADAR_REF_MEMBER)
    qc[act & ref & (mem_ref < need)] = 12
    act = qc <= 0
    if not q.USE_RADAR_VR:
        qc[act & (elm == 4002)] = 90
    act = qc <= 0
    s = ens[:, 0].copy()
    for i in range(1, member):          # sequential member order, like the reference
        s = s + ens[:, i]
    mean = s / float(member)
    ens[act, :member] = ens[act, :member] - mean[act, None]
    val[act] = dat[act] - mean[act]
    if det:
        ens[act, member] = dat[act] - ens[act, member]
    ge = lambda v: q.GROSS_ERROR if v < 0 else v
    fac = np.full(len(elm), q.GROSS_ERROR)
    fac[elm == 19999] = ge(q.GROSS_ERROR_RAIN)
    fac[ref] = ge(q.GROSS_ERROR_RADAR_REF)
    fac[elm == 4002] = ge(q.GROSS_ERROR_RADAR_VR)
    fac[elm == 4003] = ge(q.GROSS_ERROR_RADAR_PRH)
    fac[elm == 99991] = ge(q.GROSS_ERROR_TCX)
    fac[elm == 99992] = ge(q.GROSS_ERROR_TCY)
    fac[elm == 99993] = ge(q.GROSS_ERROR_TCP)
    qc[act & (np.abs(val) > fac * err)] = 5
    return qc, val, ens


VARIANTS = [dict(), dict(USE_RADAR_REF=0), dict(USE_RADAR_VR=0), dict(MIN_RADAR_REF_MEMBER=5, MIN_RADAR_REF_MEMBER_OBSREF=9),
            dict(GROSS_ERROR=3.0, GROSS_ERROR_RADAR_REF=1.5, GROSS_ERROR_RAIN=8.0, GROSS_ERROR_TCP=0.5)]


@pytest.mark.parametrize("kw", VARIANTS)
def test_oracle_matches_independent_numpy(oracle, kw):
    o = synth.make_raw_obs(**CASE)
    q = qc_defaults(**kw)
    a = oracle.obs_departure_qc(q, CASE["member"], CASE["det"], o["elm"], o["dat"], o["err"], o["qc"], o["ensval"])
    b = numpy_departure_qc(q, CASE["member"], CASE["det"], o["elm"], o["dat"], o["err"], o["qc"], o["ensval"])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    if not kw:
        assert set(np.unique(a[0])) >= {0, 5, 12, 50}   # every rule of the default configuration fires
        assert np.abs(a[2][a[0] == 0, :CASE["member"]].sum(axis=1)).max() < 1e-9   # perturbations are zero-mean


def test_oracle_matches_golden(oracle):
    g = np.load(GOLD)
    o = synth.make_raw_obs(**CASE)
    qc, val, ens = oracle.obs_departure_qc(qc_defaults(), CASE["member"], CASE["det"], o["elm"], o["dat"], o["err"],
                                           o["qc"], o["ensval"])
    assert np.array_equal(qc, g["qc"]) and np.array_equal(val, g["val"]) and np.array_equal(ens, g["ensval"])


def test_qc_config_abi(oracle):
    lib = capi.load_library()
    assert lib.letkf_b200_abi_size_qc() == C.sizeof(capi.QcConfig)
    q = capi.QcConfig()
    lib.letkf_b200_qc_config_defaults(C.byref(q))
    d = qc_defaults()
    for f, _ in capi.QcConfig._fields_:
        assert getattr(q, f) == getattr(d, f), f


@pytest.mark.gpu
@pytest.mark.parametrize("kw", VARIANTS)
@pytest.mark.parametrize("member,det,nobs", [(12, True, 600), (50, False, 5000), (100, True, 3000), (1000, False, 400)])
def test_cuda_departure_qc_bitexact(oracle, kw, member, det, nobs):
    o = synth.make_raw_obs(member=member, nobs=nobs, det=det, seed_no=70 + member)
    cfg = sl.default_config(MEMBER=member, nlon=8, nlat=8, nlev=2)
    cfg.DET_RUN = 1 if det else 0
    e = sl.LETKF(sl.resolve_config(cfg), device=0)
    q = qc_defaults(**kw)
    got = e.obs_departure_qc(o["elm"], o["dat"], o["err"], o["qc"], o["ensval"], q)
    ref = oracle.obs_departure_qc(q, member, det, o["elm"], o["dat"], o["err"], o["qc"], o["ensval"])
    for x, y in zip(got, ref):
        assert np.array_equal(x, y)
    e.close()


@pytest.mark.gpu
def test_cuda_departure_qc_golden():
    g = np.load(GOLD)
    o = synth.make_raw_obs(**CASE)
    cfg = sl.default_config(MEMBER=CASE["member"], nlon=8, nlat=8, nlev=2)
    cfg.DET_RUN = 1
    e = sl.LETKF(sl.resolve_config(cfg), device=0)
    qc, val, ens = e.obs_departure_qc(o["elm"], o["dat"], o["err"], o["qc"], o["ensval"])
    assert np.array_equal(qc, g["qc"]) and np.array_equal(val, g["val"]) and np.array_equal(ens, g["ensval"])
    e.close()


# ---- monit_dep (scale/common/common_obs_scale.f90:1851-1895) -----------------
UIDS = [2819, 2820, 3073, 3074, 3330, 3331, 14593, 19999, 4001, 4004, 4002, 4003, 8800, 99991, 99992, 99993]


def numpy_monit_dep(elm, dep, qc):
    e = np.where(elm == 3074, 3073, elm)
    e = np.where(e == 4004, 4001, e)
    n, b, r = np.zeros(16, dtype=np.int32), np.full(16, -9.99e33), np.full(16, -9.99e33)
    for i, u in enumerate(UIDS):
        m = (e == u) & (qc == 0)
        n[i] = m.sum()
        if n[i]:
            b[i] = dep[m].mean()
            r[i] = np.sqrt((dep[m] ** 2).mean())
    return n, b, r


def _monit_case():
    o = synth.make_raw_obs(member=12, nobs=20000, det=False, seed_no=17)
    q = qc_defaults()
    return o, q


def test_oracle_monit_dep_vs_numpy(oracle):
    o, q = _monit_case()
    qc, val, _ = oracle.obs_departure_qc(q, 12, False, o["elm"], o["dat"], o["err"], o["qc"], o["ensval"])
    n, b, r = oracle.monit_dep(o["elm"], val, qc)
    n2, b2, r2 = numpy_monit_dep(o["elm"], val, qc)
    assert np.array_equal(n, n2) and n[3] == 0 and n[9] == 0 and b[3] == -9.99e33   # Tv / RE0 folded into T / REF
    ok = n > 0
    assert np.abs(b[ok] - b2[ok]).max() <= 1e-12 * np.abs(b2[ok]).max() + 1e-13
    assert np.abs(r[ok] - r2[ok]).max() <= 1e-12 * np.abs(r2[ok]).max()


@pytest.mark.gpu
def test_cuda_monit_dep(oracle):
    o, q = _monit_case()
    qc, val, _ = oracle.obs_departure_qc(q, 12, False, o["elm"], o["dat"], o["err"], o["qc"], o["ensval"])
    e = sl.LETKF(sl.resolve_config(sl.default_config(MEMBER=12, nlon=8, nlat=8, nlev=2)), device=0)
    n, b, r = e.monit_dep(o["elm"], val, qc)
    n2, b2, r2 = oracle.monit_dep(o["elm"], val, qc)
    assert np.array_equal(n, n2)
    ok = n2 > 0
    assert np.array_equal(b[~ok], b2[~ok]) and np.array_equal(r[~ok], r2[~ok])      # undef where empty
    assert np.abs(b[ok] - b2[ok]).max() <= 1e-12 * np.abs(b2[ok]).max() + 1e-13     # tree vs serial summation
    assert np.abs(r[ok] - r2[ok]).max() <= 1e-12 * np.abs(r2[ok]).max()
    n3, b3, r3 = e.monit_dep(o["elm"], val, qc)
    assert np.array_equal(b, b3) and np.array_equal(r, r3)                           # deterministic
    e.close()


@pytest.mark.gpu
def test_set_letkf_obs_raw_pipeline(oracle):
    """set_letkf_obs as a whole: H(x_m) of every member -> device QC / departures -> bucket sort -> das_letkf;
    must equal the oracle fed with the oracle's own QC-passed departures."""
    from helpers import sonde_case, host_logp, relerr, TOL
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=12, nsonde=30, nsfc=100, det=True)
    k = cfg.MEMBER
    g = synth.rng(123)
    # raw H(x_m) = perturbation + (y - departure): the mean of H(x) reproduces the case's departures, the data keep
    # their physical values (PS observations enter the vertical localisation through log(dat)); 5 % of the rows get
    # an ensemble far from the data -> gross error
    dat = obs["dat"].copy()
    ens_raw = obs["ensval"].copy()
    ens_raw[:, :k] += (dat - obs["val"])[:, None]
    ens_raw[:, k] = dat - obs["ensval"][:, k]          # H(x_det) such that y - H(x_det) = the case's depd
    out_mask = g.uniform(size=len(dat)) < 0.05
    ens_raw[out_mask, :k] += (40.0 * obs["err"])[out_mask, None]
    raw = dict(obs, dat=dat, ensval=ens_raw)
    q = qc_defaults()
    e = sl.LETKF(cfg, device=0)
    info = e.set_letkf_obs_raw(raw, q)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    # reference chain on the CPU
    qc, val, ens = oracle.obs_departure_qc(q, k, True, raw["elm"], raw["dat"], raw["err"], np.zeros(len(dat), np.int32), ens_raw)
    assert np.array_equal(info["qc"], qc) and np.array_equal(info["val"], val)
    assert info["kept"] == int((qc == 0).sum()) and 0 < info["kept"] < len(dat)
    keep = qc == 0
    ref_obs = {kf: np.ascontiguousarray(np.asarray(raw[kf])[keep]) for kf in ("elm", "typ", "ri", "rj", "lev", "dat", "err")}
    ref_obs["val"], ref_obs["ensval"] = np.ascontiguousarray(val[keep]), np.ascontiguousarray(ens[keep])
    o = oracle.Oracle(cfg)
    o.set_obs(ref_obs)
    o.set_grid(rig1, rjg1, hgt1)
    ref = o.das_letkf(gues.copy(order="F"), want_nobsl=True)
    got = e.das_letkf(gues.copy(order="F"), want_nobsl=True, logp=host_logp(cfg, gues))
    assert np.array_equal(got["nobsl"], ref["nobsl"])
    slots = list(range(k)) + [k + 1]
    assert relerr(got["anal3d"][:, :, slots, :], ref["anal3d"][:, :, slots, :], axis=(0, 1, 2)) <= TOL
    n2, b2, r2 = oracle.monit_dep(raw["elm"], val, qc)
    assert np.array_equal(info["nobs"], n2)
    e.close()


@pytest.mark.gpu
def test_device_obs_chain_matches_host_chain(oracle):
    """letkf_b200_obs_departure_qc(DEVICE) -> letkf_b200_set_obs_device (qc filter, combined types, vertical coordinate
    and bucket sort on the device) gives the tables of the host chain (host filter + letkf_b200_set_obs), and the
    same analysis."""
    import torch
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    cfg = synth.config_c3(nlon=24, nlat=24, nlev=6, member=12, max_nobs=40)
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=71)
    pos = synth.make_radar_obs(cfg, radius_m=5.0e3, zmin=500.0, zmax=6000.0, dz=1000.0, seed_no=72)
    n = len(pos["elm"])
    raw = synth.make_raw_obs(member=12, nobs=n, seed_no=73)
    radar = np.isin(pos["elm"], (4001, 4004))
    raw["elm"] = pos["elm"]                      # radar elements at radar positions, H(x_m) / data of the raw generator
    raw["ensval"] = np.where(radar[:, None], np.abs(raw["ensval"]) + 8.0, raw["ensval"])
    raw["dat"] = np.where(radar, np.abs(raw["dat"]) + 8.0, raw["dat"])
    raw.update({kf: pos[kf] for kf in ("typ", "ri", "rj", "lev")})
    e1 = sl.LETKF(cfg, device=0)
    r1 = e1.set_letkf_obs_raw(raw, qc_in=raw["qc"].astype(np.int32))
    e1.set_common_mpi_grid(rig1, rjg1, hgt1)
    e2 = sl.LETKF(cfg, device=0)
    dev = torch.device("cuda", 0)
    t = {kf: torch.as_tensor(np.ascontiguousarray(raw[kf]), device=dev) for kf in ("elm", "typ", "ri", "rj", "lev", "dat", "err")}
    t["ensval"] = torch.as_tensor(np.ascontiguousarray(raw["ensval"]), device=dev)
    qc = torch.as_tensor(raw["qc"].astype(np.int32), device=dev)
    t["val"] = e2.obs_departure_qc_device(t["elm"], t["dat"], t["err"], qc, t["ensval"])
    nk = e2.set_letkf_obs_device(t, qc)
    e2.set_common_mpi_grid(rig1, rjg1, hgt1)
    assert nk == r1["kept"] and 0 < nk < n
    assert np.array_equal(qc.cpu().numpy(), r1["qc"])
    assert e1.obs_info() == e2.obs_info()
    for ic in range(e1.obs_info()[1]):
        assert np.array_equal(e1.ac_ext(ic), e2.ac_ext(ic))
    kept = e2.kept_index()
    assert np.array_equal(kept, np.nonzero(r1["qc"] == 0)[0])
    assert np.array_equal(e1.sorted_index(), e2.sorted_index())
    a = e1.das_letkf(gues.copy(order="F"), want_nobsl=True)
    b = e2.das_letkf(gues.copy(order="F"), want_nobsl=True)
    assert np.array_equal(a["nobsl"], b["nobsl"]) and a["nsolved"] > 0
    assert np.array_equal(a["anal3d"][:, :, :12, :], b["anal3d"][:, :, :12, :])
    e1.close()
    e2.close()
