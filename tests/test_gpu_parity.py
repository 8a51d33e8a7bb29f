"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs.  Selection sets / counts / bucket tables bit-exact; weights and analysis within
1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest

import scale_letkf_b200 as sl
from scale_letkf_b200 import synth
from helpers import TOL, relerr, sonde_case, radar_case, host_logp, sample_points

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def core_engine():
    cfg = sl.default_config(MEMBER=20, nlon=8, nlat=8, nlev=2)
    e = sl.LETKF(sl.resolve_config(cfg), device=0)
    yield e
    e.close()


def _cmp_core(r, ref, keys=("trans", "transm", "pao", "transmd")):
    for key in keys:
        if ref.get(key) is None:
            continue
        assert relerr(r[key], ref[key]) <= TOL, key


def test_core_batch_c1(core_engine, oracle):
    """BASELINE config 1: k=20, 10^4 points, p <= 100 (includes p = 0 and p = 1)."""
    c = synth.make_core_batch(ne=20, npts=10000, nobs=100, seed_no=1, det=True)
    args = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"], c["parm_infl"])
    ref = oracle.core_batch(*args, depd=c["depd"])
    r = core_engine.letkf_core(*args, depd=c["depd"])
    _cmp_core(r, ref)
    # p == 0 known answer is exact
    assert np.array_equal(r["trans"][0], np.eye(20)) and np.array_equal(r["transm"][0], np.zeros(20))
    # symmetric outputs, invariants at full size
    assert np.abs(r["trans"] - r["trans"].transpose(0, 2, 1)).max() == 0.0
    w = r["trans"][2:200]
    pa = r["pao"][2:200]
    assert np.abs(w @ w - 19 * pa).max() / np.abs(pa).max() < 1e-10


@pytest.mark.parametrize("ne,nobs,npts", [(7, 30, 64), (33, 60, 64), (50, 300, 48), (64, 100, 32), (100, 400, 16),
                                          (101, 150, 8), (128, 200, 8), (2, 5, 16)])
def test_core_batch_sizes(core_engine, oracle, ne, nobs, npts):
    c = synth.make_core_batch(ne=ne, npts=npts, nobs=nobs, seed_no=100 + ne, det=True, infl=1.07)
    args = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"], c["parm_infl"])
    ref = oracle.core_batch(*args, depd=c["depd"])
    r = core_engine.letkf_core(*args, depd=c["depd"])
    _cmp_core(r, ref)


def test_core_batch_options(core_engine, oracle):
    c = synth.make_core_batch(ne=20, npts=200, nobs=50, seed_no=3, infl=1.1)
    base = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"])
    # rdiag_wloc = False: rloc applied inside; adaptive inflation estimate; transm absent
    err2 = c["rdiag"] * c["rloc"]
    ref = oracle.core_batch(*base, err2, c["rloc"], c["dep"], c["parm_infl"], rdiag_wloc=False, infl_update=True)
    r = core_engine.letkf_core(*base, err2, c["rloc"], c["dep"], c["parm_infl"], rdiag_wloc=False, infl_update=True)
    _cmp_core(r, ref)
    sel = c["nobsl"] > 0
    assert relerr(r["parm_infl"][sel], ref["parm_infl"][sel]) <= TOL
    assert np.array_equal(r["parm_infl"][~sel], ref["parm_infl"][~sel])
    ref = oracle.core_batch(*base, c["rdiag"], c["rloc"], c["dep"], c["parm_infl"], want_transm=False, want_pao=False)
    r = core_engine.letkf_core(*base, c["rdiag"], c["rloc"], c["dep"], c["parm_infl"], want_transm=False, want_pao=False)
    _cmp_core(r, ref, keys=("trans",))


def _engines(cfg, rig1, rjg1, hgt1, obs, oracle):
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    e = sl.LETKF(cfg, device=0)
    e.set_letkf_obs(obs)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    return o, e


@pytest.mark.parametrize("case", ["sonde", "radar", "empty"])
def test_bucket_tables_bitexact(oracle, case):
    if case == "sonde":
        cfg, rig1, rjg1, hgt1, obs, _ = sonde_case(nsonde=40, nsfc=200)
    elif case == "radar":
        cfg, rig1, rjg1, hgt1, obs, _ = radar_case()
    else:
        cfg, rig1, rjg1, hgt1, obs, _ = sonde_case(nsonde=0, nsfc=0)
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    assert e.obs_info() == o.obs_info()
    assert np.array_equal(e.sorted_index(), o.sorted_index())
    for ic in range(o.obs_info()[1]):
        a, b = e.ctype(ic), o.ctype(ic)
        for f, _ in a._fields_:
            assert getattr(a, f) == getattr(b, f), f
        assert np.array_equal(e.ac_ext(ic), o.ac_ext(ic))
    e.close()


def test_search_bitexact_nolimit(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(nsonde=40, nsfc=200)
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    pts = sample_points(cfg, rig1, rjg1, hgt1, gues, stride=3)
    n1, i1, d1, l1 = o.obs_local(*pts, 1, 4096)
    n2, i2, d2, l2 = e.obs_local(*pts, 1, 4096)
    assert n1.max() > 50
    assert np.array_equal(n1, n2)
    assert np.array_equal(i1, i2)   # same order as the reference's bucket scan
    m = i1 >= 0
    assert relerr(d2[m], d1[m]) < 1e-14 and relerr(l2[m], l1[m]) < 1e-14
    e.close()


@pytest.mark.parametrize("criterion", [1, 2, 3])
def test_search_bitexact_limited(oracle, criterion):
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(max_nobs=30)
    cfg.MAX_NOBS_PER_GRID_CRITERION = criterion
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    pts = sample_points(cfg, rig1, rjg1, hgt1, gues, stride=5)
    n1, i1, d1, l1 = o.obs_local(*pts, 1, 4096)
    n2, i2, d2, l2 = e.obs_local(*pts, 1, 4096)
    assert n1.max() == 60 and (n1 == 0).any()
    assert np.array_equal(n1, n2)
    for p in range(len(n1)):
        a, b = np.argsort(i1[p, :n1[p]]), np.argsort(i2[p, :n2[p]])
        assert np.array_equal(i1[p, :n1[p]][a], i2[p, :n2[p]][b])
        assert np.allclose(d1[p, :n1[p]][a], d2[p, :n2[p]][b], rtol=1e-14, atol=0)
    e.close()


@pytest.mark.parametrize("criterion", [1, 3])
def test_search_limited_rescan_fallback(oracle, criterion, monkeypatch):
    """candidate buffer too small (forced): the storage-free re-scan select gives the same sets"""
    monkeypatch.setenv("LETKF_B200_CAND_CAP", "8")
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(max_nobs=30)
    cfg.MAX_NOBS_PER_GRID_CRITERION = criterion
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    pts = sample_points(cfg, rig1, rjg1, hgt1, gues, stride=5)
    n1, i1, d1, l1 = o.obs_local(*pts, 1, 4096)
    n2, i2, d2, l2 = e.obs_local(*pts, 1, 4096)
    assert np.array_equal(n1, n2)
    for p in range(len(n1)):
        assert np.array_equal(np.sort(i1[p, :n1[p]]), np.sort(i2[p, :n2[p]]))
    e.close()


def test_search_limited_ties_in_scan_order(oracle):
    """duplicated observations (exactly equal keys): the N-th place is shared; ties go to scan order,
    so the selected multiset of keys must still equal the oracle's N smallest"""
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(max_nobs=30)
    dup = {k: np.concatenate([v, v]) for k, v in obs.items()}
    o, e = _engines(cfg, rig1, rjg1, hgt1, dup, oracle)
    pts = sample_points(cfg, rig1, rjg1, hgt1, gues, stride=9)
    n1, i1, d1, l1 = o.obs_local(*pts, 1, 4096)
    n2, i2, d2, l2 = e.obs_local(*pts, 1, 4096)
    assert np.array_equal(n1, n2)
    for p in range(len(n1)):
        assert np.allclose(np.sort(d1[p, :n1[p]]), np.sort(d2[p, :n2[p]]), rtol=1e-14, atol=0)
    e.close()


def _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle, infl3d=False, check_rtps=True):
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    k = cfg.MEMBER
    nens = gues.shape[2]
    g1, g2 = gues.copy(order="F"), gues.copy(order="F")
    i1 = i2 = None
    if infl3d:
        i1 = np.full((gues.shape[0], gues.shape[1], gues.shape[3]), cfg.INFL_MUL, order="F")
        i2 = i1.copy(order="F")
    ref = o.das_letkf(g1, infl3d=i1, want_rtps=True, want_nobsl=True)
    out = e.das_letkf(g2, infl3d=i2, want_rtps=True, want_nobsl=True, logp=host_logp(cfg, gues))
    assert ref["status"] == 0 and out["status"] == 0
    assert np.array_equal(out["nobsl"], ref["nobsl"])
    assert out["npoints"] == ref["npoints"] and out["nsolved"] == ref["nsolved"]
    slots = list(range(k)) + ([k + 1] if cfg.DET_RUN else [])
    a, b = out["anal3d"][:, :, slots, :], ref["anal3d"][:, :, slots, :]
    assert relerr(a, b, axis=(0, 1, 2)) <= TOL
    assert np.array_equal(g2[:, :, :k + 1, :], g1[:, :, :k + 1, :])   # gues destroyed identically
    if check_rtps:
        assert relerr(out["rtps"], ref["rtps"]) <= TOL
    if infl3d:
        assert relerr(i2, i1) <= TOL
    e.close()
    return out, ref


@pytest.mark.parametrize("relax", ["rtps", "rtpp", "none"])
def test_das_sonde(oracle, relax):
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=20, nsonde=30, nsfc=100)
    cfg.RELAX_ALPHA_SPREAD = 0.95 if relax == "rtps" else 0.0
    cfg.RELAX_ALPHA = 0.7 if relax == "rtpp" else 0.0
    cfg.BOUNDARY_BUFFER_WIDTH = 45.0e3
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0


def test_das_det_qtop_qsprd_inflated_prior(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=12, nsonde=30, nsfc=100, det=True)
    cfg.Q_UPDATE_TOP = 300.0e2
    cfg.Q_SPRD_MAX = 0.05
    cfg.INFL_MUL = 1.15
    cfg.RELAX_TO_INFLATED_PRIOR = 1
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)


def test_das_adaptive_inflation_and_min(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=10, nsonde=30, nsfc=100)
    cfg.INFL_MUL_ADAPTIVE = 1
    cfg.INFL_MUL = 1.05
    cfg.INFL_MUL_MIN = 1.02
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle, infl3d=True)


def test_das_variable_localisation_groups(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=10, nsonde=30, nsfc=100)
    for n in range(11):
        cfg.VAR_LOCAL[2][n] = 1.0 if n >= 5 else 0.0   # Q obs only update moisture variables
        cfg.VAR_LOCAL[3][n] = 0.5 if n == 4 else 1.0   # PS obs half weight on p
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)


@pytest.mark.parametrize("member,max_nobs", [(16, 30), (50, 100)])
def test_das_radar(oracle, member, max_nobs):
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=member, max_nobs=max_nobs, nlon=32, nlat=32, nlev=6, det=True)
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0 and out["nsolved"] < out["npoints"]


def test_das_k100(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=100, max_nobs=200, nlon=20, nlat=20, nlev=4, radius=4.0e3)
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)


def test_das_device_resident_matches_host(oracle):
    import torch
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=20, nsonde=30, nsfc=100)
    e = sl.LETKF(cfg, device=0)
    e.set_letkf_obs(obs)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    host = e.das_letkf(gues.copy(order="F"))
    # torch tensor with the memory order of the Fortran array: (nv3d, nens, nlev, nij1)
    t = torch.from_numpy(np.ascontiguousarray(gues.transpose(3, 2, 1, 0))).cuda()
    dev = e.das_letkf(t)
    a = dev["anal3d"].cpu().numpy().transpose(3, 2, 1, 0)
    k = cfg.MEMBER
    assert np.array_equal(a[:, :, :k, :], host["anal3d"][:, :, :k, :])
    e.close()


def test_ensmean_and_transposes(oracle):
    import ctypes as C
    import torch
    cfg = synth.config_c2(nlon=7, nlat=5, nlev=3, member=4)
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=5)
    e = sl.LETKF(cfg, device=0)
    g = gues.copy(order="F")
    g[:, :, 4, :] = 0.0
    r = g.copy(order="F")
    e.ensmean_grd(g)
    oracle.ensmean_grd(4, r)
    assert np.array_equal(g, r)
    np_, nens, nv3d, nlev = 4, 5, cfg.nv3d, cfg.nlev
    nlevall = nlev * nv3d
    L = oracle.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = synth.rng(66)
    for rank in range(np_):
        nij1, nmax = e.nij1_of(np_, rank)
        assert (nij1, nmax) == oracle.nij1(7, 5, np_, rank)
        v3dg = np.asfortranarray(rng.standard_normal((nlev, 7, 5, nv3d)))
        ref = np.zeros((nmax, nlevall, np_), order="F")
        L.oracle_grd_to_buf(7, 5, nlev, nv3d, 0, np_, P(v3dg), None, P(ref))
        d_in = torch.from_numpy(v3dg.ravel(order="K").copy()).cuda()
        d_buf = torch.zeros(ref.size, dtype=torch.float64, device="cuda")
        e.grd_to_buf(np_, d_in, None, d_buf)
        assert np.array_equal(d_buf.cpu().numpy(), ref.ravel(order="K"))
        d_back = torch.zeros_like(d_in)
        e.buf_to_grd(np_, d_buf, d_back, None)
        assert np.array_equal(d_back.cpu().numpy(), v3dg.ravel(order="K"))
        bufr = np.asfortranarray(rng.standard_normal((nmax, nlevall, np_)))
        v3d_ref = np.zeros((nij1, nlev, nens, nv3d), order="F")
        L.oracle_buf_to_ens(7, 5, nlev, nv3d, 0, np_, rank, nens, 1, np_, P(bufr), P(v3d_ref), None)
        d_bufr = torch.from_numpy(bufr.ravel(order="K").copy()).cuda()
        d_v3d = torch.zeros(v3d_ref.size, dtype=torch.float64, device="cuda")
        e.buf_to_ens(np_, rank, nens, 1, np_, d_bufr, d_v3d, None)
        assert np.array_equal(d_v3d.cpu().numpy(), v3d_ref.ravel(order="K"))
        d_bufs = torch.zeros_like(d_bufr)
        e.ens_to_buf(np_, rank, nens, 1, np_, d_v3d, None, d_bufs)
        got = d_bufs.cpu().numpy().reshape(bufr.shape, order="F")
        assert np.array_equal(got[:nij1], bufr[:nij1])
    e.close()


@pytest.mark.parametrize("pool_mb", ["0.0005", "0.02"])
def test_das_presearch_pool_overflow_redo(oracle, monkeypatch, pool_mb):
    """A pre-search pool that is far too small: the search-free solver hands the points whose lists did
    not fit to the redo pass (in-kernel search); results must not change."""
    monkeypatch.setenv("LETKF_B200_POOL_MB", pool_mb)
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=20, nsonde=30, nsfc=100)
    cfg.BOUNDARY_BUFFER_WIDTH = 45.0e3
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0


def test_das_presearch_off_matches_on(oracle, monkeypatch):
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=16, max_nobs=30, nlon=32, nlat=32, nlev=6, det=True)
    e = sl.LETKF(cfg, device=0)
    e.set_letkf_obs(obs)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    lp = host_logp(cfg, gues)
    on = e.das_letkf(gues.copy(order="F"), want_nobsl=True, logp=lp)
    monkeypatch.setenv("LETKF_B200_PRESEARCH", "0")
    off = e.das_letkf(gues.copy(order="F"), want_nobsl=True, logp=lp)
    assert np.array_equal(on["nobsl"], off["nobsl"])
    k = cfg.MEMBER
    assert relerr(on["anal3d"][:, :, :k, :], off["anal3d"][:, :, :k, :], axis=(0, 1, 2)) <= 1e-12
    e.close()


def test_das_no_observations(oracle):
    """nobstotal = 0: every point takes the nobsl == 0 branch (common_letkf.f90:89-107): W = sqrt(infl) I."""
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=8, nsonde=0, nsfc=0)
    cfg.INFL_MUL = 1.1
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] == 0 and out["npoints"] == gues.shape[0] * gues.shape[1]


@pytest.mark.parametrize("solver", ["ns", "tiled"])
def test_das_with_2d_variables(oracle, monkeypatch, solver):
    """nv2d > 0 (the reference carries nv2d = 0 today, common_nml.f90:20, but das_letkf analyses 2-D variables at
    ilev = 1, letkf_tools.f90:300-312): gues2d / anal2d through both solver paths."""
    if solver == "tiled":
        monkeypatch.setenv("LETKF_B200_SOLVER", "tiled")
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=10, nsonde=30, nsfc=100, det=True)
    cfg.nv2d = 2
    k, nens, nij1 = cfg.MEMBER, gues.shape[2], gues.shape[0]
    g = synth.rng(77)
    gues2d = np.asfortranarray(g.standard_normal((nij1, nens, 2)) * 3.0 + 10.0)
    m = gues2d[:, 0, :].copy()
    for mm in range(1, k):
        m += gues2d[:, mm, :]
    gues2d[:, k, :] = m / k
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    g1, g2 = gues.copy(order="F"), gues.copy(order="F")
    h1, h2 = gues2d.copy(order="F"), gues2d.copy(order="F")
    ref = o.das_letkf(g1, gues2d=h1)
    out = e.das_letkf(g2, gues2d=h2, logp=host_logp(cfg, gues))
    slots = list(range(k)) + [k + 1]
    assert relerr(out["anal3d"][:, :, slots, :], ref["anal3d"][:, :, slots, :], axis=(0, 1, 2)) <= TOL
    assert relerr(out["anal2d"][:, slots, :], ref["anal2d"][:, slots, :], axis=(0, 1)) <= TOL
    assert np.array_equal(h2[:, :k + 1, :], h1[:, :k + 1, :])
    e.close()


def test_set_obs_rejects_non_positive_pressure():
    """An observation localised in ln p with a non-positive pressure (PS: dat, others: lev) is refused instead of
    silently spreading NaN through every distance test."""
    cfg, rig1, rjg1, hgt1, obs, _ = sonde_case(nsonde=10, nsfc=20)
    bad = dict(obs)
    bad["lev"] = obs["lev"].copy()
    i = int(np.argmax(obs["elm"] != 14593))
    bad["lev"][i] = -5.0
    e = sl.LETKF(cfg, device=0)
    with pytest.raises(sl.api.LetkfError) as ei:
        e.set_letkf_obs(bad)
    assert ei.value.code == sl.capi.EINVAL
    e.set_letkf_obs(obs)   # the handle stays usable
    e.close()


def test_das_without_logp_threshold_hits(oracle):
    """logp = NULL through the host-buffer path: the library takes ln(mean pressure) with the HOST's libm, exactly like
    a CPU run, so selection is bit-exact without a caller-supplied table -- also for observations constructed to sit
    on the vertical cut-off (|ln p_obs - ln p| == dist_zero_fac * sigma_v up to rounding), where one ulp of the log
    flips the decision.  > 10^6 candidate (point, observation) pairs."""
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=8, nlon=24, nlat=24, nlev=6, nsonde=40, nsfc=100, hloc=200.0e3)
    k = cfg.MEMBER
    g = synth.rng(77, 3)
    nobs = len(obs["elm"])
    sel = np.nonzero(obs["elm"] != 14593)[0]
    nij1, nlev = hgt1.shape
    dz = cfg.dist_zero_fac * cfg.VERT_LOCAL[0]
    for n in sel[: 3000]:
        ij, il = int(g.integers(0, nij1)), int(g.integers(0, nlev))
        pm = gues[ij, il, k, cfg.iv3d_p - 1]
        sgn = 1.0 if g.uniform() < 0.5 else -1.0
        obs["lev"][n] = np.exp(np.log(pm) + sgn * dz * (1.0 + g.integers(-2, 3) * 2.2e-16))
        obs["ri"][n] = rig1[ij] + g.uniform(-2.0, 2.0)
        obs["rj"][n] = rjg1[ij] + g.uniform(-2.0, 2.0)
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    ref = o.das_letkf(gues.copy(order="F"), want_nobsl=True)
    out = e.das_letkf(gues.copy(order="F"), want_nobsl=True)          # no logp
    assert nij1 * nlev * nobs > 1e6
    assert np.array_equal(out["nobsl"], ref["nobsl"])
    a, b = out["anal3d"][:, :, :k, :], ref["anal3d"][:, :, :k, :]
    assert relerr(a, b, axis=(0, 1, 2)) <= TOL
    e.close()


def test_das_infl3d_written_everywhere(oracle):
    """radar-only case: points above RADAR_ZMAX + cut-off have beta == 0 and are skipped, yet the inflation field is
    returned complete (work3d = INFL_MUL clamped by INFL_MUL_MIN at every point, letkf_tools.f90:240-267)"""
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=10, max_nobs=30, nlon=24, nlat=24, nlev=8)
    cfg.RADAR_ZMAX = 3000.0
    cfg.INFL_MUL_ADAPTIVE = 1
    cfg.INFL_MUL = 1.01
    cfg.INFL_MUL_MIN = 1.03
    o, e = _engines(cfg, rig1, rjg1, hgt1, obs, oracle)
    shp = (gues.shape[0], gues.shape[1], gues.shape[3])
    i1 = np.full(shp, max(cfg.INFL_MUL, cfg.INFL_MUL_MIN), order="F")
    i2 = np.full(shp, np.nan, order="F")          # the library must overwrite every entry
    ref = o.das_letkf(gues.copy(order="F"), infl3d=i1, want_nobsl=True)
    out = e.das_letkf(gues.copy(order="F"), infl3d=i2, want_nobsl=True, logp=host_logp(cfg, gues))
    assert (ref["nobsl"] == 0).any() and (ref["nobsl"] > 0).any()
    assert np.isfinite(i2).all()
    assert relerr(i2, i1) <= TOL
    e.close()


def test_das_rejects_2d_inflation_field():
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=6, nsonde=5, nsfc=10)
    cfg.nv2d = 1
    cfg.INFL_MUL_ADAPTIVE = 1
    e = sl.LETKF(cfg, device=0)
    e.set_letkf_obs(obs)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    g2 = np.zeros((gues.shape[0], gues.shape[2], 1), order="F")
    i3 = np.ones((gues.shape[0], gues.shape[1], gues.shape[3]), order="F")
    with pytest.raises(sl.LetkfError):
        e.das_letkf(gues.copy(order="F"), gues2d=g2, infl3d=i3)
    e.close()


def test_p2p_transposes_single_rank_match_three_pass(oracle):
    """one-pass scatter / gather (letkf_b200_scatter_grd_p2p / _gather_grd_p2p, np = 1: the peer is the rank itself)
    against the pack -> copy -> unpack path and the oracle's transposes, with and without the fused state transform"""
    import torch
    from scale_letkf_b200.transpose import EnsTranspose, EnsTransposeP2P
    cfg = synth.config_c2(nlon=37, nlat=21, nlev=35, member=5)
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    e = sl.LETKF(cfg, device=0)
    k, nens, nv, nlev = cfg.MEMBER, cfg.MEMBER + 1, cfg.nv3d, cfg.nlev
    nij1 = cfg.nlon * cfg.nlat
    g = synth.rng(5, 31)
    dev = torch.device("cuda", 0)
    for thermo in (None, e.thermo_defaults()):
        grids = []
        for m in range(k):
            a = np.abs(g.standard_normal((nv, cfg.nlat, cfg.nlon, nlev))) + 0.5      # positive: valid restart variables
            a[5:] *= 1e-3
            grids.append(torch.from_numpy(a.reshape(-1)).to(dev))
        ref = EnsTranspose(e, 1, 0, nlev, nv, 0, device=dev, thermo=thermo)
        new = EnsTransposeP2P(e, 1, 0, thermo=thermo)
        v_ref = torch.zeros((nv, nens, nlev, nij1), dtype=torch.float64, device=dev)
        v_new = torch.full_like(v_ref, -1.0)
        ref.read_ens(grids, None, v_ref, None, k, nens)
        new.read_ens(grids, v_new, k, nens)
        assert torch.equal(v_new[:, :k], v_ref[:, :k])
        out_ref = [torch.zeros_like(x) for x in grids]
        out_new = [torch.full_like(x, -1.0) for x in grids]
        ref.write_ens(v_ref, None, out_ref, None, k, nens)
        new.write_ens(v_new, out_new, k, nens)
        for a, b in zip(out_new, out_ref):
            assert torch.equal(a, b)
    e.close()
