"""Pins for the oracle's bucket tables, obs_local and das_letkf."""
import numpy as np
import pytest

from scale_letkf_b200 import synth


def small_sonde_case(member=8, nlon=24, nlat=24, nlev=6, nsonde=12, nsfc=40, **kw):
    cfg = synth.config_c2(nlon=nlon, nlat=nlat, nlev=nlev, member=member, **kw)
    cfg.HORI_LOCAL[0] = 60.0e3
    synth.resolve_config(cfg)
    for t in range(1, 24):
        cfg.HORI_LOCAL[t] = cfg.HORI_LOCAL[0]
    rig1, rjg1, hgt1 = synth.make_grid(cfg, topo_amp=300.0)
    obs = synth.make_sonde_obs(cfg, nsonde, nsfc, nlevobs=8, seed_no=40)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=41)
    return cfg, rig1, rjg1, hgt1, obs, gues


def small_radar_case(member=8, nlon=40, nlat=40, nlev=8, max_nobs=30, **kw):
    cfg = synth.config_c3(nlon=nlon, nlat=nlat, nlev=nlev, member=member, max_nobs=max_nobs, **kw)
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    obs = synth.make_radar_obs(cfg, radius_m=7.0e3, zmin=500.0, zmax=6000.0, dz=1000.0, seed_no=42)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=43)
    return cfg, rig1, rjg1, hgt1, obs, gues


def test_bucket_tables(oracle):
    cfg, rig1, rjg1, hgt1, obs, _ = small_sonde_case()
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    ntot, nctype = o.obs_info()
    assert ntot == len(obs["elm"])
    assert nctype == 5   # U,V,T,Q of ADPUPA + PS of ADPSFC
    s2o = o.sorted_index()
    assert sorted(s2o.tolist()) == list(range(ntot))
    begin = 0
    for ic in range(nctype):
        info = o.ctype(ic)
        ac = o.ac_ext(ic)   # (ngrdext_j, ngrdext_i+1)
        assert info.ac_begin == begin
        assert ac[0, 0] == begin
        flat = ac.ravel()
        assert np.all(np.diff(flat) >= 0)
        assert np.array_equal(ac[1:, 0], ac[:-1, -1])   # rows chained (letkf_obs.f90:948-950)
        assert ac[-1, -1] - begin == info.tot_ext
        # every obs of a bucket has this ctype and lies in that bucket; stable in-bucket order
        seg = s2o[begin:begin + info.tot_ext]
        assert np.all(obs["elm"][seg] == info.elm) and np.all(obs["typ"][seg] == info.typ)
        for j in range(info.ngrdext_j):
            for i in range(1, info.ngrdext_i + 1):
                ids = s2o[ac[j, i - 1]:ac[j, i]]
                assert np.all(np.diff(ids) > 0)
                for n in ids:
                    gi = int(np.ceil((obs["ri"][n] - cfg.IHALO - 0.5) * info.ngrd_i / cfg.nlon))
                    gj = int(np.ceil((obs["rj"][n] - cfg.JHALO - 0.5) * info.ngrd_j / cfg.nlat))
                    gi = min(max(gi, 1), info.ngrd_i) + info.ngrdsch_i
                    gj = min(max(gj, 1), info.ngrd_j) + info.ngrdsch_j
                    assert (gi, gj) == (i, j + 1)
        begin += info.tot_ext


def _points(cfg, rig1, rjg1, hgt1, gues, stride=7):
    nij1, nlev = hgt1.shape
    k = cfg.MEMBER
    ri, rj, rlev, rz = [], [], [], []
    for il in range(nlev):
        for ij in range(0, nij1, stride):
            ri.append(rig1[ij]); rj.append(rjg1[ij])
            rlev.append(gues[ij, il, k, cfg.iv3d_p - 1]); rz.append(hgt1[ij, il])
    return map(np.array, (ri, rj, rlev, rz))


def test_search_semantic_vs_bruteforce_nolimit(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = small_sonde_case()
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    ri, rj, rlev, rz = _points(cfg, rig1, rjg1, hgt1, gues)
    n1, i1, d1, l1 = o.obs_local(ri, rj, rlev, rz, 1, 2048)
    n2, i2, d2, l2 = o.obs_local(ri, rj, rlev, rz, 1, 2048, brute=True)
    assert n1.max() > 0 and n1.min() == 0 or n1.max() > 0
    assert np.array_equal(n1, n2)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2) and np.array_equal(l1, l2)


@pytest.mark.parametrize("criterion", [1, 2, 3])
def test_search_semantic_vs_bruteforce_limited(oracle, criterion):
    cfg, rig1, rjg1, hgt1, obs, gues = small_radar_case()
    cfg.MAX_NOBS_PER_GRID_CRITERION = criterion
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    ri, rj, rlev, rz = _points(cfg, rig1, rjg1, hgt1, gues, stride=11)
    n1, i1, d1, l1 = o.obs_local(ri, rj, rlev, rz, 1, 4096)
    n2, i2, d2, l2 = o.obs_local(ri, rj, rlev, rz, 1, 4096, brute=True)
    assert np.array_equal(n1, n2)
    assert n1.max() == 60   # REF+RE0 budget 30 + VR budget 30
    for p in range(len(ri)):
        assert sorted(i1[p, :n1[p]].tolist()) == sorted(i2[p, :n2[p]].tolist())


def test_das_no_obs_is_identity_with_inflation(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = small_sonde_case(nsonde=1, nsfc=0)
    # move the only sounding far outside every cut-off: still 4 ctypes, no local obs anywhere
    cfg.HORI_LOCAL[0] = 1.0e3
    for t in range(1, 24):
        cfg.HORI_LOCAL[t] = 1.0e3
    cfg.RELAX_ALPHA_SPREAD = 0.0
    cfg.INFL_MUL = 1.21
    o = oracle.Oracle(cfg)
    obs["ri"][:] = 3.3
    obs["rj"][:] = 3.3
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    g0 = gues.copy(order="F")
    r = o.das_letkf(gues, want_nobsl=True)
    k = cfg.MEMBER
    far = r["nobsl"] == 0
    assert far.sum() > 0
    mean = g0[:, :, k, :]
    for m in range(k):
        expect = mean + np.sqrt(1.21) * (g0[:, :, m, :] - mean)
        got = r["anal3d"][:, :, m, :]
        assert np.allclose(got[far], expect[far], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("relax", ["rtps", "rtpp", "none"])
def test_das_matches_pointwise_numpy(oracle, relax):
    """das_letkf == per-point numpy LETKF built from obs_local + letkf_core (two code paths)."""
    cfg, rig1, rjg1, hgt1, obs, gues = small_sonde_case(member=6, nlon=12, nlat=12, nlev=4)
    cfg.RELAX_ALPHA_SPREAD = 0.95 if relax == "rtps" else 0.0
    cfg.RELAX_ALPHA = 0.7 if relax == "rtpp" else 0.0
    cfg.BOUNDARY_BUFFER_WIDTH = 40.0e3
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    k = cfg.MEMBER
    g0 = gues.copy(order="F")
    r = o.das_letkf(gues, want_nobsl=True, want_rtps=True)
    assert r["status"] == 0
    s2o = o.sorted_index()
    ens_sorted = obs["ensval"][s2o]
    val_sorted = obs["val"][s2o]
    nij1, nlev = hgt1.shape
    checked = 0
    for il in range(nlev):
        for ij in range(0, nij1, 5):
            pm = g0[ij, il, k, cfg.iv3d_p - 1]
            n, idx, rd, rl = o.obs_local([rig1[ij]], [rjg1[ij]], [pm], [hgt1[ij, il]], 1, 4096)
            p = int(n[0])
            assert p == r["nobsl"][ij, il]
            beta = min(1.0, min(min(rig1[ij] - cfg.IHALO, cfg.nlon + cfg.IHALO + 1 - rig1[ij]) * cfg.DX,
                                min(rjg1[ij] - cfg.JHALO, cfg.nlat + cfg.JHALO + 1 - rjg1[ij]) * cfg.DY)
                       / cfg.BOUNDARY_BUFFER_WIDTH)
            y = ens_sorted[idx[0, :p], :k]
            a = y.T @ (y / rd[0, :p, None]) + (k - 1) * np.eye(k)
            lam, v = np.linalg.eigh(a)
            pa = (v / lam) @ v.T
            w = (v * np.sqrt((k - 1) / lam)) @ v.T
            wm = pa @ (y.T @ (val_sorted[idx[0, :p]] / rd[0, :p]))
            for nvar in range(cfg.nv3d):
                dx = g0[ij, il, :k, nvar] - g0[ij, il, k, nvar]
                if relax == "rtps":
                    vg, va = dx @ dx, dx @ pa @ dx
                    f = 0.95 * np.sqrt(vg / (va * (k - 1))) - 0.95 + 1.0 if (vg > 0 and va > 0) else 1.0
                    wr = w * f
                elif relax == "rtpp":
                    wr = 0.3 * w + 0.7 * np.eye(k)
                else:
                    wr = w
                t = (wr + wm[:, None]) * beta + (1 - beta) * np.eye(k)
                xa = g0[ij, il, k, nvar] + dx @ t
                got = r["anal3d"][ij, il, :k, nvar]
                assert np.allclose(got, xa, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(xa).max()))
            checked += 1
    assert checked > 10
    # gues is destroyed into perturbations with the mean kept (letkf_tools.f90:209-230)
    assert np.allclose(gues[:, :, :k, :], g0[:, :, :k, :] - g0[:, :, k:k + 1, :], rtol=0, atol=0)


def test_das_mask_equals_full(oracle):
    cfg, rig1, rjg1, hgt1, obs, gues = small_sonde_case(member=6, nlon=12, nlat=12, nlev=4)
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    g1, g2 = gues.copy(order="F"), gues.copy(order="F")
    full = o.das_letkf(g1)
    mask = np.zeros(hgt1.shape, dtype=np.uint8)
    mask[::3, :] = 1
    part = o.das_letkf(g2, point_mask=mask)
    assert part["npoints"] == mask.sum()
    sel = mask.astype(bool)
    k = cfg.MEMBER
    assert np.array_equal(part["anal3d"][:, :, :k, :][sel], full["anal3d"][:, :, :k, :][sel])


def test_transpose_roundtrip(oracle):
    import ctypes as C
    nlon, nlat, nlev, nv3d, nv2d, np_ = 7, 5, 3, 2, 1, 4
    g = synth.rng(77)
    nens = np_
    nlevall = nlev * nv3d + nv2d
    _, nmax = oracle.nij1(nlon, nlat, np_, 0)
    v3dg = [np.asfortranarray(g.standard_normal((nlev, nlon, nlat, nv3d))) for _ in range(np_)]
    v2dg = [np.asfortranarray(g.standard_normal((nlon, nlat, nv2d))) for _ in range(np_)]
    L = oracle.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    bufs = [np.zeros((nmax, nlevall, np_), order="F") for _ in range(np_)]
    for r in range(np_):
        L.oracle_grd_to_buf(nlon, nlat, nlev, nv3d, nv2d, np_, P(v3dg[r]), P(v2dg[r]), P(bufs[r]))
    # all-to-all: rank r receives block r of every sender s into slot s
    bufr = [np.asfortranarray(np.stack([bufs[s][:, :, r] for s in range(np_)], axis=2)) for r in range(np_)]
    back = []
    for r in range(np_):
        n1, _ = oracle.nij1(nlon, nlat, np_, r)
        v3d = np.zeros((n1, nlev, nens, nv3d), order="F")
        v2d = np.zeros((n1, nens, nv2d), order="F")
        L.oracle_buf_to_ens(nlon, nlat, nlev, nv3d, nv2d, np_, r, nens, 1, np_, P(bufr[r]), P(v3d), P(v2d))
        ilon, ilat = synth.column_deal(nlon, nlat, np_, r)
        for m in range(np_):
            assert np.array_equal(v3d[:, :, m, :], v3dg[m][:, ilon - 1, ilat - 1, :].transpose(1, 0, 2))
            assert np.array_equal(v2d[:, m, :], v2dg[m][ilon - 1, ilat - 1, :])
        b = np.zeros((nmax, nlevall, np_), order="F")
        L.oracle_ens_to_buf(nlon, nlat, nlev, nv3d, nv2d, np_, r, nens, 1, np_, P(v3d), P(v2d), P(b))
        back.append(b)
    for r in range(np_):
        br = np.asfortranarray(np.stack([back[s][:, :, r] for s in range(np_)], axis=2))
        o3 = np.zeros_like(v3dg[r]); o2 = np.zeros_like(v2dg[r])
        L.oracle_buf_to_grd(nlon, nlat, nlev, nv3d, nv2d, np_, P(br), P(o3), P(o2))
        assert np.array_equal(o3, v3dg[r]) and np.array_equal(o2, v2dg[r])


def test_das_radar_matches_independent_bruteforce_numpy(oracle):
    """C3-type case (radar, MAX_NOBS_PER_GRID(22) > 0, REF+RE0 merged budget, RTPS, boundary taper, radar lid):
    das_letkf of the oracle against a restatement that shares NO code with oracle/ -- brute-force selection over all
    observations (tests/golden/make_golden.py:select_bruteforce), LAPACK eigh for letkf_core, numpy for the rest."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import select_bruteforce
    from helpers import radar_case
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=6, nlon=24, nlat=24, nlev=5, max_nobs=12, seed=902, radius=5.0e3)
    o = oracle.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    k = cfg.MEMBER
    g0 = gues.copy(order="F")
    r = o.das_letkf(gues.copy(order="F"), want_nobsl=True)
    assert r["status"] == 0
    dzf = cfg.dist_zero_fac
    zcut = cfg.RADAR_ZMAX + max(cfg.VERT_LOCAL[21], cfg.VERT_LOCAL_RADAR_VR) * dzf
    nij1, nlev = hgt1.shape
    checked = solved = 0
    for il in range(nlev):
        for ij in range(0, nij1, 13):
            ri, rj, rz = rig1[ij], rjg1[ij], hgt1[ij, il]
            if rz > zcut:
                beta = 0.0
            else:
                d = min(min(ri - cfg.IHALO, cfg.nlon + cfg.IHALO + 1 - ri) * cfg.DX,
                        min(rj - cfg.JHALO, cfg.nlat + cfg.JHALO + 1 - rj) * cfg.DY) / cfg.BOUNDARY_BUFFER_WIDTH
                beta = 1.0 if d >= 1.0 else max(d, 0.0)
            mean = g0[ij, il, k, :]
            dx = g0[ij, il, :k, :] - mean                      # (k, nv)
            got = r["anal3d"][ij, il, :k, :]
            if beta == 0.0:
                assert np.array_equal(got, mean + dx)
                checked += 1
                continue
            pm = g0[ij, il, k, cfg.iv3d_p - 1]
            ids = select_bruteforce(cfg, obs, (np.array([ri]), np.array([rj]), np.array([pm]), np.array([rz])), 1)[0]
            assert len(ids) == r["nobsl"][ij, il]
            if len(ids) == 0:
                xa = mean + dx                                  # infl = 1: W = I, wbar = 0
            else:
                rdiag = []
                for n in ids:
                    e = obs["elm"][n]
                    hl = cfg.HORI_LOCAL_RADAR_OBSNOREF if e == 4004 else cfg.HORI_LOCAL_RADAR_VR if e == 4002 else cfg.HORI_LOCAL[21]
                    vl = cfg.VERT_LOCAL_RADAR_VR if e == 4002 else cfg.VERT_LOCAL[21]
                    nd_v = abs(obs["lev"][n] - rz) / vl
                    nd_h = np.sqrt(((ri - obs["ri"][n]) * cfg.DX) ** 2 + ((rj - obs["rj"][n]) * cfg.DY) ** 2) / hl
                    rdiag.append(obs["err"][n] ** 2 / np.exp(-0.5 * (nd_h * nd_h + nd_v * nd_v)))
                rdiag = np.array(rdiag)
                y = obs["ensval"][ids, :k]
                a = y.T @ (y / rdiag[:, None]) + (k - 1) * np.eye(k)
                lam, v = np.linalg.eigh(a)
                pa = (v / lam) @ v.T
                w = (v * np.sqrt((k - 1) / lam)) @ v.T
                wm = pa @ (y.T @ (obs["val"][ids] / rdiag))
                xa = np.empty_like(dx)
                for nvar in range(cfg.nv3d):
                    x = dx[:, nvar]
                    vg, va = x @ x, x @ pa @ x
                    f = 0.95 * np.sqrt(vg / (va * (k - 1))) - 0.95 + 1.0 if (vg > 0 and va > 0) else 1.0
                    t = (w * f + wm[:, None]) * beta + (1 - beta) * np.eye(k)
                    xa[:, nvar] = mean[nvar] + x @ t
                solved += 1
            sc = np.maximum(np.abs(xa).max(axis=0), 1e-300)
            assert (np.abs(got - xa) / sc).max() <= 1e-10, (ij, il)
            checked += 1
    assert checked > 40 and solved > 5
