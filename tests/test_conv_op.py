"""Conventional (prepbufr) observation operator and monit_obs: phys2ijk + Trans_XtoY
(scale/common/common_obs_scale.f90:999-1110, 264-337, prsadj :600-617) as obsope_cal calls them (scale/obs/obsope_tools.f90:
466-473) and the observation loop of monit_obs (:1516-1572) followed by monit_dep (:1851-1895).

CPU: the oracle restatement (oracle/oracle_conv.cpp) against an INDEPENDENT vectorised numpy evaluation of the same Fortran
formulas written here (no oracle code shared).  GPU: letkf_b200_obsope_conv / letkf_b200_monit_obs_set against the oracle, QC
flags exact, values to 1e-12 (ln / pow of CUDA vs glibc differ by an ulp; the level index amplifies ln by ~1/dln p)."""
import numpy as np
import pytest

from scale_letkf_b200 import capi

RD, RV, GG = 287.05, 461.50, 9.81
FVIRT = RV / RD - 1.0
U, V, T, TV, Q, RH, PS = 2819, 2820, 3073, 3074, 3330, 3331, 14593


def make_case(nobs=3000, nmem=3, nlev=18, nlon=17, nlat=15, halo=2, seed=3, stggrd=0):
    g = np.random.Generator(np.random.PCG64(20260400 + seed))
    nlevh, nlonh, nlath = nlev + 2 * halo, nlon + 2 * halo, nlat + 2 * halo
    zlev = np.concatenate([np.zeros(halo), 100.0 + 450.0 * np.arange(nlev) ** 1.1, np.zeros(halo)])
    g3, g2 = [], []
    for m in range(nmem):
        v = np.zeros((nlevh, nlonh, nlath, 13), order="F")
        w = np.zeros((nlonh, nlath, 7), order="F")
        topo = np.abs(40.0 * g.standard_normal((nlonh, nlath)))
        hgt = zlev[:, None, None] + topo[None]
        p = 1.0e5 * np.exp(-hgt / 7600.0) * (1.0 + 2e-3 * g.standard_normal(hgt.shape))
        p[:halo] = -1.0                       # halo levels carry no valid pressure (p_full >= 0 marks valid levels)
        p[halo + nlev:] = -1.0
        if m == 0:
            p[halo, 3:6, 4:7] = -1.0          # a patch whose lowest model level is below ground: ks moves up
        v[..., 0] = 8.0 + 4.0 * g.standard_normal(hgt.shape)
        v[..., 1] = -2.0 + 4.0 * g.standard_normal(hgt.shape)
        v[..., 3] = 290.0 - 6.0e-3 * hgt + g.standard_normal(hgt.shape)
        v[..., 4] = p
        v[..., 5] = 9e-3 * np.exp(-hgt / 2800.0) * (1.0 + 0.1 * g.standard_normal(hgt.shape))
        v[..., 11] = np.clip(0.6 + 0.2 * g.standard_normal(hgt.shape), 0.0, 1.1)
        v[..., 12] = hgt
        w[..., 0] = topo
        w[..., 1] = 1.0e5 * np.exp(-topo / 7600.0)
        w[..., 5] = 288.0 + g.standard_normal(topo.shape)
        w[..., 6] = 8e-3 * (1.0 + 0.1 * g.standard_normal(topo.shape))
        g3.append(v)
        g2.append(w)
    elm = g.choice([U, V, T, TV, Q, RH, PS, 4001], size=nobs, p=[0.17, 0.17, 0.17, 0.1, 0.15, 0.1, 0.12, 0.02]).astype(np.int32)
    ril = g.uniform(0.6, nlonh + 0.4, nobs)       # a few outside of the halo'ed domain
    rjl = g.uniform(0.6, nlath + 0.4, nobs)
    if stggrd:                                     # the staggered u / v interpolation reaches half a cell further west / south
        ril = np.maximum(ril, 1.6)
        rjl = np.maximum(rjl, 1.6)
    lev = np.exp(g.uniform(np.log(2.0e4), np.log(1.03e5), nobs))      # some above the model top / below the surface
    ps = elm == PS
    lev[ps] = g.uniform(0.0, 260.0, ps.sum())      # station height [m]; some beyond PS_ADJUST_THRES of the model topography
    ang = g.uniform(-0.1, 0.1, nobs)
    rotc = np.ascontiguousarray(np.stack([np.cos(ang), np.sin(ang)]))
    c = capi.ConvConfig()
    c.nlevh, c.nlonh, c.nlath, c.nlev, c.KHALO, c.nv3dd, c.nv2dd, c.stggrd = nlevh, nlonh, nlath, nlev, halo, 13, 7, stggrd
    c.PS_ADJUST_THRES = 100.0
    return c, elm, ril, rjl, lev, g3, g2, rotc


# ---- independent numpy evaluation ------------------------------------------------------------------------------------
def np_itpl3(var, r1, r2, r3):
    i, j, k = int(np.ceil(r1)), int(np.ceil(r2)), int(np.ceil(r3))
    a, b, c = r1 - (i - 1), r2 - (j - 1), r3 - (k - 1)
    s = var[i - 2:i, j - 2:j, k - 2:k]            # 0-based corners (i-1, i) x (j-1, j) x (k-1, k) of the 1-based indices
    wa, wb, wc = np.array([1 - a, a]), np.array([1 - b, b]), np.array([1 - c, c])
    return float(np.einsum("abc,a,b,c->", s, wa, wb, wc))


def np_itpl2(var, ri, rj):
    i, j = int(np.ceil(ri)), int(np.ceil(rj))
    a, b = ri - (i - 1), rj - (j - 1)
    return float(np.array([1 - a, a]) @ var[i - 2:i, j - 2:j] @ np.array([1 - b, b]))


def np_operator(c, elm, ri, rj, lev, v3, v2, r1, r2):
    """-> (yobs, qc) of one observation on one member"""
    nlev, kh = c.nlev, c.KHALO
    if ri < 1.0 or ri > c.nlonh or rj < 1.0 or rj > c.nlath:
        return None, 98
    if elm > 9999:
        rk = lev
    else:
        i, j = int(np.ceil(ri)), int(np.ceil(rj))
        p = v3[:, i - 2:i, j - 2:j, 4]
        valid = p[kh:kh + nlev] >= 0.0                               # (nlev, 2, 2)
        ks = kh + 1 + int(np.argmax(valid, axis=0).max())            # 1-based lowest level valid in all four columns
        a, b = ri - (i - 1), rj - (j - 1)
        with np.errstate(invalid="ignore", divide="ignore"):
            plev = np.einsum("kab,a,b->k", np.log(p), np.array([1 - a, a]), np.array([1 - b, b]))    # plev[k-1] = level k
        lr = np.log(lev)
        if lr < plev[nlev + kh - 1]:
            return None, 20
        if lr > plev[ks - 1]:
            return None, 21
        below = np.nonzero(plev[ks:nlev + kh] < lr)[0]
        k = ks + 1 + int(below[0]) if len(below) else nlev + kh
        rk = (k - 1) + (lr - plev[k - 2]) / (plev[k - 1] - plev[k - 2])
    qc = 0
    if elm in (U, V):
        if c.stggrd == 1:
            u, v = np_itpl3(v3[..., 0], rk, ri - 0.5, rj), np_itpl3(v3[..., 1], rk, ri, rj - 0.5)
        else:
            u, v = np_itpl3(v3[..., 0], rk, ri, rj), np_itpl3(v3[..., 1], rk, ri, rj)
        y = u * r1 - v * r2 if elm == U else u * r2 + v * r1
    elif elm == T:
        y = np_itpl3(v3[..., 3], rk, ri, rj)
    elif elm == TV:
        y = np_itpl3(v3[..., 3], rk, ri, rj) * (1.0 + FVIRT * np_itpl3(v3[..., 5], rk, ri, rj))
    elif elm == Q:
        y = np_itpl3(v3[..., 5], rk, ri, rj)
    elif elm == RH:
        y = np_itpl3(v3[..., 11], rk, ri, rj)
    elif elm == PS:
        t, q, topo = np_itpl2(v2[..., 5], ri, rj), np_itpl2(v2[..., 6], ri, rj), np_itpl2(v2[..., 0], ri, rj)
        y = np_itpl2(v2[..., 1], ri, rj)
        dz = rk - topo
        if dz != 0:
            tv = t * (1.0 + 0.608 * q)
            y = y * ((-5.0e-3 * dz + tv) / tv) ** (GG / (5.0e-3 * RD))
        if abs(dz) > c.PS_ADJUST_THRES:
            qc = 10
    else:
        return None, 90
    return y, qc


@pytest.mark.parametrize("stggrd", [0, 1])
def test_oracle_conv_operator_matches_numpy(oracle, stggrd):
    c, elm, ril, rjl, lev, g3, g2, rotc = make_case(nobs=1500, stggrd=stggrd)
    y, q = oracle.obsope_conv(c, elm, ril, rjl, lev, g3, g2, rotc=rotc)
    seen = set()
    for m in range(len(g3)):
        for n in range(len(elm)):
            want, wq = np_operator(c, int(elm[n]), ril[n], rjl[n], lev[n], g3[m], g2[m], rotc[0, n], rotc[1, n])
            assert q[n, m] == wq, (n, m, elm[n], q[n, m], wq)
            seen.add(wq)
            if want is None:
                assert y[n, m] == capi.UNDEF
            else:
                assert abs(y[n, m] - want) <= 1e-12 * max(abs(want), 1e-300), (n, m, elm[n], y[n, m], want)
    assert seen == {0, 10, 20, 21, 90, 98}         # every QC outcome of the operator occurs in the case


def _monit_sets(seed=3):
    c, elm, ril, rjl, lev, g3, g2, rotc = make_case(nobs=2500, nmem=1, seed=seed, stggrd=1)
    from test_radar_op import make_case as radar_make_case
    g = np.random.Generator(np.random.PCG64(99 + seed))
    conv = dict(cfg=c, elm=elm, ril=ril, rjl=rjl, lev=lev, rotc=rotc, dat=g.standard_normal(len(elm)) + 280.0,
                dif=g.uniform(-4000.0, 4000.0, len(elm)))
    # a radar set on the SAME grid: Trans_XtoY_radar reads the 13 history variables of the conventional case
    r, relm, rril, rrjl, rlon, rlat, rlev, rgrids, rrotc = radar_make_case(nobs=1200, nmem=1, nlev=c.nlev, nlon=c.nlonh - 4,
                                                                          nlat=c.nlath - 4, halo=c.KHALO, seed=seed)
    v3 = g3[0].copy(order="F")
    v3[..., 6:11] = rgrids[0][..., 6:11]          # hydrometeors from the radar case, everything else from the conventional one
    radar = dict(cfg=r, elm=relm, ril=rril, rjl=rrjl, lon=rlon, lat=rlat, lev=np.minimum(rlev, 6000.0), rotc=rrotc,
                 dat=g.uniform(0.0, 40.0, len(relm)), dif=g.uniform(-4000.0, 4000.0, len(relm)))
    return [conv, radar], v3, g2[0]


def test_oracle_monit_obs_matches_numpy(oracle):
    sets, v3, v2 = _monit_sets()
    t_range = 3000.0
    out = oracle.monit_obs(sets, v3, v2, t_range=t_range)
    # conventional part: numpy operator + the rules of the observation loop; statistics by plain numpy
    st = sets[0]
    c = st["cfg"]
    n0 = len(st["elm"])
    for n in range(n0):
        if abs(st["dif"][n]) > t_range:
            assert out["oqc"][n] == -1
            continue
        want, wq = np_operator(c, int(st["elm"][n]), st["ril"][n], st["rjl"][n], st["lev"][n], v3, v2, st["rotc"][0, n], st["rotc"][1, n])
        assert out["oqc"][n] == wq
        if wq == 0:
            assert abs(out["ohx"][n] - (st["dat"][n] - want)) <= 1e-12 * max(abs(want), 1.0)
        else:
            assert out["ohx"][n] == capi.UNDEF
    # radar part: the radar operator of the oracle (pinned by tests/test_radar_op.py) without the RADAR_ZMAX test
    st = sets[1]
    r = capi.RadarConfig.from_buffer_copy(st["cfg"])
    r.RADAR_ZMAX = 1e300
    y, q = oracle.obsope_radar(r, st["elm"], st["ril"], st["rjl"], st["lon"], st["lat"], st["lev"], [v3], rotc=st["rotc"])
    keep = np.abs(st["dif"]) <= t_range
    assert np.array_equal(out["oqc"][n0:][keep], q[keep, 0]) and (out["oqc"][n0:][~keep] == -1).all()
    good = keep & (q[:, 0] == 0)
    assert np.array_equal(out["ohx"][n0:][good], st["dat"][good] - y[good, 0]) and good.sum() > 100
    # monit_dep over everything (element uid: Tv counts as T, RE0 as REF -- common_obs_scale.f90:1866-1875)
    uid = {U: 1, V: 2, T: 3, TV: 3, Q: 5, RH: 6, PS: 7, 4001: 9, 4004: 9, 4002: 11}
    for e_uid in set(uid.values()):
        sel = np.array([uid.get(int(e), -1) == e_uid for e in out["elm"]]) & (out["oqc"] == 0)
        assert out["nobs"][e_uid - 1] == sel.sum()
        if sel.sum():
            d = out["ohx"][sel]
            assert abs(out["bias"][e_uid - 1] - d.mean()) <= 1e-12 * max(abs(d).max(), 1.0)
            assert abs(out["rmse"][e_uid - 1] - np.sqrt((d * d).mean())) <= 1e-12 * max(abs(d).max(), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("stggrd", [0, 1])
def test_gpu_conv_operator_matches_oracle(oracle, space, stggrd):
    import torch
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    c, elm, ril, rjl, lev, g3, g2, rotc = make_case(nobs=6000, nmem=4, stggrd=stggrd)
    y0, q0 = oracle.obsope_conv(c, elm, ril, rjl, lev, g3, g2, rotc=rotc)
    eng = sl.LETKF(synth.config_c2(nlon=8, nlat=8, nlev=4, member=4), device=0)
    if space == "device":
        d3 = [torch.from_numpy(np.ascontiguousarray(np.transpose(a, (3, 2, 1, 0)))).cuda() for a in g3]   # F-order memory
        d2 = [torch.from_numpy(np.ascontiguousarray(np.transpose(a, (2, 1, 0)))).cuda() for a in g2]
        y, q = eng.obsope_conv(c, elm, ril, rjl, lev, d3, d2, rotc=rotc)
        y, q = y.cpu().numpy(), q.cpu().numpy()
    else:
        y, q = eng.obsope_conv(c, elm, ril, rjl, lev, g3, g2, rotc=rotc)
    eng.close()
    assert np.array_equal(q, q0)
    good = q0 != 98
    ok = (q0 == 0) | (q0 == 10)
    assert np.array_equal(y[~ok], y0[~ok])                                         # undef where the operator gave up
    for e in (U, V, T, TV, Q, RH, PS):               # per element: wind components pass through zero, so the scale is the field's
        sel = ok & (elm == e)[:, None]
        assert sel.any()
        assert np.abs(y[sel] - y0[sel]).max() <= 1e-12 * np.abs(y0[sel]).max(), e
    assert ok.sum() > 0.5 * q0.size and good.any()


@pytest.mark.gpu
def test_gpu_monit_obs_matches_oracle(oracle):
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    sets, v3, v2 = _monit_sets(seed=4)
    want = oracle.monit_obs(sets, v3, v2, t_range=3000.0)
    eng = sl.LETKF(synth.config_c2(nlon=8, nlat=8, nlev=4, member=4), device=0)
    got = eng.monit_obs(sets, v3, v2, t_range=3000.0)
    eng.close()
    assert np.array_equal(got["oqc"], want["oqc"]) and np.array_equal(got["nobs"], want["nobs"])
    good = want["oqc"] == 0
    assert np.array_equal(got["ohx"][~good], want["ohx"][~good])
    conv = np.isin(want["elm"], [U, V, T, TV, Q, RH, PS])
    sc = np.maximum(np.abs(want["ohx"]), 1.0)
    assert (np.abs(got["ohx"] - want["ohx"])[good & conv] <= 1e-11 * sc[good & conv]).all()
    assert (np.abs(got["ohx"] - want["ohx"])[good & ~conv] <= 1e-7 * sc[good & ~conv]).all()      # Doppler velocity: see test_radar_op.py
    m = want["nobs"] > 0
    assert (np.abs(got["bias"] - want["bias"])[m] <= 1e-7 * np.maximum(np.abs(want["bias"][m]), 1.0)).all()
    assert (np.abs(got["rmse"] - want["rmse"])[m] <= 1e-7 * np.maximum(np.abs(want["rmse"][m]), 1.0)).all()
    assert m.sum() >= 6
