"""Radar observation operator: the obsfmt_radar branch of obsope_cal
(scale/obs/obsope_tools.f90:476-494) = phys2ijkz + Trans_XtoY_radar + calc_ref_vr
(scale/common/common_obs_scale.f90:1116-1237, 342-493, 626-990) for all members.

CPU: the oracle restatement (oracle/oracle_radar.cpp) against an INDEPENDENT vectorised numpy evaluation of the same
Fortran formulas written here (no oracle code shared).  GPU: letkf_b200_obsope_radar against the oracle, QC flags
exact, reflectivities to 1e-12 (pow / log10 of CUDA vs glibc differ by a few ulp).  Doppler velocities to 1e-7 only: the
elevation angle comes from com_distll_1 (common/common.f90:401-424), dist = acos(cos d) re, and acos is conditioned
like 1/sin d -- 10^3 at the few-km ranges of a radar volume -- so ONE ulp of sin/cos between two libms (numpy's SIMD
routines, glibc, CUDA) moves vr by ~1e-9.  The formula is the reference's; the tolerance states its conditioning."""
import numpy as np
import pytest

from scale_letkf_b200 import capi

PI = 3.1415926535
RD, GG, RE = 287.05, 9.81, 6371.3e3
F32 = lambda x: float(np.float32(x))


def make_case(nobs=4000, nmem=3, nlev=20, nlon=18, nlat=16, halo=2, seed=5, method=3, use_tv=0):
    g = np.random.Generator(np.random.PCG64(20260300 + seed))
    nlevh, nlonh, nlath = nlev + 2 * halo, nlon + 2 * halo, nlat + 2 * halo
    zlev = np.concatenate([np.full(halo, -999.0), 200.0 + 500.0 * np.arange(nlev) ** 1.15, np.full(halo, -999.0)])
    grids = []
    for m in range(nmem):
        v = np.zeros((nlevh, nlonh, nlath, 13), order="F")
        topo = 30.0 * g.standard_normal((nlonh, nlath))
        hgt = zlev[:, None, None] + topo[None] * np.linspace(1.0, 0.0, nlevh)[:, None, None]
        hgt[:halo] = -999.0
        hgt[halo + nlev:] = -999.0
        v[..., 12] = hgt
        v[..., 0] = 10.0 + 5.0 * g.standard_normal(hgt.shape)
        v[..., 1] = -3.0 + 5.0 * g.standard_normal(hgt.shape)
        v[..., 2] = 0.5 * g.standard_normal(hgt.shape)
        v[..., 3] = 295.0 - 6.5e-3 * np.maximum(hgt, 0.0) + g.standard_normal(hgt.shape)     # crosses 273.16 K
        v[..., 4] = 1.0e5 * np.exp(-np.maximum(hgt, 0.0) / 7500.0) * (1.0 + 1e-3 * g.standard_normal(hgt.shape))
        v[..., 5] = 8e-3 * np.exp(-np.maximum(hgt, 0.0) / 3000.0)
        for n, sc in ((6, 3e-4), (7, 1e-3), (8, 2e-4), (9, 6e-4), (10, 8e-4)):     # hydrometeors: patchy, many exact zeros
            f = sc * np.maximum(g.standard_normal(hgt.shape) - 0.3, 0.0)
            v[..., n] = f
        v[..., 11] = 0.6
        grids.append(v)
    elm = g.choice([4001, 4004, 4002, 2819], size=nobs, p=[0.4, 0.2, 0.38, 0.02]).astype(np.int32)
    ril = g.uniform(0.5, nlonh + 0.5, nobs)       # some outside of the halo'ed domain
    rjl = g.uniform(0.5, nlath + 0.5, nobs)
    lev = g.uniform(0.0, zlev[halo + nlev - 1] * 1.1, nobs)     # some below the lowest / above the highest level
    radar_lon, radar_lat, radar_z = 135.5, 34.8, 120.0
    lon = radar_lon + (ril - nlonh / 2) * 0.005
    lat = radar_lat + (rjl - nlath / 2) * 0.0045
    lon[:3], lat[:3] = radar_lon, radar_lat                      # the radar site itself: iqc_out_h
    ang = g.uniform(-0.1, 0.1, nobs)
    rotc = np.ascontiguousarray(np.stack([np.cos(ang), np.sin(ang)]))
    r = capi.RadarConfig()
    r.METHOD_REF_CALC, r.USE_TERMINAL_VELOCITY = method, use_tv
    r.nlevh, r.nlonh, r.nlath, r.nlev, r.KHALO, r.nv3dd = nlevh, nlonh, nlath, nlev, halo, 13
    r.MIN_RADAR_REF_DBZ, r.LOW_REF_SHIFT, r.RADAR_ZMAX = 5.0, -2.0, float(zlev[halo + nlev - 1] * 1.05)
    r.radar_lon, r.radar_lat, r.radar_z = radar_lon, radar_lat, radar_z
    return r, elm, ril, rjl, lon, lat, lev, grids, rotc


# ---- independent numpy evaluation -------------------------------------------------------------------------
def np_gamma(x):
    G = [1.0, 0.5772156649015329, -0.6558780715202538, -0.420026350340952e-1, 0.1665386113822915, -.421977345555443e-1,
         -.96219715278770e-2, .72189432466630e-2, -.11651675918591e-2, -.2152416741149e-3, .1280502823882e-3,
         -.201348547807e-4, -.12504934821e-5, .11330272320e-5, -.2056338417e-6, .61160950e-8, .50020075e-8,
         -.11812746e-8, .1043427e-9, .77823e-11, -.36968e-11, .51e-12, -.206e-13, -.54e-14, .14e-14, .1e-15]
    z, m = abs(x), int(abs(x))
    rr = 1.0
    for k in range(1, m + 1):
        rr *= (z - k)
    z -= m
    gr = G[25]
    for k in range(24, -1, -1):
        gr = gr * z + G[k]
    return 1.0 / (gr * z) * rr


def np_ref_vr(method, use_tv, qr, qs, qg, u, v, w, t, p, az, elev):
    ro = p / (RD * t)
    P = np.power
    with np.errstate(divide="ignore", invalid="ignore"):
        if method == 1:
            qt = qr + qs + qg
            ref = np.where(qt > 0, 10.0e18 * 72 * P(ro * qt, 1.75) / (P(PI, 1.75) * P(8.0e6, 0.75) * P(1000.0, 1.75)), 0.0)
            wt = np.where(qt > 0, 5.40 * P(1.0e5 / p, F32(0.4)) * P(np.where(qt > 0, qt, 1.0), 0.125), 0.0)
        elif method == 2:
            pip, cf = P(PI, 1.75), 1.0e18 * 720
            zr = np.where(qr > 0, cf * P(ro * qr, 1.75) / (pip * P(8.0e6, 0.75) * P(1000.0, 1.75)), 0.0)
            zs_c = cf * 0.176 * P(100.0, 0.25) * P(ro * qs, 1.75) / (pip * 0.930 * P(3.0e6, 0.75) * 917.0 ** 2)
            zs_w = cf * P(ro * qs, 1.75) / (pip * P(3.0e6, 0.75) * P(917.0, 1.75))
            zs = np.where(qs > 0, np.where(t <= F32(273.16), zs_c, zs_w), 0.0)
            zg = np.where(qg > 0, P(cf / (pip * P(4.0e4, 0.75) * P(913.0, 1.75)), F32(0.95)) * P(ro * qg, F32(1.6625)), 0.0)
            ref = zr + zs + zg
            e3 = F32(1e-3)
            nor, nos, nog, ror, ros, rog, roo, ro2 = 8.0e6 * e3, 3.0e6 * e3, 4.0e4 * e3, 1000.0 * e3, 100.0 * e3, 913.0 * e3, 1.0 * e3, ro * e3
            rof = P(roo / ro2, 0.25)
            wr = np.where(qr > 0, 1.0e-2 * (2115.0 * np_gamma(4.8) / (6.0 * P(P(PI * ror * nor / (ro2 * qr), 0.25), 0.8))) * rof, 0.0)
            ws = np.where(qs > 0, 1.0e-2 * (152.93 * np_gamma(4.25) / (6.0 * P(P(PI * ros * nos / (ro2 * qs), 0.25), 0.25))) * rof, 0.0)
            wg = np.where(qg > 0, 1.0e-2 * (np_gamma(4.5) * P(4.0 * GG * 100.0 * rog / (3.0 * 0.6 * ro2), 0.5)) /
                          (6.0 * P(P(PI * rog * nog / (ro2 * qg), 0.25), 0.5)), 0.0)
            wt = np.where(ref > 0, (wr * zr + ws * zs + wg * zg) / (zr + zs + zg), 0.0)
        else:
            both_g, both_s = (qr > 0) & (qg > 0), (qr > 0) & (qs > 0)
            Fg = np.where(both_g, 0.5 * P(np.minimum(qr / qg, qg / qr), 1.0 / 3.0), 0.0)
            fwg = np.where(both_g, qr / (qr + qg), 0.0)
            Fs = np.where(both_s, 0.5 * P(np.minimum(qr / qs, qs / qr), 1.0 / 3.0), 0.0)
            fws = np.where(both_s, qr / (qr + qs), 0.0)
            qrp, qsp, qgp = (1.0 - Fs - Fg) * qr, (1.0 - Fs) * qs, (1.0 - Fg) * qg
            qms, qmg = Fs * (qr + qs), Fg * (qr + qg)
            zr = np.where(qrp > 0, 2.53e4 * P(ro * qrp * 1.0e3, F32(1.84)), 0.0)
            zs = np.where(qsp > 0, 3.48e3 * P(ro * qsp * 1.0e3, F32(1.66)), 0.0)
            zg = np.where(qgp > 0, 5.54e3 * P(ro * qgp * 1.0e3, F32(1.70)), 0.0)
            zms = np.where(qms > 0, (F32(0.00491) + F32(5.75) * fws - F32(5.588) * fws ** 2) * 1.0e5 *
                           P(ro * qms * 1.0e3, F32(1.67) - F32(0.202) * fws + F32(0.398) * fws ** 2), 0.0)
            zmg = np.where(qmg > 0, (F32(0.809) + F32(10.13) * fwg - F32(5.98) * fwg ** 2) * 1.0e5 *
                           P(ro * qmg * 1.0e3, F32(1.48) + F32(0.0448) * fwg - F32(0.0313) * fwg ** 2), 0.0)
            ref = zr + zg + zs + zms + zmg
            ro2 = 1.0e-3 * ro
            rof = P(0.001 / ro2, 0.5)
            wr = np.where(qr > 0, 1.0e-2 * (2115.0 * np_gamma(4.8) / (6.0 * P(P(PI * 1.0 * 8.0e-2 / (ro2 * qr), 0.25), 0.8))) * rof, 0.0)
            ws = np.where(qs > 0, 1.0e-2 * (152.93 * np_gamma(4.25) / (6.0 * P(P(PI * 0.1 * 3.0e-2 / (ro2 * qs), 0.25), 0.25))) * rof, 0.0)
            wg = np.where(qg > 0, 1.0e-2 * (np_gamma(4.5) * P(4.0 * GG * 100.0 * 0.917 / (3.0 * 0.6 * ro2), 0.5)) /
                          (6.0 * P(P(PI * 0.917 * 4.0e-4 / (ro2 * qg), 0.25), 0.5)), 0.0)
            wt = np.where(ref > 0, (wr * zr + ws * zs + ws * zms + wg * zg + wg * zmg) / (zr + zs + zg + zms + zmg), 0.0)
    d2r = PI / 180.0
    vr = u * np.cos(elev * d2r) * np.sin(az * d2r) + v * np.cos(elev * d2r) * np.cos(az * d2r)
    vr = vr + (w - wt if use_tv else w) * np.sin(elev * d2r)
    return ref, vr


def np_operator(r, elm, ril, rjl, lon, lat, lev, grid, rotc):
    """one member; returns (yobs, qc)"""
    nobs = len(elm)
    y = np.full(nobs, -9.99e33)
    qc = np.zeros(nobs, dtype=np.int32)
    H, nlev = r.KHALO, r.nlev
    hgt = grid[..., 12]
    for n in range(nobs):
        if lev[n] > r.RADAR_ZMAX:
            qc[n] = 19
            continue
        if ril[n] < 1.0 or ril[n] > r.nlonh or rjl[n] < 1.0 or rjl[n] > r.nlath:
            qc[n] = 98
            continue
        i, j = int(np.ceil(ril[n])), int(np.ceil(rjl[n]))
        ai, aj = ril[n] - (i - 1), rjl[n] - (j - 1)
        cols = hgt[:, i - 2:i, j - 2:j]                      # 1-based (i-1, i) x (j-1, j)
        ks = H + 1
        for a in range(2):
            for b in range(2):
                ok = np.nonzero((cols[H:H + nlev, a, b] > -300.0) & (cols[H:H + nlev, a, b] < 10000.0))[0]
                kf = H + 1 + (ok[0] if len(ok) else nlev)
                ks = max(ks, kf)
        z = cols[:, 0, 0] * (1 - ai) * (1 - aj) + cols[:, 1, 0] * ai * (1 - aj) + cols[:, 0, 1] * (1 - ai) * aj + cols[:, 1, 1] * ai * aj
        if lev[n] > z[H + nlev - 1]:
            qc[n] = 20
            continue
        if lev[n] < z[ks - 1]:
            qc[n] = 21
            continue
        k = ks + 1
        while k <= nlev + H and not z[k - 1] > lev[n]:
            k += 1
        k = min(k, nlev + H)
        rk = (k - 1) + (lev[n] - z[k - 2]) / (z[k - 1] - z[k - 2])
        kk = int(np.ceil(rk))
        ak = rk - (kk - 1)

        def it(v):
            c = grid[kk - 2:kk, i - 2:i, j - 2:j, v]
            return (c[0, 0, 0] * (1 - ak) * (1 - ai) * (1 - aj) + c[1, 0, 0] * ak * (1 - ai) * (1 - aj)
                    + c[0, 1, 0] * (1 - ak) * ai * (1 - aj) + c[1, 1, 0] * ak * ai * (1 - aj)
                    + c[0, 0, 1] * (1 - ak) * (1 - ai) * aj + c[1, 0, 1] * ak * (1 - ai) * aj
                    + c[0, 1, 1] * (1 - ak) * ai * aj + c[1, 1, 1] * ak * ai * aj)
        u0, v0, w0, t0, p0 = it(0), it(1), it(2), it(3), it(4)
        qr, qs, qg = it(7), it(9), it(10)
        u1 = u0 * rotc[0, n] - v0 * rotc[1, n]
        v1 = u0 * rotc[1, n] + v0 * rotc[0, n]
        dlon, dlat = lon[n] - r.radar_lon, lat[n] - r.radar_lat
        if dlon == 0.0 and dlat == 0.0:
            qc[n] = 98
            continue
        az = 180.0 / PI * np.arctan2(dlon * np.cos(r.radar_lat * PI / 180.0), dlat)
        if az < 0:
            az += 360.0
        l1, l2, b1, b2 = lon[n] * PI / 180.0, r.radar_lon * PI / 180.0, lat[n] * PI / 180.0, r.radar_lat * PI / 180.0
        cosd = min(1.0, max(-1.0, np.sin(b1) * np.sin(b2) + np.cos(b1) * np.cos(b2) * np.cos(l2 - l1)))
        elev = 180.0 / PI * np.arctan2(lev[n] - r.radar_z, np.arccos(cosd) * RE)
        ref, vr = np_ref_vr(r.METHOD_REF_CALC, r.USE_TERMINAL_VELOCITY, *(np.float64(x) for x in (qr, qs, qg, u1, v1, w0, t0, p0, az, elev)))
        minref = 10.0 ** (r.MIN_RADAR_REF_DBZ / 10.0)
        if elm[n] in (4001, 4004):
            y[n] = r.MIN_RADAR_REF_DBZ + r.LOW_REF_SHIFT if ref < minref else 10.0 * np.log10(ref)
        elif elm[n] == 4002:
            y[n] = vr
        else:
            qc[n] = 90
    return y, qc


def _close(a, b, elm, tol_ref=1e-12, tol_vr=1e-7):
    tol = np.where(np.asarray(elm) == 4002, tol_vr, tol_ref)
    tol = tol[:, None] if a.ndim == 2 else tol
    ok = (a == b) | (np.abs(a - b) <= tol * np.maximum(np.abs(b), 1.0))
    return bool(ok.all()), float((np.abs(a - b) / np.maximum(np.abs(b), 1.0) / tol).max())


@pytest.mark.parametrize("method,use_tv", [(1, 1), (2, 1), (3, 0), (3, 1)])
def test_oracle_radar_matches_independent_numpy(oracle, method, use_tv):
    r, elm, ril, rjl, lon, lat, lev, grids, rotc = make_case(nobs=1500, nmem=2, method=method, use_tv=use_tv, seed=method)
    y, q = oracle.obsope_radar(r, elm, ril, rjl, lon, lat, lev, grids, rotc)
    seen = set()
    for m in range(2):
        y2, q2 = np_operator(r, elm, ril, rjl, lon, lat, lev, grids[m], rotc)
        assert np.array_equal(q[:, m], q2)
        ok, err = _close(y[:, m], y2, elm)
        assert ok, err
        seen |= set(q2.tolist())
    assert {0, 19, 20, 21, 90, 98} <= seen          # every QC exit of the branch is exercised


@pytest.mark.gpu
@pytest.mark.parametrize("method,use_tv", [(1, 1), (2, 1), (3, 0), (3, 1)])
def test_gpu_radar_operator_matches_oracle(oracle, method, use_tv):
    import torch
    import scale_letkf_b200 as sl
    r, elm, ril, rjl, lon, lat, lev, grids, rotc = make_case(nobs=20000, nmem=5, method=method, use_tv=use_tv, seed=10 + method)
    y0, q0 = oracle.obsope_radar(r, elm, ril, rjl, lon, lat, lev, grids, rotc)
    e = sl.LETKF(sl.resolve_config(sl.default_config(MEMBER=5, nlon=8, nlat=8, nlev=2)), device=0)
    y1, q1 = e.obsope_radar(r, elm, ril, rjl, lon, lat, lev, grids, rotc)                      # host buffers
    assert np.array_equal(q1, q0)
    ok, err = _close(y1, y0, elm)
    assert ok, err
    dg = [torch.from_numpy(np.ascontiguousarray(g.ravel(order="F"))).cuda() for g in grids]   # device-resident members
    y2, q2 = e.obsope_radar(r, elm, ril, rjl, lon, lat, lev, dg, rotc)
    assert np.array_equal(q2.cpu().numpy(), q1) and np.array_equal(y2.cpu().numpy(), y1)
    e.close()
