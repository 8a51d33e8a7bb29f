"""Shared case builders and comparison helpers for the parity tests."""
import numpy as np

from scale_letkf_b200 import synth

TOL = 1e-10   # BASELINE.json north_star: analysis ensemble and weights within 1e-10 relative


def relerr(a, b, axis=None):
    """max |a-b| / max(|b|) (per trailing-variable scale when axis is given)."""
    if axis is None:
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    sc = np.maximum(np.abs(b).max(axis=axis, keepdims=True), 1e-300)
    return float((np.abs(a - b) / sc).max())


def sonde_case(member=8, nlon=24, nlat=24, nlev=6, nsonde=12, nsfc=40, hloc=60.0e3, det=False, seed=40, **kw):
    cfg = synth.config_c2(nlon=nlon, nlat=nlat, nlev=nlev, member=member, **kw)
    cfg.DET_RUN = 1 if det else 0
    for t in range(24):
        cfg.HORI_LOCAL[t] = hloc
    rig1, rjg1, hgt1 = synth.make_grid(cfg, topo_amp=300.0)
    obs = synth.make_sonde_obs(cfg, nsonde, nsfc, nlevobs=8, seed_no=seed)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=seed + 1)
    return cfg, rig1, rjg1, hgt1, obs, gues


def radar_case(member=8, nlon=40, nlat=40, nlev=8, max_nobs=30, det=False, seed=42, radius=7.0e3, **kw):
    cfg = synth.config_c3(nlon=nlon, nlat=nlat, nlev=nlev, member=member, max_nobs=max_nobs, **kw)
    cfg.DET_RUN = 1 if det else 0
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    obs = synth.make_radar_obs(cfg, radius_m=radius, zmin=500.0, zmax=6000.0, dz=1000.0, seed_no=seed)
    gues = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=seed + 1)
    return cfg, rig1, rjg1, hgt1, obs, gues


def host_logp(cfg, gues):
    return np.asfortranarray(np.log(gues[:, :, cfg.MEMBER, cfg.iv3d_p - 1]))


def sample_points(cfg, rig1, rjg1, hgt1, gues, stride=7):
    nij1, nlev = hgt1.shape
    k = cfg.MEMBER
    ri, rj, rlev, rz = [], [], [], []
    for il in range(nlev):
        for ij in range(0, nij1, stride):
            ri.append(rig1[ij]); rj.append(rjg1[ij])
            rlev.append(gues[ij, il, k, cfg.iv3d_p - 1]); rz.append(hgt1[ij, il])
    return tuple(np.array(x) for x in (ri, rj, rlev, rz))


def truth_invsqrt(A):
    """A^-1/2 of a symmetric positive definite matrix to long-double accuracy: double eigh, then Newton
    corrections with the Sylvester equation solved in the (double) eigenbasis, all products in 80-bit."""
    ld = np.longdouble
    Al = A.astype(ld)
    lam, V = np.linalg.eigh(A.astype(np.float64))
    Vl, s = V.astype(ld), np.sqrt(lam.astype(ld))
    Z = (Vl / s) @ Vl.T
    I = np.eye(A.shape[0], dtype=ld)
    for _ in range(4):
        R = I - Z @ Al @ Z
        Z = Z + Vl @ ((Vl.T @ R @ Vl) / (s[:, None] + s[None, :])) @ Vl.T
        Z = (Z + Z.T) / 2
    return Z


def truth_analysis_point(cfg, y, rd, dep, dx, xm, infl=1.0):
    """Long-double LETKF analysis of one grid point without relaxation or with RTPS (cfg.RELAX_ALPHA_SPREAD):
    y (p, k) local obs-space perturbations, rd (p) rdiag, dep (p), dx (k, nv) state perturbations, xm (nv) mean.
    Returns xa (k, nv) and the RTPS factors (nv).  (common_letkf.f90:111-226, letkf_tools.f90:457-486, 1971-2002)"""
    ld = np.longdouble
    k = y.shape[1]
    yl, w = y.astype(ld), 1.0 / rd.astype(ld)
    A = yl.T @ (yl * w[:, None]) + (ld(k - 1) / ld(infl)) * np.eye(k, dtype=ld)
    Z = truth_invsqrt(A)
    pa = Z @ Z
    W = np.sqrt(ld(k - 1)) * Z
    wm = pa @ (yl.T @ (dep.astype(ld) * w))
    xa = np.empty(dx.shape, dtype=ld)
    fac = np.ones(dx.shape[1], dtype=ld)
    for n in range(dx.shape[1]):
        d = dx[:, n].astype(ld)
        f = ld(1)
        if cfg.RELAX_ALPHA_SPREAD != 0.0:
            vg, va = d @ d, d @ pa @ d
            if vg > 0 and va > 0:
                a = ld(cfg.RELAX_ALPHA_SPREAD)
                f = a * np.sqrt(vg / (va * (k - 1))) - a + 1
        fac[n] = f
        xa[:, n] = ld(xm[n]) + d @ (W * f + wm[:, None])
    return xa, fac
