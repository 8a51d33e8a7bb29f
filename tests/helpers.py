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
