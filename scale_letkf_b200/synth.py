"""Seeded synthetic inputs for the BASELINE.json configs.

Everything the reference would read from NetCDF / observation files is replaced by
numpy `PCG64` draws with fixed seeds (20260101 + config number).  The same arrays feed
the CPU oracle and the CUDA library, so parity tests compare like with like.

Layouts follow the reference: state `gues3d(nij1, nlev, nens, nv3d)` Fortran order
(scale/letkf/letkf_tools.f90:54-57), observation `ensval(nensobs, nobs)` member fastest
(scale/common/common_obs_scale.f90:130), grid coordinates `rig1 = i + IHALO`
(scale/common/common_mpi_scale.f90:303-308).
"""
import numpy as np

from . import capi
from .config import default_config, resolve_config

SEED0 = 20260101


def rng(config_no, extra=0):
    return np.random.Generator(np.random.PCG64(SEED0 + config_no + 1000 * extra))


# ----------------------------------------------------------------------------- C1
def make_core_batch(ne=20, npts=10000, nobs=100, seed_no=1, det=False, infl=1.0):
    """BASELINE config 1: independent letkf_core problems, p ~ U{0..nobs} (p = 0 and p = 1
    forced for the first two points), rows of hdxb have zero member-mean
    (scale/letkf/letkf_obs.f90:474-489), rloc = exp(-ndist/2), rdiag = err^2/rloc."""
    g = rng(seed_no)
    nobsl = g.integers(0, nobs + 1, size=npts).astype(np.int32)
    if npts >= 2:
        nobsl[0], nobsl[1] = 0, 1
    hdxb = g.standard_normal((npts, ne, nobs))
    hdxb -= hdxb.mean(axis=1, keepdims=True)
    ndist = g.uniform(0.0, 13.33, size=(npts, nobs))
    rloc = np.exp(-0.5 * ndist)
    rdiag = 1.0 / rloc
    dep = 2.0 * g.standard_normal((npts, nobs))
    depd = 2.0 * g.standard_normal((npts, nobs)) if det else None
    parm_infl = np.full(npts, infl)
    return dict(ne=ne, nobs=nobs, npts=npts, nobsl=nobsl, hdxb=hdxb, rdiag=rdiag, rloc=rloc,
                dep=dep, depd=depd, parm_infl=parm_infl)


def make_raw_obs(member=12, nobs=600, det=False, seed_no=7):
    """Raw observation-space ensemble for the departure/QC step (scale/letkf/letkf_obs.f90:355-560):
    every element family of the gross-error `select case`, radar reflectivities scattered around
    RADAR_REF_THRES_DBZ (so that the member-count rules fire), undef radar data, pre-rejected
    observations (qc > 0) and gross outliers.  ensval (nobs, nensobs) holds H(x_m), not perturbations."""
    g = rng(seed_no, 9)
    elms = np.array([capi.ID_U, capi.ID_V, capi.ID_T, capi.ID_Q, capi.ID_PS, capi.ID_RAIN, capi.ID_RADAR_REF,
                     capi.ID_RADAR_REF_ZERO, capi.ID_RADAR_VR, capi.ID_RADAR_PRH, 99991, 99992, 99993], dtype=np.int32)
    elm = elms[g.integers(0, len(elms), size=nobs)]
    nens = member + (1 if det else 0)
    err = g.uniform(0.5, 3.0, size=nobs)
    radar = (elm == capi.ID_RADAR_REF) | (elm == capi.ID_RADAR_REF_ZERO)
    base = np.where(radar, g.uniform(5.0, 30.0, size=nobs), g.normal(0.0, 10.0, size=nobs))
    spread = np.where(radar, 8.0, 1.5)
    ens = base[:, None] + spread[:, None] * g.standard_normal((nobs, nens))
    dat = base + err * g.standard_normal(nobs) * 2.0
    out = g.uniform(size=nobs) < 0.08
    dat = np.where(out, dat + 12.0 * err * np.sign(g.standard_normal(nobs)), dat)   # gross errors
    dat = np.where(radar & (g.uniform(size=nobs) < 0.05), capi.UNDEF, dat)
    qc = np.where(g.uniform(size=nobs) < 0.1, g.choice([10, 20, 98], size=nobs), 0).astype(np.int32)
    return dict(elm=elm, dat=dat, err=err, qc=qc, ensval=np.ascontiguousarray(ens))


# ----------------------------------------------------------------------------- grids
def z_levels(nlev, zbot=100.0, ztop=28000.0):
    """stretched model-level heights (m)"""
    s = np.linspace(0.0, 1.0, nlev)
    return zbot + (ztop - zbot) * (0.35 * s + 0.65 * s * s)


def pressure_of_z(z):
    return 1.0e5 * np.exp(-z / 7500.0)


def column_deal(nlon, nlat, nprocs_e=1, myrank_e=0):
    """cyclic column deal of grd_to_buf (common_mpi_scale.f90:1428-1440): local column i
    (1-based) of e-rank m is global j = (m-1) + np*(i-1), ilon = mod(j,nlon)+1."""
    j = np.arange(myrank_e, nlon * nlat, nprocs_e)
    ilon = j % nlon + 1
    ilat = (j - ilon + 1) // nlon + 1
    return ilon, ilat


def make_grid(cfg, nprocs_e=1, myrank_e=0, topo_amp=0.0, seed_no=0):
    ilon, ilat = column_deal(cfg.nlon, cfg.nlat, nprocs_e, myrank_e)
    rig1 = (ilon + cfg.IHALO).astype(np.float64)
    rjg1 = (ilat + cfg.JHALO).astype(np.float64)
    z = z_levels(cfg.nlev)
    topo = topo_amp * (0.5 + 0.5 * np.sin(2 * np.pi * ilon / cfg.nlon) * np.cos(2 * np.pi * ilat / cfg.nlat))
    decay = np.linspace(1.0, 0.0, cfg.nlev)
    hgt1 = np.asfortranarray(z[None, :] + topo[:, None] * decay[None, :])
    return rig1, rjg1, hgt1


_VAR_SCALE = np.array([3.0, 3.0, 0.3, 1.0, 80.0, 8e-4, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4])
_VAR_MEAN = np.array([10.0, 2.0, 0.0, 280.0, 0.0, 5e-3, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4])


def make_state(cfg, rig1, rjg1, hgt1, seed_no=0, xp=np, device=None, gen=None):
    """Background ensemble gues3d(nij1,nlev,nens,nv3d): smooth mean + N(0, sigma_v) member
    noise; variable iv3d_p is a hydrostatic pressure (Pa) so vertical localisation in ln p
    is realistic.  Slot MEMBER+1 (mean) and MEMBER+2 (DET_RUN) are filled like
    ensmean_grd would (the caller may recompute the mean through the library)."""
    nij1, nlev = hgt1.shape
    k = cfg.MEMBER
    nens = k + 2 if cfg.DET_RUN else k + 1
    nv3d = cfg.nv3d
    if xp is np:
        g = rng(seed_no, 7)
        gues = np.empty((nij1, nlev, nens, nv3d), order="F")
        pz = pressure_of_z(hgt1)
        for n in range(nv3d):
            sc, mu = _VAR_SCALE[n % 11], _VAR_MEAN[n % 11]
            base = mu + sc * np.sin(0.07 * rig1)[:, None] * np.cos(0.05 * rjg1)[:, None] * np.ones((1, nlev))
            if n + 1 == cfg.iv3d_p:
                base = pz + 50.0 * np.sin(0.03 * rig1)[:, None]
            noise = sc * g.standard_normal((nij1, nlev, k))
            gues[:, :, :k, n] = base[:, :, None] + noise
            if cfg.DET_RUN:
                gues[:, :, k + 1, n] = base + sc * g.standard_normal((nij1, nlev))
            m = gues[:, :, 0, n].copy()
            for mm in range(1, k):
                m += gues[:, :, mm, n]
            gues[:, :, k, n] = m / k
        return gues
    # torch path (device-resident synthetic state for the full-size bench)
    import torch
    rig = torch.as_tensor(rig1, device=device)
    rjg = torch.as_tensor(rjg1, device=device)
    hg = torch.as_tensor(np.ascontiguousarray(hgt1.T), device=device)   # (nlev, nij1)
    # memory order of Fortran (nij1,nlev,nens,nv3d) == C order (nv3d,nens,nlev,nij1)
    gues = torch.empty((nv3d, nens, nlev, nij1), dtype=torch.float64, device=device)
    pz = 1.0e5 * torch.exp(-hg / 7500.0)
    for n in range(nv3d):
        sc, mu = float(_VAR_SCALE[n % 11]), float(_VAR_MEAN[n % 11])
        base = mu + sc * (torch.sin(0.07 * rig) * torch.cos(0.05 * rjg))[None, :].expand(nlev, nij1)
        if n + 1 == cfg.iv3d_p:
            base = pz + 50.0 * torch.sin(0.03 * rig)[None, :]
        gues[n, :k] = torch.randn((k, nlev, nij1), dtype=torch.float64, device=device, generator=gen)
        gues[n, :k].mul_(sc).add_(base[None])
        if cfg.DET_RUN:
            gues[n, k + 1] = base + sc * torch.randn((nlev, nij1), dtype=torch.float64, device=device, generator=gen)
        gues[n, k] = gues[n, :k].mean(dim=0)
    return gues


def _finish_obs(g, elm, typ, ri, rj, lev, dat, err, k, det, ens_sigma):
    nobs = len(elm)
    nensobs = k + 1 if det else k
    ens = g.standard_normal((nobs, nensobs)) * np.asarray(ens_sigma)[:, None]
    ens[:, :k] -= ens[:, :k].mean(axis=1, keepdims=True)   # perturbation form, zero member mean
    val = 2.0 * np.asarray(ens_sigma) * g.standard_normal(nobs)
    perm = g.permutation(nobs)   # arrival order is not sorted in the reference either
    return dict(elm=np.asarray(elm, np.int32)[perm], typ=np.asarray(typ, np.int32)[perm],
                ri=np.asarray(ri)[perm], rj=np.asarray(rj)[perm], lev=np.asarray(lev)[perm],
                dat=np.asarray(dat)[perm], err=np.asarray(err)[perm], val=val[perm],
                ensval=np.ascontiguousarray(ens[perm]))


def make_sonde_obs(cfg, nsonde, nsfc, nlevobs=25, seed_no=2):
    """Config-2 style conventional obs: `nsonde` soundings at continuous random positions x
    `nlevobs` pressure levels x {U,V,T,Q} as ADPUPA(1), plus `nsfc` surface-pressure obs as
    ADPSFC(8)."""
    g = rng(seed_no, 3)
    k, det = cfg.MEMBER, bool(cfg.DET_RUN)
    lo_i, hi_i = cfg.IHALO + 0.5, cfg.nlon + cfg.IHALO + 0.5
    lo_j, hi_j = cfg.JHALO + 0.5, cfg.nlat + cfg.JHALO + 0.5
    sri = g.uniform(lo_i, hi_i, nsonde)
    srj = g.uniform(lo_j, hi_j, nsonde)
    plev = np.exp(np.linspace(np.log(1.0e5), np.log(5.0e3), nlevobs))
    elms = [capi.ID_U, capi.ID_V, capi.ID_T, capi.ID_Q]
    errs = {capi.ID_U: 2.0, capi.ID_V: 2.0, capi.ID_T: 1.0, capi.ID_Q: 1.0e-3}
    elm, typ, ri, rj, lev, dat, err, sig = [], [], [], [], [], [], [], []
    for s in range(nsonde):
        for p in plev * np.exp(g.uniform(-0.02, 0.02, nlevobs)):
            for e in elms:
                elm.append(e); typ.append(capi.TYP_ADPUPA); ri.append(sri[s]); rj.append(srj[s])
                lev.append(p); dat.append(0.0); err.append(errs[e]); sig.append(errs[e] * 1.2)
    for s in range(nsfc):
        elm.append(capi.ID_PS); typ.append(capi.TYP_ADPSFC)
        ri.append(g.uniform(lo_i, hi_i)); rj.append(g.uniform(lo_j, hi_j))
        lev.append(10.0); dat.append(1.0e5 + 1500.0 * g.standard_normal()); err.append(100.0)
        sig.append(120.0)
    return _finish_obs(g, elm, typ, ri, rj, lev, dat, err, k, det, sig)


def make_radar_obs(cfg, radius_m=60.0e3, zmin=500.0, zmax=11000.0, dz=500.0, mesh_m=500.0,
                   jitter_m=200.0, frac_ref=0.5, seed_no=3, center=None):
    """Config-3 style phased-array radar obs (type 22 PHARAD): positions on a `mesh_m` mesh
    with +-`jitter_m` uniform jitter (so that equal-distance ties have measure zero, SURVEY
    H2) inside a cylinder; each position is REF (4001) or RE0 (4004), and REF positions
    also carry a Doppler velocity VR (4002)."""
    g = rng(seed_no, 5)
    k, det = cfg.MEMBER, bool(cfg.DET_RUN)
    cx = (cfg.nlon * 0.5 + cfg.IHALO + 0.5) if center is None else center[0]
    cy = (cfg.nlat * 0.5 + cfg.JHALO + 0.5) if center is None else center[1]
    nx = int(radius_m / mesh_m)
    xs = np.arange(-nx, nx + 1) * mesh_m
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    m = (X * X + Y * Y) <= radius_m * radius_m
    X, Y = X[m], Y[m]
    zs = np.arange(zmin, zmax + 0.5 * dz, dz)
    elm, typ, ri, rj, lev, dat, err, sig = [], [], [], [], [], [], [], []
    for z in zs:
        n = X.size
        x = X + g.uniform(-jitter_m, jitter_m, n)
        y = Y + g.uniform(-jitter_m, jitter_m, n)
        zz = z + g.uniform(-0.2 * dz, 0.2 * dz, n)
        r_i = cx + x / cfg.DX
        r_j = cy + y / cfg.DY
        ok = (r_i > cfg.IHALO + 0.5) & (r_i < cfg.nlon + cfg.IHALO + 0.5) & \
             (r_j > cfg.JHALO + 0.5) & (r_j < cfg.nlat + cfg.JHALO + 0.5)
        r_i, r_j, zz = r_i[ok], r_j[ok], zz[ok]
        n = r_i.size
        is_ref = g.uniform(size=n) < frac_ref
        e = np.where(is_ref, capi.ID_RADAR_REF, capi.ID_RADAR_REF_ZERO)
        elm += list(e); typ += [capi.TYP_PHARAD] * n; ri += list(r_i); rj += list(r_j)
        lev += list(zz); dat += list(np.where(is_ref, 30.0, 5.0)); err += [5.0] * n; sig += [6.0] * n
        nv = int(is_ref.sum())
        elm += [capi.ID_RADAR_VR] * nv; typ += [capi.TYP_PHARAD] * nv
        ri += list(r_i[is_ref]); rj += list(r_j[is_ref]); lev += list(zz[is_ref])
        dat += [0.0] * nv; err += [3.0] * nv; sig += [3.5] * nv
    return _finish_obs(g, elm, typ, ri, rj, lev, dat, err, k, det, sig)


# ----------------------------------------------------------------------------- named configs
def config_c2(nlon=256, nlat=256, nlev=60, member=50, **kw):
    """BASELINE config 2: 15 km mesh, sonde/surface obs, R-localisation, RTPS 0.95."""
    c = default_config(MEMBER=member, nlon=nlon, nlat=nlat, nlev=nlev, DX=15000.0, DY=15000.0,
                       RELAX_ALPHA_SPREAD=0.95, INFL_MUL=1.0)
    c.HORI_LOCAL[0] = 200.0e3
    c.VERT_LOCAL[0] = 0.4
    c.MAX_NOBS_PER_GRID[0] = 0
    for key, v in kw.items():
        setattr(c, key, v)
    return resolve_config(c)


def config_c3(nlon=256, nlat=256, nlev=60, member=100, max_nobs=500, **kw):
    """BASELINE config 3 / 5: 500 m mesh, dense radar, settings of
    scale/run/config/BDA_d4_500m_9p_bf30/config.nml.letkf:34,44,52-61,84."""
    c = default_config(MEMBER=member, nlon=nlon, nlat=nlat, nlev=nlev, DX=500.0, DY=500.0,
                       RELAX_ALPHA_SPREAD=0.95, INFL_MUL=1.0, BOUNDARY_BUFFER_WIDTH=15.0e3,
                       RADAR_ZMAX=11.0e3)
    c.HORI_LOCAL[0] = 4.0e3
    c.VERT_LOCAL[0] = 0.3
    c.VERT_LOCAL[21] = 2.0e3
    c.HORI_LOCAL_RADAR_OBSNOREF = 2.0e3
    c.MAX_NOBS_PER_GRID[0] = max_nobs
    c.MAX_NOBS_PER_GRID_CRITERION = 1
    for key, v in kw.items():
        setattr(c, key, v)
    return resolve_config(c)


# ----------------------------------------------------------------------------- correlated ensemble + H(x)
class SmoothEnsemble:
    """Background ensemble and observation operator of the benchmark workloads ("gues3d =
    smooth random fields + member noise; H(x) = linear interpolation of the synthetic members to the obs
    location, so that ensval is consistent with gues3d").

    Every draw d (members 0..k-1, deterministic run k, truth k+1) of variable n is a FUNCTION of the global grid
    index, so any rank can evaluate any grid value without communication and 1/2/4/8 GPUs analyse the same
    problem:
        x_d(ri, rj, lev) = base_n + sigma_n [ rho S_d(ri, rj, z) + sqrt(1 - rho^2) W_d(ri, rj, lev) ]
    S_d = sum_j (alpha_dj cos th_j + beta_dj sin th_j) / sqrt(J), th_j = kx_j ri + ky_j rj + kz_j z with Gaussian
    wave numbers: a Gaussian random field of unit variance, horizontal correlation length `hlen` grid cells,
    vertical `vlen` metres (the localisation scales).  W_d is white noise from an integer hash of
    (ri, rj, lev, d, n).  Observations are y = H(x_truth) + err N(0,1), H = tri-linear interpolation of the
    gridded draws (letkf_obs.f90:474-490 semantics: ensval = H(x_m) - mean, val = y - mean, DET row =
    y - H(x_det)).  Local observations of one storm/sounding are therefore strongly correlated and
    lambda_max(Yr^T Y) / c0 reaches 10^2-10^4, unlike i.i.d. rows.  torch is used as the array library (CPU or CUDA)."""

    VAR_OF_ELM = {2819: 0, 2820: 1, 3073: 3, 3330: 5}   # U, V, T, Q -> iv3d - 1

    def __init__(self, cfg, seed_no=4, rho=0.9, hlen=None, vlen=3000.0, nmodes=None, topo_amp=0.0, device=None):
        import torch
        self.torch, self.cfg, self.dev = torch, cfg, device
        k = cfg.MEMBER
        self.k, self.ndraw, self.rho, self.topo_amp = k, k + 2, float(rho), float(topo_amp)
        J = nmodes or max(32, (k + 2) // 2 + 8)
        g = rng(seed_no, 13)
        hlen = hlen or cfg.HORI_LOCAL[0] / cfg.DX
        self.kx = torch.as_tensor(g.normal(0.0, 1.0 / hlen, J), device=device)
        self.ky = torch.as_tensor(g.normal(0.0, 1.0 / hlen, J), device=device)
        self.kz = torch.as_tensor(g.normal(0.0, 1.0 / vlen, J), device=device)
        self.coef = torch.as_tensor(g.standard_normal((cfg.nv3d, 2 * J, self.ndraw)) / np.sqrt(J), device=device)
        self.zlev = torch.as_tensor(z_levels(cfg.nlev), device=device)
        self.decay = torch.as_tensor(np.linspace(1.0, 0.0, cfg.nlev), device=device)
        self.sigma = [float(_VAR_SCALE[n % 11]) for n in range(cfg.nv3d)]
        self.mu = [float(_VAR_MEAN[n % 11]) for n in range(cfg.nv3d)]

    # -- pieces -------------------------------------------------------------------------------
    def hgt(self, ri, rj, lev):
        """model-level height at (integer) grid coordinates, same formula as make_grid"""
        t = self.torch
        c = self.cfg
        topo = self.topo_amp * (0.5 + 0.5 * t.sin(2 * np.pi * (ri - c.IHALO) / c.nlon) * t.cos(2 * np.pi * (rj - c.JHALO) / c.nlat))
        return self.zlev[lev] + topo * self.decay[lev]

    def _white(self, ri, rj, lev, n, draws):
        """N(0,1) white noise, (npts, ndraws), from an integer hash of (ri, rj, lev, draw, variable)"""
        t = self.torch
        base = ((lev.to(t.int64) * 2097169 + rj.to(t.int64)) * 4194319 + ri.to(t.int64))[:, None]
        x = base + (draws.to(t.int64)[None, :] + 4099 * n) * 1048583 * 8388617

        def mix(v):
            v = (v ^ (v >> 30)) * -4658895280553007687     # splitmix64 constants as signed int64
            v = (v ^ (v >> 27)) * -7723592293110705685
            return v ^ (v >> 31)

        def unif(v):
            return ((v >> 11) & ((1 << 53) - 1)).to(t.float64) * (2.0 ** -53) + 2.0 ** -54

        u1 = unif(mix(x))
        u2 = unif(mix(x + 0x632BE59BD9B4E019))
        return t.sqrt(-2.0 * t.log(u1)) * t.cos(2.0 * np.pi * u2)

    def base(self, n, ri, rj, z):
        t = self.torch
        if n + 1 == self.cfg.iv3d_p:
            return 1.0e5 * t.exp(-z / 7500.0) + 50.0 * t.sin(0.03 * ri)
        return self.mu[n] + self.sigma[n] * t.sin(0.07 * ri) * t.cos(0.05 * rj) + 0.0 * z

    def values(self, n, ri, rj, lev, draws=None, chunk=1 << 18):
        """x_d of variable n (0-based) at integer grid coordinates ri, rj (rig1-style, halo offset included) and
        level indices lev (all 1-D, same length) -> (npts, ndraws)"""
        t = self.torch
        draws = t.arange(self.ndraw, device=self.dev) if draws is None else draws
        out = t.empty((ri.numel(), draws.numel()), dtype=t.float64, device=self.dev)
        cf = self.coef[n][:, draws]
        for s in range(0, ri.numel(), chunk):
            a, b, l = ri[s:s + chunk].to(t.float64), rj[s:s + chunk].to(t.float64), lev[s:s + chunk]
            z = self.hgt(a, b, l)
            th = a[:, None] * self.kx[None, :] + b[:, None] * self.ky[None, :] + z[:, None] * self.kz[None, :]
            sm = t.cat([t.cos(th), t.sin(th)], dim=1) @ cf
            w = self._white(a, b, l, n, draws)
            out[s:s + chunk] = self.base(n, a, b, z)[:, None] + self.sigma[n] * (
                self.rho * sm + np.sqrt(1.0 - self.rho ** 2) * w)
        return out

    # -- state -----------------------------------------------------------------------------------
    def state(self, rig1, rjg1, as_numpy=False):
        """gues3d of the columns (rig1, rjg1): torch (nv3d, nens, nlev, nij1) [the memory order of the Fortran
        (nij1, nlev, nens, nv3d)] or, as_numpy, that Fortran-ordered numpy array.  Slot k = member mean summed in
        member order like ensmean_grd, slot k+1 = deterministic run when DET_RUN."""
        t = self.torch
        c = self.cfg
        k, nlev, nij1 = self.k, c.nlev, len(rig1)
        nens = k + 2 if c.DET_RUN else k + 1
        ri = t.as_tensor(np.asarray(rig1), device=self.dev).repeat(nlev)
        rj = t.as_tensor(np.asarray(rjg1), device=self.dev).repeat(nlev)
        lev = t.arange(nlev, device=self.dev).repeat_interleave(nij1)
        draws = t.arange(k + 1 if c.DET_RUN else k, device=self.dev)
        gues = t.empty((c.nv3d, nens, nlev, nij1), dtype=t.float64, device=self.dev)
        for n in range(c.nv3d):
            v = self.values(n, ri, rj, lev, draws)            # (nlev*nij1, ndraws)
            gues[n, :k] = v[:, :k].T.reshape(k, nlev, nij1)
            if c.DET_RUN:
                gues[n, k + 1] = v[:, k].reshape(nlev, nij1)
            m = gues[n, 0].clone()
            for mm in range(1, k):
                m += gues[n, mm]
            gues[n, k] = m / k
        if as_numpy:
            return np.asfortranarray(gues.cpu().numpy().T)
        return gues

    # -- observation operator ---------------------------------------------------------------------
    def _interp(self, n, ri, rj, rk, draws):
        """tri-linear interpolation of variable n at real coordinates (ri, rj) and fractional level rk"""
        t = self.torch
        i0, j0 = t.floor(ri), t.floor(rj)
        k0 = t.clamp(t.floor(rk), 0, self.cfg.nlev - 2) if self.cfg.nlev > 1 else t.zeros_like(rk)
        fi, fj, fk = ri - i0, rj - j0, t.clamp(rk - k0, 0.0, 1.0)
        out = None
        for di in (0, 1):
            for dj in (0, 1):
                for dk in ((0, 1) if self.cfg.nlev > 1 else (0,)):
                    wgt = (fi if di else 1 - fi) * (fj if dj else 1 - fj) * ((fk if dk else 1 - fk) if self.cfg.nlev > 1 else 1.0)
                    v = self.values(n, i0 + di, j0 + dj, (k0 + dk).to(t.int64), draws) * wgt[:, None]
                    out = v if out is None else out + v
        return out

    def hx(self, elm, ri, rj, lev, radar_xyz=None):
        """H(x_d) for all draws: (nobs, ndraw).  lev = pressure (Pa) for conventional obs, height (m) for radar."""
        t = self.torch
        c = self.cfg
        elm = np.asarray(elm)
        ri_t = t.as_tensor(np.asarray(ri, dtype=np.float64), device=self.dev)
        rj_t = t.as_tensor(np.asarray(rj, dtype=np.float64), device=self.dev)
        lev_t = t.as_tensor(np.asarray(lev, dtype=np.float64), device=self.dev)
        draws = t.arange(self.ndraw, device=self.dev)
        radar = np.isin(elm, (capi.ID_RADAR_REF, capi.ID_RADAR_REF_ZERO, capi.ID_RADAR_VR))
        zobs = t.where(t.as_tensor(radar, device=self.dev), lev_t, -7500.0 * t.log(t.clamp(lev_t, min=1.0) / 1.0e5))
        idx = t.arange(c.nlev, dtype=t.float64, device=self.dev)
        # fractional level index from the flat level heights (piecewise linear inverse)
        pos = t.clamp(t.searchsorted(self.zlev, zobs.contiguous()), 1, max(c.nlev - 1, 1))
        z0, z1 = self.zlev[pos - 1], self.zlev[t.clamp(pos, max=c.nlev - 1)]
        rk = idx[pos - 1] + t.clamp((zobs - z0) / t.clamp(z1 - z0, min=1e-9), 0.0, 1.0)
        out = t.empty((len(elm), self.ndraw), dtype=t.float64, device=self.dev)
        cx, cy = radar_xyz if radar_xyz is not None else (c.nlon * 0.5 + c.IHALO + 0.5, c.nlat * 0.5 + c.JHALO + 0.5)
        for e in np.unique(elm):
            sel = t.as_tensor(np.nonzero(elm == e)[0], device=self.dev)
            a, b, r = ri_t[sel], rj_t[sel], rk[sel]
            if int(e) in self.VAR_OF_ELM:
                v = self._interp(self.VAR_OF_ELM[int(e)], a, b, r, draws)
            elif e == capi.ID_PS:
                v = self._interp(c.iv3d_p - 1, a, b, t.zeros_like(r), draws)
            elif e in (capi.ID_RADAR_REF, capi.ID_RADAR_REF_ZERO):   # linearised reflectivity (dBZ) of rain + snow + graupel
                v = 25.0 + 4.0e4 * (self._interp(7, a, b, r, draws) + self._interp(9, a, b, r, draws)
                                    + self._interp(10, a, b, r, draws) - 3.0e-4)
            elif e == capi.ID_RADAR_VR:   # radial velocity seen from a radar at the domain centre, z = 0
                dx, dy, dz = (a - cx) * c.DX, (b - cy) * c.DY, zobs[sel]
                dist = t.sqrt(dx * dx + dy * dy + dz * dz).clamp(min=1.0)
                v = (self._interp(0, a, b, r, draws) * (dx / dist)[:, None] + self._interp(1, a, b, r, draws) * (dy / dist)[:, None]
                     + self._interp(2, a, b, r, draws) * (dz / dist)[:, None])
            else:
                raise ValueError(f"SmoothEnsemble.hx: element {e} not supported")
            out[sel] = v
        return out

    def attach(self, obs, seed_no=5):
        """replace the i.i.d. ensval / val / dat of an obs dict (make_sonde_obs / make_radar_obs positions) by
        H(x)-consistent ones"""
        t = self.torch
        k, det = self.k, bool(self.cfg.DET_RUN)
        h = self.hx(obs["elm"], obs["ri"], obs["rj"], obs["lev"])
        g = rng(seed_no, 17)
        err = t.as_tensor(np.asarray(obs["err"]), device=self.dev)
        y = h[:, k + 1] + err * t.as_tensor(g.standard_normal(len(err)), device=self.dev)
        mean = h[:, 0].clone()
        for m in range(1, k):
            mean += h[:, m]
        mean /= k
        ens = t.empty((len(err), k + 1 if det else k), dtype=t.float64, device=self.dev)
        ens[:, :k] = h[:, :k] - mean[:, None]
        if det:
            ens[:, k] = y - h[:, k]
        out = dict(obs)
        out["ensval"] = np.ascontiguousarray(ens.cpu().numpy())
        out["val"] = (y - mean).cpu().numpy()
        out["dat"] = y.cpu().numpy()
        return out
