#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory.

    python tests/golden/make_golden.py

PARITY UNPINNED BY THE REFERENCE: gylien/scale-letkf ships no golden vectors, KATs or fixtures
for the analysis path and its Fortran cannot be built in the build container (no Fortran
front-end, no MPI, SCALE-RM / NetCDF not vendored), so nothing here is an output of the
reference binary.  The fixtures pin the behaviour of this repo against two independent
restatements of the reference's algorithm:

  core_k20.npz      letkf_core (common/common_letkf.f90:52-257) on a C1-shaped batch (k = 20,
                    p in 0..40).  `trans/transm/pao/transmd` come from an INDEPENDENT numpy/LAPACK
                    evaluation written straight from the reference's formulas (eigh of
                    hdxb^T R^-1 hdxb + (k-1)/rho I, :140-226) -- NOT from oracle/ -- so the C++
                    oracle (EISPACK tred2/tql2 restatement) and the CUDA path are both checked
                    against it.
  select_radar.npz  obs_local (scale/letkf/letkf_tools.f90:1325-1759) selection sets of a small
                    radar case with MAX_NOBS_PER_GRID(22) = 12, criterion 1: per point the sorted
                    ORIGINAL observation ids, produced by an O(nobs) brute-force numpy scan that
                    restates obs_local_cal (:1793-1906) -- again independent of oracle/.
  das_small.npz     das_letkf (letkf_tools.f90:50-932) analysis of a small sonde case, from the
                    C++ oracle (regression pin for the oracle and target for the CUDA path).

  obsqc.npz         departure + QC half of set_letkf_obs (scale/letkf/letkf_obs.f90:355-560): qc codes,
                    departures and H(x) perturbations of a 600-observation mixed-type case, from the
                    vectorised numpy restatement in tests/test_obs_qc.py (independent of oracle/).

Inputs are regenerated from seeds by scale_letkf_b200.synth in the tests; only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from scale_letkf_b200 import synth, config  # noqa: E402

CORE = dict(ne=20, npts=48, nobs=40, seed_no=901, det=True, infl=1.03)
SELECT = dict(member=4, nlon=24, nlat=24, nlev=5, max_nobs=12, seed=902, radius=5.0e3)
DAS = dict(member=10, nlon=12, nlat=12, nlev=4, nsonde=10, nsfc=30, seed=903)


def letkf_core_numpy(hdxb, rdiag, dep, depd, infl, nobsl):
    """common_letkf.f90:111-226 with rdiag_wloc = .true., via LAPACK eigh."""
    ne = hdxb.shape[0]
    if nobsl == 0:   # :89-107
        return np.sqrt(infl) * np.eye(ne), np.zeros(ne), infl / (ne - 1) * np.eye(ne), np.zeros(ne)
    Y = hdxb[:, :nobsl].T                       # (p, k)
    Yr = Y / rdiag[:nobsl, None]                # hdxb_rinv (:111-116)
    A = Yr.T @ Y + (ne - 1) / infl * np.eye(ne)  # :127-143
    lam, V = np.linalg.eigh(A)                  # mtx_eigen (:149)
    pa = (V / lam) @ V.T                        # :153-159
    trans = (V * np.sqrt((ne - 1) / lam)) @ V.T  # :201-208
    transm = pa @ (Yr.T @ dep[:nobsl])          # :169-195
    transmd = pa @ (Yr.T @ depd[:nobsl])
    return trans, transm, pa, transmd


def make_core():
    c = synth.make_core_batch(**CORE)
    ne, npts = c["ne"], len(c["nobsl"])
    out = dict(trans=np.zeros((npts, ne, ne)), transm=np.zeros((npts, ne)), pao=np.zeros((npts, ne, ne)),
               transmd=np.zeros((npts, ne)))
    for i in range(npts):
        t, m, p, md = letkf_core_numpy(c["hdxb"][i], c["rdiag"][i], c["dep"][i], c["depd"][i],
                                       float(c["parm_infl"][i]), int(c["nobsl"][i]))
        out["trans"][i], out["transm"][i], out["pao"][i], out["transmd"][i] = t.T, m, p.T, md
    np.savez_compressed(os.path.join(HERE, "core_k20.npz"), nobsl=c["nobsl"], **out)


def select_bruteforce(cfg, obs, pts, nvar):
    """obs_local semantics by brute force: every observation passing obs_local_cal
    (letkf_tools.f90:1793-1906), then per combined-type budget the N nearest (criterion 1)."""
    dzf, dzf2 = cfg.dist_zero_fac, cfg.dist_zero_fac_square
    ri, rj, rlev, rz = pts
    elm, typ = obs["elm"], obs["typ"]
    res = []
    for p in range(len(ri)):
        sel = []
        # budget groups: REF + RE0 of type 22 merged, VR separate (letkf_tools.f90:167-192)
        for grp in ((4001, 4004), (4002,)):
            cand = []
            for e in grp:
                hl = cfg.HORI_LOCAL_RADAR_OBSNOREF if e == 4004 else cfg.HORI_LOCAL_RADAR_VR if e == 4002 \
                    else cfg.HORI_LOCAL[21]
                vl = cfg.VERT_LOCAL_RADAR_VR if e == 4002 else cfg.VERT_LOCAL[21]
                for n in np.nonzero((elm == e) & (typ == 22))[0]:
                    nd_v = abs(obs["lev"][n] - rz[p]) / vl                       # :1857-1858
                    if nd_v > dzf:
                        continue
                    rdx = (ri[p] - obs["ri"][n]) * cfg.DX
                    rdy = (rj[p] - obs["rj"][n]) * cfg.DY
                    nd_h = np.sqrt(rdx * rdx + rdy * rdy) / hl                   # :1876-1881
                    if nd_h > dzf:
                        continue
                    nd = nd_h * nd_h + nd_v * nd_v
                    if nd > dzf2:
                        continue
                    cand.append((nd, int(n)))
            cand.sort()
            N = cfg.MAX_NOBS_PER_GRID[21]
            if N <= 0:                      # no obs-number limit (:1438-1476): everything inside the cut-off
                sel += [n for _, n in cand]
                continue
            if len(cand) > N:
                assert cand[N - 1][0] < cand[N][0], "tie at the N-th distance: change the seed"
            sel += [n for _, n in cand[:N]]
        res.append(sorted(sel))
    return res


def make_select():
    from helpers import radar_case, sample_points
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(**SELECT)
    pts = sample_points(cfg, rig1, rjg1, hgt1, gues, stride=11)
    res = select_bruteforce(cfg, obs, pts, 1)
    nobsl = np.array([len(r) for r in res], dtype=np.int32)
    ids = np.full((len(res), max(int(nobsl.max()), 1)), -1, dtype=np.int32)
    for i, r in enumerate(res):
        ids[i, :len(r)] = r
    assert nobsl.max() == 24 and (nobsl == 0).any()
    np.savez_compressed(os.path.join(HERE, "select_radar.npz"), nobsl=nobsl, ids=ids)


def make_das():
    from helpers import sonde_case
    from oracle import oracle_py
    oracle_py.build()
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(**DAS)
    cfg.RELAX_ALPHA_SPREAD = 0.95
    o = oracle_py.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    r = o.das_letkf(gues.copy(order="F"), want_nobsl=True)
    k = cfg.MEMBER
    np.savez_compressed(os.path.join(HERE, "das_small.npz"), anal3d=r["anal3d"][:, :, :k, :], nobsl=r["nobsl"])


if __name__ == "__main__":
    make_core()
    make_select()
    make_das()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


def make_obsqc():
    from test_obs_qc import numpy_departure_qc, qc_defaults, CASE, GOLD
    o = synth.make_raw_obs(**CASE)
    qc, val, ens = numpy_departure_qc(qc_defaults(), CASE["member"], CASE["det"], o["elm"], o["dat"], o["err"],
                                      o["qc"], o["ensval"])
    np.savez_compressed(GOLD, qc=qc, val=val, ensval=ens)


if __name__ == "__main__" and "--obsqc" in sys.argv:
    make_obsqc()
