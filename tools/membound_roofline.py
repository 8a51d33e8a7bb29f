#!/usr/bin/env python
"""Roofline of the memory-bound kernels either side of the solver (north_star: "achieved HBM GB/s for the update and
search kernels"): ensmean_grd, enssprd_grd, state_trans, the one-pass member<->grid transposes (np = 1: the peer is the
rank itself), obs_departure_qc, the device observation chain (set_obs_device) and the stand-alone obs_local search.

Every kernel is called through the C ABI (scale_letkf_b200.LETKF) on device-resident arrays of the C2 shape
(256 x 256 x 60, 50 members) -- --small for a quick run -- timed with CUDA events after warm-up, with a buffer larger than
the 126 MB L2 written between repetitions (L2 flush).  `achieved` = ALGORITHMIC bytes (each value that must be read or
written, once) / time; `peak` = MEASURED_PEAKS.json hbm_gbs.  Under `ncu --metrics dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum` the same script gives the DRAM traffic per launch (profiles/).

    python tools/membound_roofline.py [--small] [--reps 5] > profiles/r02_membound_roofline.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    from scale_letkf_b200.transpose import EnsTransposeP2P
    import bench

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peaks, src = bench.measured_peaks()
    peak = float(peaks.get("hbm_gbs"))
    name = "c2small" if args.small else "c2"
    cfg, obs, rig1, rjg1, hgt1, ens = bench.make_workload(name, device=dev)
    k, nv, nlev, nij = cfg.MEMBER, cfg.nv3d, cfg.nlev, len(rig1)
    nens = k + 1
    eng = sl.LETKF(cfg, device=0)
    eng.set_letkf_obs(obs)
    eng.set_common_mpi_grid(rig1, rjg1, hgt1)
    v3d = ens.state(rig1, rjg1)                      # (nv3d, nens, nlev, nij) = Fortran (nij, nlev, nens, nv3d)
    flush = torch.empty(160 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 160 MB > 126 MB L2

    def timed(fn, reps=args.reps):
        fn()
        ms = []
        for _ in range(reps):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    rows = []

    def rec(kernel, what, nbytes, ms, bound="hbm", note=None):
        gbs = nbytes / (ms * 1e-3) * 1e-9
        r = {"kernel": kernel, "role": what, "algorithmic_bytes": int(nbytes), "ms": round(ms, 4), "achieved_gbs": round(gbs, 1),
             "peak_gbs": peak, "frac": round(gbs / peak, 4), "bound": bound}
        if note:
            r["note"] = note
        rows.append(r)

    def guard(what, fn):   # one failing section must not lose the others
        try:
            fn()
        except Exception as e:
            rows.append({"kernel": what, "error": repr(e)[:300]})
            torch.cuda.empty_cache()

    def _means():
        val = v3d.numel() // nens                        # values of one member slot
        # ensmean_grd: reads k member slots, writes the mean slot
        rec("ensmean_kernel", "ensmean_grd (common_scale.f90:1513)", (k + 1) * val * 8, timed(lambda: eng.ensmean_grd(v3d)))
        # enssprd_grd: reads k members + the mean, writes one spread field
        rec("enssprd_kernel", "enssprd_grd (common_scale.f90:1557)", (k + 2) * val * 8, timed(lambda: eng.enssprd_grd(v3d)))
    guard("means", _means)

    def _additive():
        val = v3d.numel() // nens
        an = torch.zeros_like(v3d)
        # additive inflation: the additive ensemble read (mean + update: the second read of a column comes from L2 at best),
        # the analysis members read and written
        rec("additive_inflation_kernel", "additive inflation block of das_letkf (letkf_tools.f90:869-925)", 3 * k * val * 8,
            timed(lambda: eng.additive_inflation(v3d, an, 0.1)))
        del an
    guard("additive", _additive)
    def _transposes():
        # one-pass transposes, np = 1 (local peer): every value read once and written once, state_trans fused
        thermo = eng.thermo_defaults()
        p2p = EnsTransposeP2P(eng, 1, 0, thermo=thermo)
        gsz = nlev * cfg.nlon * cfg.nlat * nv
        nm = min(k, 8)                                   # a block of members is enough to time the kernels
        grids = [torch.empty(gsz, dtype=torch.float64, device=dev) for _ in range(nm)]
        outg = [torch.empty(gsz, dtype=torch.float64, device=dev) for _ in range(nm)]
        p2p.write_ens(v3d, grids, nm, nens)              # physically sensible restart variables for state_trans
        rec("scatter_grd_p2p_kernel", "scatter_grd_mpi_alltoall + state_trans, one pass (common_mpi_scale.f90:1279)", 2 * nm * gsz * 8,
            timed(lambda: p2p.read_ens(grids, v3d, nm, nens)), note="%d members, np = 1" % nm)
        rec("gather_grd_p2p_kernel", "gather_grd_mpi_alltoall + state_trans_inv, one pass (common_mpi_scale.f90:1340)", 2 * nm * gsz * 8,
            timed(lambda: p2p.write_ens(v3d, outg, nm, nens)), note="%d members, np = 1" % nm)
        # stand-alone state_trans on one member-major grid (in place: read + write, +5 derived reads)
        g0 = grids[0].clone()
        rec("state_trans_kernel", "state_trans (common_scale.f90:1181), in place on one member", 2 * gsz * 8, timed(lambda: eng.state_trans(g0, thermo)))
    guard("transposes", _transposes)
    def _obs_chain():
        # observation chain on the device: departure + QC, then filter / sort / gather
        od = {kf: torch.as_tensor(np.ascontiguousarray(obs[kf], dtype=np.int32 if kf in ("elm", "typ") else np.float64), device=dev)
              for kf in ("elm", "typ", "ri", "rj", "lev", "dat", "err", "val", "ensval")}
        nobs, nensobs = od["ensval"].shape
        hx = od["ensval"] + od["dat"][:, None]           # H(x_m) again (the generator stored perturbation-like rows)
        qc = torch.zeros(nobs, dtype=torch.int32, device=dev)

        def dep():
            e = hx.clone()
            q = qc.clone()
            eng.obs_departure_qc_device(od["elm"], od["dat"], od["err"], q, e)
        t_clone = timed(lambda: (hx.clone(), qc.clone()))
        rec("obs_departure_qc_kernel", "departure + QC of set_letkf_obs (letkf_obs.f90:355-560)", 2 * nobs * nensobs * 8,
            max(timed(dep) - t_clone, 1e-4), note="clone of the input subtracted; %d observations x %d" % (nobs, nensobs))
        rec("set_obs_device (obs_prepare/compact/bucket_*/obs_gather kernels)", "qc filter + combined types + bucket sort + row gather "
            "(letkf_obs.f90:308-342, 747-805)", nobs * (nensobs * 8 + (k + 6) * 8 + 6 * 8), timed(lambda: eng.set_letkf_obs_device(od, None)),
            note="whole call incl. two small D2H synchronisations; latency-bound for %d observations" % nobs)
    guard("obs_chain", _obs_chain)
    # stand-alone search at the grid points of one level (latency / L2 bound: reported for completeness)
    try:
        npt = min(nij, 65536)
        lev = np.ascontiguousarray(hgt1[:npt, nlev // 2])
        plev = np.full(npt, 50000.0)
        maxl = 1024
        import time as _t
        eng.obs_local(rig1[:npt], rjg1[:npt], plev, lev, 1, maxl)
        t0 = _t.perf_counter()
        eng.obs_local(rig1[:npt], rjg1[:npt], plev, lev, 1, maxl)
        t = (_t.perf_counter() - t0) * 1e3
        rows.append({"kernel": "search_kernel (letkf_b200_obs_local, host buffers in and out)",
                     "role": "obs_local at %d points of one level (letkf_tools.f90:1325)" % npt, "ms": round(t, 3),
                     "points_per_s": round(npt / (t * 1e-3)),
                     "bound": "L2 latency (32 B geometry record per candidate) + the D2H of the index lists; inside das_letkf the same "
                              "search runs as presearch_kernel, overlapped with the solver"})
    except Exception as e:   # the stand-alone entry has a host-array signature on some builds
        rows.append({"kernel": "search_kernel (obs_local)", "note": "not timed here: " + repr(e)[:160]})
    print(json.dumps({"workload": name, "grid": [cfg.nlon, cfg.nlat, nlev], "k": k, "hbm_peak_gbs": peak, "peak_source": src + " MEASURED_PEAKS.json",
                      "l2": "160 MB buffer rewritten before every timed call", "rows": rows}, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
