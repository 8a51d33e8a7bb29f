#!/usr/bin/env python
"""bench.py -- analysed grid points/s of the LETKF analysis hot path (das_letkf twin).

Workload (BASELINE.json configs[1], "C2"): SCALE-LETKF regional 256x256x60 grid, 50 members,
synthetic sonde/surface observations, R-localisation, RTPS 0.95.  One "step" = one das_letkf
over the whole domain: local-observation search + weight solve + relaxation + ensemble update
of every (ij, lev) point.  With N ranks the horizontal domain is dealt to the ranks column by
column exactly like the reference's e-rank deal (scale/common/common_mpi_scale.f90:264-283,
1428-1440), every rank holds all observations, and there is no data-path collective
(strong scaling: total work fixed).

    python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
    python bench.py --impl reference ...                     CPU restatement of the reference
                                                             (oracle/, all host threads), rank 0

Prints ONE JSON line (see the task contract): value = device-resident throughput,
e2e = the same call through the C ABI with host buffers (H2D + D2H inside the timed region),
roofline for the dominant kernel (das_kernel), cpu_baseline = oracle on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "analysed grid points/s"
UNIT = "points/s"

WORKLOADS = {
    # name: (config factory kwargs, obs kwargs)
    "c2": dict(kind="sonde", nlon=256, nlat=256, nlev=60, member=50, nsonde=50, nsfc=500,
               desc="C2: 256x256x60, 50 members, sonde+surface obs, R-localisation, RTPS 0.95"),
    "c2small": dict(kind="sonde", nlon=64, nlat=64, nlev=20, member=50, nsonde=50, nsfc=500,
                    desc="C2 shape on a 64x64x20 grid (smoke-size)"),
    "c3": dict(kind="radar", nlon=256, nlat=256, nlev=60, member=100, max_nobs=500,
               desc="C3: 256x256x60 500 m mesh, 100 members, dense radar, MAX_NOBS_PER_GRID(22)=500"),
    "c3small": dict(kind="radar", nlon=96, nlat=96, nlev=20, member=100, max_nobs=500,
                    desc="C3 shape on a 96x96x20 grid"),
    "c4": dict(kind="radar", nlon=128, nlat=128, nlev=40, member=1000, max_nobs=100,
               desc="C4: 128x128x40, 1000 members, dense radar, MAX_NOBS_PER_GRID(22)=100 (tiled path)"),
    "c4small": dict(kind="radar", nlon=48, nlat=48, nlev=10, member=1000, max_nobs=100,
                    desc="C4 shape on a 48x48x10 grid (tiled path)"),
    "c5": dict(kind="radar", nlon=400, nlat=400, nlev=60, member=100, max_nobs=100,
               desc="C5: 400x400x60 500 m mesh, 100 members, phased-array radar, MAX_NOBS_PER_GRID(22)=100"),
}


def make_workload(name, nprocs_e=1, myrank_e=0, device=None, synth_kind="hx"):
    """-> cfg, obs, rig1, rjg1, hgt1, ens.  synth_kind "hx" (default): members are smooth random fields + gridpoint
    noise and ensval = H(x_m) by tri-linear interpolation of those members, one global
    state whatever the rank count; "iid": the round-1 generator (independent rows, benign conditioning); "grid": no
    observations (grid of a column sample only)."""
    from scale_letkf_b200 import synth
    w = WORKLOADS[name]
    if w["kind"] == "sonde":
        cfg = synth.config_c2(nlon=w["nlon"], nlat=w["nlat"], nlev=w["nlev"], member=w["member"])
        obs = None if synth_kind == "grid" else synth.make_sonde_obs(cfg, w["nsonde"], w["nsfc"], nlevobs=25, seed_no=2)
        hlen, vlen = cfg.HORI_LOCAL[0] / cfg.DX, 3000.0
    else:
        cfg = synth.config_c3(nlon=w["nlon"], nlat=w["nlat"], nlev=w["nlev"], member=w["member"],
                              max_nobs=w["max_nobs"])
        rad = min(60.0e3, 0.47 * w["nlon"] * 500.0)
        obs = None if synth_kind == "grid" else synth.make_radar_obs(cfg, radius_m=rad, zmin=500.0, zmax=11000.0, dz=500.0, seed_no=3)
        hlen, vlen = cfg.HORI_LOCAL[21] / cfg.DX, 2000.0
    rig1, rjg1, hgt1 = synth.make_grid(cfg, nprocs_e=nprocs_e, myrank_e=myrank_e)
    ens = None
    if synth_kind == "hx":
        ens = synth.SmoothEnsemble(cfg, seed_no=4, hlen=hlen, vlen=vlen, device=device)
        obs = ens.attach(obs)
    return cfg, obs, rig1, rjg1, hgt1, ens


def host_threads():
    """threads the CPU legs may use: the affinity mask, not OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_to_gpu_numa(index):
    """Pin this rank (and hence the pinned host buffers it allocates afterwards: first touch) to the CPUs of the NUMA
    node its GPU hangs off.  Round 1 ran all eight ranks on node 0: the host-buffer (e2e) leg moved 35 GB through one
    socket's memory at 131 GB/s aggregate.  Returns a small record for the JSON line; never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:   # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "note": "the platform reports no NUMA node for this GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return {"node": node, "note": "no allowed CPU on the GPU's node"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as e:
        return {"node": None, "note": repr(e)[:120]}


def algorithmic_work(k, nv, npoints, nsolved, nobsl_sum):
    """F(k,p) = 2pk^2 + 9k^3 + 4k^3 + (2pk + 2k^2) + nv(2k^2 + 2k^2) per
    solved point, B(k,p) = 2 nv (k+1) 8 per analysed point + p (k+4) 8 per solved point."""
    flops = nobsl_sum * (2.0 * k * k + 2.0 * k) + nsolved * (13.0 * k ** 3 + 2.0 * k * k + nv * 4.0 * k * k)
    bytes_ = npoints * 2.0 * nv * (k + 1) * 8.0 + nobsl_sum * (k + 4) * 8.0
    return flops, bytes_


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:   # NVML missing: report it, never fake clocks
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2.0)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def fp64_peak_live():
    """FP64 FMA / DMMA peak measured live by tools/fp64_peak (MEASURED_PEAKS.json has no FP64
    figure).  Falls back to the value recorded under profiles/ on the same pool."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip()
        j = json.loads(out.splitlines()[-1])
        return max(j["fp64_fma_tflops"], j["fp64_dmma_tflops"]), "measured live (tools/fp64_peak.cu)"
    except Exception:
        return 36.6, "recorded (profiles/fp64_peak_r01.json)"


# ---------------------------------------------------------------------------------------------
class CpuSample:
    """The oracle (CPU restatement of the reference algorithm, OpenMP over ij inside a level loop
    like letkf_tools.f90:289-320) on a cyclic-deal sample of the workload's columns: every
    `npe`-th column of the plane, all levels.  calibrate() sizes the sample for ~target_s."""

    def __init__(self, name, nthreads=0, synth_kind="hx"):
        from oracle import oracle_py
        oracle_py.build()
        self.name, self.w = name, WORKLOADS[name]
        self.oracle_py = oracle_py
        self.nth = nthreads or host_threads()
        self.synth_kind = synth_kind
        self.ncol = self.w["nlon"] * self.w["nlat"]

    def _deal_width(self, target_cols):   # coprime with nlon: spreads the sample over the plane
        npe = max(1, self.ncol // max(1, target_cols))
        while npe > 1 and np.gcd(npe, self.w["nlon"]) != 1:
            npe += 1
        return npe

    def prepare(self, npe):
        from scale_letkf_b200 import synth
        if getattr(self, "_obs", None) is None:   # the observation set does not depend on the column sample
            _, self._obs, _, _, _, _ = make_workload(self.name, synth_kind=self.synth_kind)
        cfg, _, rig1, rjg1, hgt1, ens = make_workload(self.name, nprocs_e=npe, myrank_e=npe // 3, synth_kind="grid")
        o = self.oracle_py.Oracle(cfg)
        o.set_obs(self._obs)
        o.set_grid(rig1, rjg1, hgt1)
        self.o, self.npe = o, npe
        if self.synth_kind == "hx":
            w = self.w
            hl, vl = (cfg.HORI_LOCAL[0] / cfg.DX, 3000.0) if w["kind"] == "sonde" else (cfg.HORI_LOCAL[21] / cfg.DX, 2000.0)
            self.gues0 = synth.SmoothEnsemble(cfg, seed_no=4, hlen=hl, vlen=vl).state(rig1, rjg1, as_numpy=True)
        else:
            self.gues0 = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=4)

    def run(self):
        g = self.gues0.copy(order="F")
        t0 = time.perf_counter()
        r = self.o.das_letkf(g, nthreads=self.nth)
        dt = time.perf_counter() - t0
        return r["npoints"], dt, r["nsolved"]

    def calibrate(self, target_s, calib_points=1500):
        # the eigensolve costs O(k^3): shrink the calibration sample for large ensembles
        calib_points = max(self.w["nlev"], int(calib_points * min(1.0, (100.0 / self.w["member"]) ** 3)))
        self.prepare(self._deal_width(max(1, calib_points // self.w["nlev"])))
        npts, dt, _ = self.run()
        want_cols = max(1, int(npts / dt * target_s / self.w["nlev"]))
        self.prepare(self._deal_width(min(want_cols, self.ncol)))

    def describe(self, npts, nsolved, dt):
        w = self.w
        return (f"oracle das_letkf on every {self.npe}-th column of the {w['nlon']}x{w['nlat']} plane x "
                f"{w['nlev']} levels = {npts} points ({nsolved} solved) per pass, {self.nth} OpenMP threads, "
                f"{dt:.1f} s per pass")


def run_c1(args):
    """BASELINE config 1 (the reference's own CPU-runnable case): letkf_core batch, 20 members, 10^4 grid points,
    <= 100 local obs each, fp64 -- through letkf_b200_core_batch with HOST buffers (this call has no
    device-resident form in the reference's interface), next to the oracle on all host threads."""
    import torch
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    c = synth.make_core_batch(ne=20, npts=10000, nobs=100, seed_no=1)
    a = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"], c["parm_infl"])
    if args.impl == "reference":
        from oracle import oracle_py
        oracle_py.build()
        run = lambda: oracle_py.core_batch(*a)
        cores = oracle_py.max_threads()
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference)")
        eng = sl.LETKF(sl.resolve_config(sl.default_config(MEMBER=20, nlon=8, nlat=8, nlev=2)), device=0)
        run = lambda: eng.letkf_core(*a)
        cores = 0
    ts = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        run()
        if args.impl != "reference":
            torch.cuda.synchronize()
        if i >= args.warmup:
            ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    k, p = 20, float(c["nobsl"].mean())
    flops = c["npts"] * (2.0 * p * k * k + 13.0 * k ** 3 + 2.0 * p * k + 2.0 * k * k)
    nbytes = c["hdxb"].nbytes + 3 * c["rdiag"].nbytes + 2 * c["npts"] * k * k * 8 + c["npts"] * k * 8
    line = {"metric": METRIC, "value": c["npts"] / dt, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C1: letkf_core batch, 20 members, 10^4 grid points, <= 100 local obs each",
                       "k": k, "points": c["npts"], "mean_local_obs": p,
                       "note": "host buffers in and out (trans, transm, pao): H2D/D2H inside the timed call"},
            "e2e": {"value": c["npts"] / dt, "unit": UNIT, "h2d_bytes_per_step": int(c["hdxb"].nbytes + 3 * c["rdiag"].nbytes),
                    "d2h_bytes_per_step": int(2 * c["npts"] * k * k * 8 + c["npts"] * k * 8)},
            "gpu_launches": 0 if args.impl == "reference" else args.steps,
            "roofline": {"kernel": "core_kernel<20>", "bound": "hbm", "achieved": nbytes / dt * 1e-9, "peak": None,
                         "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "PCIe-bound through host buffers; algorithmic %.2f GFLOP per call" % (flops * 1e-9)}}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": line["value"], "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "oracle letkf_core on all 10^4 points"}
        line["e2e"] = {"value": line["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))
    return 0


def c1_record():
    """BASELINE config 1 inside the default line: letkf_core batch (k = 20, 10^4 points) through the C ABI with host buffers,
    next to the oracle on 1 thread ("1 MPI rank") and on all host threads."""
    import torch
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    from oracle import oracle_py
    oracle_py.build()
    c = synth.make_core_batch(ne=20, npts=10000, nobs=100, seed_no=1)
    a = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"], c["parm_infl"])
    eng = sl.LETKF(sl.resolve_config(sl.default_config(MEMBER=20, nlon=8, nlat=8, nlev=2)), device=torch.cuda.current_device())

    def tm(fn, n):
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n

    tg = tm(lambda: eng.letkf_core(*a), 5)
    t1 = tm(lambda: oracle_py.core_batch(*a, nthreads=1), 1)
    nth = host_threads()
    tn = tm(lambda: oracle_py.core_batch(*a, nthreads=nth), 2)
    r = eng.letkf_core(*a)
    ref = oracle_py.core_batch(*a, nthreads=nth)
    err = max(float(np.abs(r[q] - ref[q]).max() / np.abs(ref[q]).max()) for q in ("trans", "transm", "pao"))
    eng.close()
    return {"workload": "C1: letkf_core batch, 20 members, 10^4 grid points, <= 100 local obs each (host buffers)",
            "points_per_s": c["npts"] / tg, "ms": tg * 1e3, "cpu_1_thread_points_per_s": c["npts"] / t1,
            "cpu_all_threads_points_per_s": c["npts"] / tn, "cpu_threads": nth, "max_rel_err_vs_oracle": err}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (C++ restatement under oracle/ -- the Fortran
    original cannot be built here: no Fortran compiler, no MPI, SCALE-RM/NetCDF not vendored)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    name = args.workload
    cs = CpuSample(name, synth_kind=args.synth)
    cs.calibrate(args.cpu_seconds / 4.0)
    tot_pts, tot_s, nsolved, npts = 0, 0.0, 0, 0
    for i in range(args.warmup + args.steps):
        npts, dt, nsolved = cs.run()
        if i >= args.warmup:
            tot_pts += npts
            tot_s += dt
    value = tot_pts / tot_s
    w = WORKLOADS[name]
    desc = cs.describe(npts, nsolved, tot_s / max(args.steps, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["desc"], "k": w["member"], "grid": [w["nlon"], w["nlat"], w["nlev"]],
                   "step": "one pass over a bounded sample of the workload (see cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cs.nth, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def measure(name, ctx, args, steps, warmup, legs):
    """One workload on all ranks: device-resident throughput, in-bench parity, roofline, full cycle, host-buffer
    end to end, CPU baseline (rank 0, N = 1).  `legs`: set of {"parity", "cycle", "e2e", "cpu"}.  Returns a dict."""
    import torch
    import torch.distributed as dist
    import scale_letkf_b200 as sl
    from scale_letkf_b200 import synth
    world, rank, local, dev, barrier = ctx["world"], ctx["rank"], ctx["local"], ctx["dev"], ctx["barrier"]
    w = WORKLOADS[name]
    cfg, obs, rig1, rjg1, hgt1, ens = make_workload(name, nprocs_e=world * args.subsample, myrank_e=rank, device=dev,
                                                    synth_kind=args.synth)
    k, nv, nlev = cfg.MEMBER, cfg.nv3d, cfg.nlev
    nij1 = len(rig1)
    eng = sl.LETKF(cfg, device=local)
    eng.set_letkf_obs(obs)
    eng.set_common_mpi_grid(rig1, rjg1, hgt1)
    if ens is not None:      # one global state: every value is a function of the global grid index
        gues0 = ens.state(rig1, rjg1)
    else:
        gen = torch.Generator(device=dev)
        gen.manual_seed(20260102 + rank)
        gues0 = synth.make_state(cfg, rig1, rjg1, hgt1, xp=torch, device=dev, gen=gen)
    gues = torch.empty_like(gues0)
    anal = torch.empty_like(gues0)
    state_bytes = gues0.numel() * 8

    # ---- device-resident leg ------------------------------------------------------------------
    def step():
        gues.copy_(gues0)            # das_letkf destroys gues (INTENT(INOUT)); restore is untimed
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = eng.das_letkf(gues, anal3d=anal)
        e1.record()
        return e0, e1, out

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    for _ in range(steps):
        evs.append(step())
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b, _ in evs]
    kern_ms = [o["kernel_ms"] for _, _, o in evs]
    out = evs[-1][2]
    ph = np.array(out["phase_clocks"], dtype=np.float64)
    phases = dict(zip(["load", "search", "gram", "factor", "eigen", "apply", "store", "sched"],
                      (ph / max(ph.sum(), 1.0)).round(4).tolist()))
    phases["solver_iterations_per_solve"] = out["solver_iterations"] / max(out["nsolved"], 1)
    phases["refined_share_of_solves"] = out.get("nrefined", 0) / max(out["nsolved"], 1)   # ill-conditioned path (explicit Z + refinement)
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms, float(sum(kern_ms))], dtype=torch.float64, device=dev)
    cnt = torch.tensor([out["npoints"], out["nsolved"], out["nobsl_sum"], out["launches"]], dtype=torch.float64,
                       device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms, kernel_ms_sum = float(t[0]), float(t[1])
    npoints, nsolved, nobsl_sum, launches = (float(x) for x in cnt)
    value = npoints * steps / (total_ms * 1e-3)
    ms_per_step = total_ms / steps

    # ---- parity inside the bench: the analysis just timed against the oracle on a sample of this rank's columns ----
    parity = None
    if "parity" in legs:
        perr = None
        pv = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=torch.float64, device=dev)   # err, nobsl mismatch, points, failed
        try:
            from oracle import oracle_py
            oracle_py.build()
            want_pts = max(4, int(args.parity_points * min(1.0, (50.0 / k) ** 3)))
            ncs = max(1, min(nij1, want_pts // nlev))   # whole columns; fewer points than one column: a level mask
            cols = np.unique(np.linspace(0, nij1 - 1, ncs).astype(np.int64))
            ct = torch.as_tensor(cols, device=dev)
            g_s = np.asfortranarray(gues0[:, :, :, ct].cpu().numpy().T)          # (ncols, nlev, nens, nv3d)
            gues.copy_(gues0)
            nb = eng.das_letkf(gues, anal3d=anal, want_nobsl=True)["nobsl"][:, ct].cpu().numpy().T
            a_gpu = anal[:, :k, :, ct].cpu().numpy().T                           # (ncols, nlev, k, nv3d)
            o = oracle_py.Oracle(cfg)
            o.set_obs(obs)
            o.set_grid(rig1[cols], rjg1[cols], np.asfortranarray(hgt1[cols]))
            mask = None
            if want_pts < len(cols) * nlev:   # large ensembles: the EISPACK oracle costs ~10 k^3 flops per point
                mask = np.zeros((len(cols), nlev), dtype=np.uint8)
                mask[0, np.unique(np.linspace(0, nlev - 1, max(2, want_pts)).astype(int))] = 1
            ref = o.das_letkf(g_s, want_nobsl=True, point_mask=mask, nthreads=max(1, host_threads() // world))
            sel = np.ones((len(cols), nlev), dtype=bool) if mask is None else mask.astype(bool)
            b = ref["anal3d"][:, :, :k, :]
            sc = np.maximum(np.abs(b[sel]).max(axis=(0, 1), keepdims=True), 1e-300)
            pv = torch.tensor([float((np.abs(a_gpu[sel] - b[sel]) / sc).max()),
                               0.0 if np.array_equal(nb[sel], ref["nobsl"][sel]) else 1.0,
                               float(sel.sum()), 0.0], dtype=torch.float64, device=dev)
            del o
        except Exception as e:
            perr = repr(e)[:300]
        if world > 1:   # outside the try block: every rank takes part whatever happened on it
            mx = pv.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(pv, op=dist.ReduceOp.SUM)
            pv[0], pv[1], pv[3] = mx[0], mx[1], mx[3]
        if float(pv[3]) != 0.0:
            parity = {"error": perr or "the parity leg failed on another rank"}
        else:
            parity = {"nobsl_equal": bool(pv[1] == 0.0), "max_rel_err": float(pv[0]), "points": int(pv[2]),
                      "what": "GPU anal3d / nobsl of the timed workload vs the CPU oracle on a column sample of every rank, "
                              "max |a-b| / max|b| per variable", "tolerance": 1e-10}

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks, peaks_src = measured_peaks()
    flops, abytes = algorithmic_work(k, nv, npoints, nsolved, nobsl_sum)   # whole job, per step
    kms = kernel_ms_sum / steps                                            # max-rank kernel time per step
    if "fp64_peak" not in ctx:
        ctx["fp64_peak"] = fp64_peak_live() if rank == 0 else (36.6, "")
    fp64_peak, fp64_src = ctx["fp64_peak"]
    ach_tf = flops / world / (kms * 1e-3) * 1e-12                          # per GPU
    ach_gbs = abytes / world / (kms * 1e-3) * 1e-9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "das_kernel_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                tj = json.load(f).get(name)
            traffic = tj["bytes_per_point"] * npoints / world if tj else None
        except Exception:
            traffic = None
    roofline = {
        "kernel": "das_ns_kernel" if k <= 102 else "das_tiled", "bound": "tensor", "pipe": "fp64 DMMA (mma.sync m8n8k4.f64)",
        "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak, "traffic": traffic,
        "traffic_source": "ncu capture under profiles/ (bytes per point x points of this run), not measured in this run",
        "peak_source": fp64_src + "; MEASURED_PEAKS.json holds no FP64 figure",
        "kernel_ms_per_step": kms, "launches_per_step": int(launches / world),
        "kernel_ms_per_launch": kms / max(int(launches / world), 1),
        "traffic_per_launch": None if traffic is None else traffic / max(int(launches / world), 1),
        "algorithmic_flops_per_step": flops / world, "algorithmic_bytes_per_step": abytes / world,
        "hbm": {"achieved": ach_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": ach_gbs / peaks.get("hbm_gbs"), "peak_source": peaks_src + " MEASURED_PEAKS.json"},
    }

    if ctx.get("on_partial") is not None:   # from here on only optional legs follow (cycle, e2e, cpu): a watchdog may print this
        ctx["on_partial"](dict(value=value, ms_per_step=ms_per_step, clocks=clocks, phases=phases, parity=parity, roofline=roofline,
                               cycle={"error": "the leg did not finish before the deadline"}, e2e=None, cpu=None, launches=launches,
                               npoints=npoints, nsolved=nsolved, nobsl_sum=nobsl_sum, kms=kms, state_bytes=state_bytes, k=k,
                               nobs=int(len(obs["elm"])), desc=w["desc"], grid=[w["nlon"], w["nlat"], w["nlev"]]))

    # ---- full analysis cycle (SURVEY.md section 8d metric ii) ----------------------------------------
    cycle = None
    free_b, _ = torch.cuda.mem_get_info()
    if world > 1:   # one decision for all ranks: the cycle leg is full of collectives
        fb = torch.tensor([float(free_b)], dtype=torch.float64, device=dev)
        dist.all_reduce(fb, op=dist.ReduceOp.MIN)
        free_b = float(fb[0])
    if "cycle" in legs and args.subsample == 1 and free_b < 2.2 * state_bytes:
        cycle = {"skipped": "member-major input and output grids (2 x %.1f GB) do not fit next to the three state arrays "
                            "on this GPU count" % (state_bytes / 1e9)}
    elif "cycle" in legs and args.subsample == 1:
        try:
            from scale_letkf_b200.transpose import EnsTranspose, EnsTransposeP2P
            nens = gues0.shape[1]
            thermo = eng.thermo_defaults()
            if args.transpose == "p2p":   # one pass: local reads, stores straight into the receiving rank's array (NVLink)
                p2p = EnsTransposeP2P(eng, world, rank, thermo=thermo)

                class _Tr:   # the interface of EnsTranspose the loop below uses
                    block = 0
                    rounds = staticmethod(p2p.rounds)
                    read_ens = staticmethod(lambda g3, g2, v3, v2, km, ne: p2p.read_ens(g3, v3, km, ne))
                    write_ens = staticmethod(lambda v3, v2, g3, g2, km, ne: p2p.write_ens(v3, g3, km, ne))
                tr = _Tr()
            else:
                tr = EnsTranspose(eng, world, rank, nlev, nv, 0, group=None, device=dev, thermo=thermo)
            gsz = nlev * w["nlon"] * w["nlat"] * nv
            rounds = list(tr.rounds(k))
            gin = [torch.empty(gsz, dtype=torch.float64, device=dev) if im is not None else None for _, im, _, _ in rounds]
            gout = [torch.empty(gsz, dtype=torch.float64, device=dev) if im is not None else None for _, im, _, _ in rounds]
            gues.copy_(gues0)
            tr.write_ens(gues, None, gin, None, k, nens)      # untimed: the member-major input of the cycle
            chk_in = float(sum(float(g.sum()) for g in gin if g is not None))
            names = ["transpose_in+state_trans", "ensmean", "set_obs", "das_letkf", "anal_mean",
                     "transpose_out+state_trans_inv"]
            obs_dev = None
            if args.obs_chain == "device":   # the observation tables stay in HBM (as the device observation operator and
                #                              departure/QC kernels leave them); set_obs filters / sorts them on the device
                obs_dev = {kf: torch.as_tensor(np.ascontiguousarray(obs[kf], dtype=np.int32 if kf in ("elm", "typ") else np.float64),
                                               device=dev) for kf in ("elm", "typ", "ri", "rj", "lev", "dat", "err", "val", "ensval")}
            recs = []
            for i in range(2 + args.cycle_steps):
                barrier()
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
                ev[0].record()
                tr.read_ens(gin, None, gues, None, k, nens)   # state_trans fused in the pack
                ev[1].record()
                eng.ensmean_grd(gues)
                ev[2].record()
                th0 = time.perf_counter()
                if obs_dev is not None:
                    eng.set_letkf_obs_device(obs_dev, None)
                else:
                    eng.set_letkf_obs(obs)
                host_setobs_ms = (time.perf_counter() - th0) * 1e3
                ev[3].record()
                eng.das_letkf(gues, anal3d=anal)
                ev[4].record()
                eng.ensmean_grd(anal)
                ev[5].record()
                tr.write_ens(anal, None, gout, None, k, nens)  # state_trans_inv fused in the unpack
                ev[6].record()
                barrier()
                if i >= 2:
                    recs.append([ev[j].elapsed_time(ev[j + 1]) for j in range(len(names))])
            arr = torch.tensor(recs, dtype=torch.float64, device=dev)       # (steps, phases)
            tot = arr.sum(dim=1)
            # round trip of the transposes alone: in -> out must give the member-major input back
            tr.read_ens(gin, None, gues, None, k, nens)
            tr.write_ens(gues, None, gout, None, k, nens)
            rt = torch.tensor([max([float((a - b).abs().max() / a.abs().max()) for a, b in zip(gin, gout) if a is not None]
                                   + [0.0])], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(arr, op=dist.ReduceOp.MAX)
                dist.all_reduce(tot, op=dist.ReduceOp.MAX)
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            cycle = {"ms_median": float(tot.median()), "ms_min": float(tot.min()), "steps": args.cycle_steps,
                     "phases_ms_median": dict(zip(names, [round(float(x), 3) for x in arr.median(dim=0).values])),
                     "transposes": ("one pass over peer memory (letkf_b200_scatter_grd_p2p / _gather_grd_p2p: local reads, "
                                    "NVLink stores into the receiving rank's array)") if args.transpose == "p2p" else
                                   "CUDA pack + NCCL all_to_all_single + CUDA unpack",
                     "transpose_gbs_per_gpu": {nm: round(2.0 * state_bytes * k / nens / (float(x) * 1e-3) * 1e-9, 1)
                                               for nm, x in zip(names, arr.median(dim=0).values) if nm.startswith("transpose")},
                     "set_obs_host_ms_last": round(host_setobs_ms, 3),
                     "set_obs": "device-resident tables (letkf_b200_set_obs_device)" if obs_dev is not None else "host tables (H2D inside)",
                     "transpose_round_trip_max_rel_err": float(rt[0]), "input_checksum_rank0": chk_in,
                     "what": "restart variables -> transpose in (state_trans fused) + mean + obs bucketing + analysis + mean + transpose out (state_trans_inv fused), n_gpus ranks"}
            del gin, gout, tr, obs_dev
        except Exception as e:   # never lose the main measurement to the optional leg
            cycle = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- end-to-end leg: host buffers through the C ABI -----------------------------------------
    e2e = None
    if "e2e" in legs:
        hg = torch.empty(gues0.shape, dtype=torch.float64, pin_memory=True)
        hg.copy_(gues0)
        ha = torch.empty(gues0.shape, dtype=torch.float64, pin_memory=True)
        del gues0, gues, anal
        gues0 = gues = anal = None
        torch.cuda.empty_cache()
        g_np = hg.numpy().T      # (nij1, nlev, nens, nv3d) Fortran order view of the same memory
        a_np = ha.numpy().T
        nst = args.e2e_steps or steps
        e2e_ms = []
        for i in range(min(warmup, 2) + nst):
            barrier()                # host gues3d is left untouched by the call (copy_back_gues=False)
            t0 = time.perf_counter()
            eng.das_letkf(g_np, anal3d=a_np, copy_back_gues=False)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            if i >= min(warmup, 2):
                e2e_ms.append(dt)
        te = torch.tensor([float(sum(e2e_ms))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": npoints * nst / (float(te[0]) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": state_bytes * world, "d2h_bytes_per_step": state_bytes * world,
               "ms_per_step": float(te[0]) / nst, "steps": nst,
               "api": "letkf_b200_das_letkf(mem_space=HOST) via scale_letkf_b200.LETKF.das_letkf; pinned host "
                      "gues3d in, anal3d out (gues3d perturbations are not copied back; ln p on the host inside the call)"}
        del hg, ha

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and "cpu" in legs:
        cs = CpuSample(name, synth_kind=args.synth)
        cs.calibrate(args.cpu_seconds)
        npts_c, dt_c, nsolved_c = cs.run()
        cpu = {"value": npts_c / dt_c, "unit": UNIT, "cores": cs.nth, "kind": "port",
               "sample": cs.describe(npts_c, nsolved_c, dt_c)}
    del gues0, gues, anal
    eng.close()
    torch.cuda.empty_cache()
    return dict(value=value, ms_per_step=ms_per_step, clocks=clocks, phases=phases, parity=parity, roofline=roofline,
                cycle=cycle, e2e=e2e, cpu=cpu, launches=launches, npoints=npoints, nsolved=nsolved, nobsl_sum=nobsl_sum,
                kms=kms, state_bytes=state_bytes, k=k, nobs=int(len(obs["elm"])), desc=w["desc"],
                grid=[w["nlon"], w["nlat"], w["nlev"]])


def brief(r, steps):
    """sub-record of a secondary workload inside the default line"""
    return {"workload": r["desc"], "k": r["k"], "points": int(r["npoints"]), "mean_local_obs": r["nobsl_sum"] / max(r["nsolved"], 1.0),
            "analysis_ms": r["ms_per_step"], "points_per_s": r["value"], "steps": steps,
            "roofline_frac": r["roofline"]["frac"], "roofline_tflops_per_gpu": r["roofline"]["achieved"],
            "solver_iterations_per_solve": r["phases"]["solver_iterations_per_solve"],
            "cycle_ms": None if not r["cycle"] else r["cycle"].get("ms_median"),
            "cycle_phases_ms": None if not r["cycle"] else r["cycle"].get("phases_ms_median"),
            "cycle_note": None if not r["cycle"] else (r["cycle"].get("error") or r["cycle"].get("skipped")),
            "parity": r["parity"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c1"])
    ap.add_argument("--cpu-seconds", type=float, default=16.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check against the oracle")
    ap.add_argument("--parity-points", type=int, default=4000, help="points of the in-bench parity sample at k = 50")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed e2e steps (default: --steps)")
    ap.add_argument("--no-cycle", dest="cycle", action="store_false",
                    help="skip the full-cycle leg (transposes + bucketing + analysis)")
    ap.set_defaults(cycle=True)
    ap.add_argument("--cycle-steps", type=int, default=3)
    ap.add_argument("--obs-chain", default="device", choices=["device", "host"],
                    help="cycle leg: observation tables resident in HBM (letkf_b200_set_obs_device) or uploaded from the host every cycle")
    ap.add_argument("--transpose", default="p2p", choices=["p2p", "nccl"],
                    help="member<->grid transposes of the cycle leg: one-pass peer-memory kernels (default) or pack + NCCL + unpack")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the secondary records of the default line (k100 = C3, c1, and at 8 GPUs c5_cycle and c4)")
    ap.add_argument("--all-extra", action="store_true", help="run the 8-GPU secondary records (c5_cycle, c4) at any rank count")
    ap.add_argument("--main-deadline", type=float, default=420.0,
                    help="seconds the optional legs of the main workload (cycle, e2e, cpu) may take before the line is printed without them")
    ap.add_argument("--extra-deadline", type=float, default=480.0,
                    help="seconds after which the secondary records are abandoned and the line is printed without them")
    ap.add_argument("--synth", default="hx", choices=["hx", "iid"],
                    help="hx: smooth correlated members, ensval = H(x_m) (default); iid: round-1 generator")
    ap.add_argument("--subsample", type=int, default=1,
                    help="profiling aid: analyse only every S-th column of the plane (same per-point work)")
    args = ap.parse_args()
    if args.workload == "c1":
        return run_c1(args) if int(os.environ.get("RANK", "0")) == 0 else 0
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout_fd = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on file descriptor 1; stdout must carry the ONE JSON line only, so fd 1
        # points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    numa = bind_to_gpu_numa(local) if world > 1 else None   # (one rank: keep every core for the CPU baseline leg)
    ctx = dict(world=world, rank=rank, local=local, dev=dev, barrier=barrier)
    legs = set()
    if not args.no_parity:
        legs.add("parity")
    if args.cycle:
        legs.add("cycle")
    if not args.no_e2e:
        legs.add("e2e")
    if not args.no_cpu:
        legs.add("cpu")
    # ---- the line (rank 0); emitted once, by the main thread or -- if an optional leg hangs -- by a deadline thread ----
    extra = {}
    emitted = threading.Lock()
    res = {}

    def emit(note=None):
        if not emitted.acquire(blocking=False):
            return
        if rank != 0:
            return
        r = res["r"]
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": r["desc"], "k": r["k"], "grid": r["grid"],
                       "nobs": r["nobs"], "mean_local_obs": r["nobsl_sum"] / max(r["nsolved"], 1.0),
                       "points": int(r["npoints"]), "solved_points": int(r["nsolved"]),
                       "synthetic_inputs": "members = smooth random fields + gridpoint noise, ensval = H(x_m) by tri-linear "
                                           "interpolation (one global state for every rank count)" if args.synth == "hx"
                                           else "i.i.d. observation-space ensemble (round-1 generator)",
                       "decomposition": f"cyclic column deal over {world} rank(s), obs replicated, no collective",
                       "l2": "inputs (%.1f GB state per rank) far larger than the 126 MB L2" % (r["state_bytes"] / 1e9)},
            "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": int(r["launches"] / world) * args.steps,
            "roofline": r["roofline"], "cpu_baseline": r["cpu"], "parity": r["parity"], "cycle": r["cycle"],
            "kernel_ms_per_step": r["kms"], "phase_share_rank0": r["phases"],
        }
        if numa is not None:
            line["numa_rank0"] = numa
        line.update(dict(extra))
        if note:
            line["extras_note"] = note
        if args.subsample > 1:
            line["config"]["subsample"] = f"every {args.subsample}-th column only (profiling aid, not a bench value)"
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)

    # The device-timed measurement and its parity check are never lost to an optional leg (cycle: peer-memory transposes
    # and barriers; e2e; cpu baseline) that hangs in a collective: every rank arms the same deadline when measure() hands out
    # the partial result, rank 0 prints the line without the unfinished legs, every rank exits 0.
    main_done = threading.Event()

    def on_partial(rp):
        res["r"] = rp

        def main_deadline():
            if not main_done.wait(args.main_deadline):
                emit("optional legs (cycle / e2e / cpu_baseline) stopped after %.0f s (deadline); value, roofline and parity "
                     "are complete" % args.main_deadline)
                sys.stdout.flush()
                os._exit(0)
        threading.Thread(target=main_deadline, daemon=True).start()
    ctx["on_partial"] = on_partial
    res["r"] = measure(args.workload, ctx, args, args.steps, args.warmup, legs)
    main_done.set()
    ctx["on_partial"] = None

    # ---- secondary records of the default line (the other BASELINE.json shapes) ------------------------
    if args.workload == "c2" and not args.no_extra and args.subsample == 1:
        done = threading.Event()

        def deadline():   # the main measurement is never lost to a secondary record that hangs in a collective
            if not done.wait(args.extra_deadline):
                emit("secondary records stopped after %.0f s (deadline); the records present are complete" % args.extra_deadline)
                sys.stdout.flush()
                os._exit(0)
        threading.Thread(target=deadline, daemon=True).start()

        def sub(key, name, steps, warmup, lg):
            try:
                extra[key] = brief(measure(name, ctx, args, steps, warmup, lg), steps)
            except Exception as e:
                extra[key] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()
        sub("k100", "c3", 3, 1, {"parity", "cycle"})            # C3: 256x256x60, k = 100, dense radar
        if world >= 8 or args.all_extra:
            sub("c5_cycle", "c5", 2, 1, {"cycle"})              # C5: 400x400x60 full cycle
            sub("c4", "c4", 1, 1, {"parity"})                   # C4: 128x128x40, k = 1000, one full-size pass
        if rank == 0:
            try:
                extra["c1"] = c1_record()
            except Exception as e:
                extra["c1"] = {"error": repr(e)[:300]}
        done.set()
    emit()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
