#!/usr/bin/env python
"""Prints the achieved relative error of das_letkf (CUDA vs oracle) on a few test-size cases:
the margin under the 1e-10 acceptance bar.  Needs a GPU.   python tools/accuracy_report.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scale_letkf_b200 as sl   # noqa: E402
from helpers import sonde_case, radar_case, host_logp, relerr   # noqa: E402
from oracle import oracle_py   # noqa: E402

oracle_py.build()
CASES = [("sonde k=20", sonde_case(member=20, nsonde=30, nsfc=100)),
         ("sonde k=50", sonde_case(member=50, nsonde=40, nsfc=150)),
         ("radar k=50 max_nobs=100", radar_case(member=50, max_nobs=100, nlon=32, nlat=32, nlev=6)),
         ("radar k=100 max_nobs=200", radar_case(member=100, max_nobs=200, nlon=20, nlat=20, nlev=4, radius=4.0e3)),
         ("radar k=136 (tiled)", radar_case(member=136, max_nobs=40, nlon=20, nlat=20, nlev=4, radius=4.0e3))]
for name, (cfg, rig1, rjg1, hgt1, obs, gues) in CASES:
    o = oracle_py.Oracle(cfg)
    o.set_obs(obs)
    o.set_grid(rig1, rjg1, hgt1)
    e = sl.LETKF(cfg, device=0)
    e.set_letkf_obs(obs)
    e.set_common_mpi_grid(rig1, rjg1, hgt1)
    ref = o.das_letkf(gues.copy(order="F"))
    out = e.das_letkf(gues.copy(order="F"), logp=host_logp(cfg, gues))
    k = cfg.MEMBER
    err = relerr(out["anal3d"][:, :, :k, :], ref["anal3d"][:, :, :k, :], axis=(0, 1, 2))
    print(f"{name:28s} max rel err {err:.2e}  iterations/solve {out['solver_iterations'] / max(out['nsolved'], 1):.2f}")
    e.close()
