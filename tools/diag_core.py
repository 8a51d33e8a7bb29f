"""GPU diagnostic: per-point error of the letkf_core batch vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scale_letkf_b200 as sl
from scale_letkf_b200 import synth
from oracle import oracle_py

cfg = sl.resolve_config(sl.default_config(MEMBER=20, nlon=8, nlat=8, nlev=2))
e = sl.LETKF(cfg, device=0)
for ne, nobs, npts in [(20, 100, 2000), (50, 300, 200), (100, 400, 64)]:
    c = synth.make_core_batch(ne=ne, npts=npts, nobs=nobs, seed_no=1, det=True)
    args = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"], c["parm_infl"])
    ref = oracle_py.core_batch(*args, depd=c["depd"])
    r = e.letkf_core(*args, depd=c["depd"])
    for key in ("trans", "pao", "transm", "transmd"):
        d = np.abs(r[key] - ref[key]).reshape(npts, -1).max(axis=1)
        sc = np.abs(ref[key]).reshape(npts, -1).max(axis=1)
        rel = d / np.maximum(sc, 1e-300)
        w = np.argsort(rel)[::-1][:8]
        print(ne, key, "max rel", rel.max(), "median", np.median(rel), "n>1e-10:", int((rel > 1e-10).sum()))
        print("   worst pts", w.tolist(), "nobsl", c["nobsl"][w].tolist(), "rel", [f"{x:.1e}" for x in rel[w]])
    # invariants of the GPU result itself
    W, Pa = r["trans"], r["pao"]
    inv = np.abs(W @ W - (ne - 1) * Pa).reshape(npts, -1).max(axis=1) / np.abs(Pa).reshape(npts, -1).max(axis=1)
    print(ne, "GPU  W W - (k-1) Pa:", inv.max())
    W, Pa = ref["trans"], ref["pao"]
    inv = np.abs(W @ W - (ne - 1) * Pa).reshape(npts, -1).max(axis=1) / np.abs(Pa).reshape(npts, -1).max(axis=1)
    print(ne, "ORCL W W - (k-1) Pa:", inv.max())
    # independent: numpy eigh
    bad = 0
    for i in range(min(npts, 300)):
        p = c["nobsl"][i]
        if p == 0:
            continue
        Y = c["hdxb"][i][:, :p].T          # (p, ne)
        A = (Y / c["rdiag"][i][:p, None]).T @ Y + (ne - 1) / c["parm_infl"][i] * np.eye(ne)
        lam, V = np.linalg.eigh(A)
        Wn = (V * np.sqrt((ne - 1) / lam)) @ V.T
        eg = np.abs(r["trans"][i] - Wn).max() / np.abs(Wn).max()
        eo = np.abs(ref["trans"][i] - Wn).max() / np.abs(Wn).max()
        if eg > 1e-10 or eo > 1e-10:
            bad += 1
            if bad < 6:
                print("   pt", i, "p", p, "gpu-vs-eigh", eg, "oracle-vs-eigh", eo)
    print(ne, "points off vs numpy eigh:", bad)
