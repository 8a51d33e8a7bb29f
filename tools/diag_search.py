"""GPU diagnostic: obs_local selection vs the oracle (no obs-number limit case)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scale_letkf_b200 as sl
from oracle import oracle_py
from helpers import sonde_case, sample_points

cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(nsonde=40, nsfc=200)
o = oracle_py.Oracle(cfg); o.set_obs(obs); o.set_grid(rig1, rjg1, hgt1)
e = sl.LETKF(cfg, device=0); e.set_letkf_obs(obs); e.set_common_mpi_grid(rig1, rjg1, hgt1)
pts = sample_points(cfg, rig1, rjg1, hgt1, gues, stride=3)
n1, i1, d1, l1 = o.obs_local(*pts, 1, 4096)
n2, i2, d2, l2 = e.obs_local(*pts, 1, 4096)
print("npts", len(n1), "n equal", np.array_equal(n1, n2), "max", n1.max(), n2.max())
bad = np.nonzero(n1 != n2)[0]
print("count mismatches", len(bad), bad[:10], n1[bad[:10]], n2[bad[:10]])
print("sorted idx equal", np.array_equal(o.sorted_index(), e.sorted_index()))
nb = 0
for p in range(len(n1)):
    a, b = i1[p, :n1[p]], i2[p, :n2[p]]
    if not np.array_equal(a, b):
        nb += 1
        if nb <= 5:
            same_set = np.array_equal(np.sort(a), np.sort(b))
            j = np.nonzero(a[:min(len(a), len(b))] != b[:min(len(a), len(b))])[0]
            print("pt", p, "n", n1[p], n2[p], "same set", same_set, "first diff pos", j[:5], a[j[:5]], b[j[:5]])
            if not same_set:
                print("   only oracle", np.setdiff1d(a, b)[:10], "only gpu", np.setdiff1d(b, a)[:10])
print("points with different lists", nb)
m = i1 >= 0
if np.array_equal(i1, i2):
    print("rdiag rel", np.abs(d2[m] - d1[m]).max() / np.abs(d1[m]).max(), "rloc rel", np.abs(l2[m] - l1[m]).max() / np.abs(l1[m]).max())
    r = np.abs(d2[m] - d1[m]) / np.abs(d1[m])
    print("rdiag per-element rel max", r.max())
