"""GPU parity tests of the large-ensemble ("tiled") analysis path (scale_letkf_b200/csrc/tiled.cuh):
whole-GPU batched DMMA GEMMs + Newton-Schulz over HBM-resident matrices, in the primal (k x k) and
the low-rank dual (p x p) form, against the CPU oracle (EISPACK eigensolve restatement of
common/common_letkf.f90:52-257 inside scale/letkf/letkf_tools.f90:50-932).  Same bar as the small-
ensemble path: local-observation counts bit-exact, analysis within 1e-10 relative."""
import pytest

from helpers import sonde_case, radar_case
from test_gpu_parity import _das_compare

pytestmark = pytest.mark.gpu

FORMS = ["primal", "dual"]


def _force(monkeypatch, form):
    monkeypatch.setenv("LETKF_B200_SOLVER", "tiled")
    monkeypatch.setenv("LETKF_B200_TILED_FORM", form)


@pytest.mark.parametrize("form", FORMS)
@pytest.mark.parametrize("relax", ["rtps", "rtpp", "none"])
def test_tiled_forced_sonde(oracle, monkeypatch, form, relax):
    """The tiled kernels on a small ensemble (forced), where p > k and p < k points both occur."""
    _force(monkeypatch, form)
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=20, nsonde=30, nsfc=100)
    cfg.RELAX_ALPHA_SPREAD = 0.95 if relax == "rtps" else 0.0
    cfg.RELAX_ALPHA = 0.7 if relax == "rtpp" else 0.0
    cfg.BOUNDARY_BUFFER_WIDTH = 45.0e3
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0 and out["launches"] > 10


@pytest.mark.parametrize("form", FORMS)
def test_tiled_forced_det_qtop_qsprd_inflated_prior(oracle, monkeypatch, form):
    _force(monkeypatch, form)
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=12, nsonde=30, nsfc=100, det=True)
    cfg.Q_UPDATE_TOP = 300.0e2
    cfg.Q_SPRD_MAX = 0.05
    cfg.INFL_MUL = 1.15
    cfg.RELAX_TO_INFLATED_PRIOR = 1
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)


@pytest.mark.parametrize("form", FORMS)
def test_tiled_forced_adaptive_inflation(oracle, monkeypatch, form):
    _force(monkeypatch, form)
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=10, nsonde=30, nsfc=100)
    cfg.INFL_MUL_ADAPTIVE = 1
    cfg.INFL_MUL = 1.05
    cfg.INFL_MUL_MIN = 1.02
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle, infl3d=True)


@pytest.mark.parametrize("form", FORMS)
def test_tiled_forced_variable_localisation_groups(oracle, monkeypatch, form):
    _force(monkeypatch, form)
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=10, nsonde=30, nsfc=100)
    for n in range(11):
        cfg.VAR_LOCAL[2][n] = 1.0 if n >= 5 else 0.0
        cfg.VAR_LOCAL[3][n] = 0.5 if n == 4 else 1.0
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)


@pytest.mark.parametrize("form", FORMS)
def test_tiled_forced_radar(oracle, monkeypatch, form):
    _force(monkeypatch, form)
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=50, max_nobs=100, nlon=32, nlat=32, nlev=6, det=True)
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert 0 < out["nsolved"] < out["npoints"]


def test_tiled_small_batches(oracle, monkeypatch):
    """A tiny scratch budget forces many batches (batch boundaries inside a level)."""
    _force(monkeypatch, "primal")
    monkeypatch.setenv("LETKF_B200_TILED_MB", "0.01")
    cfg, rig1, rjg1, hgt1, obs, gues = sonde_case(member=20, nsonde=30, nsfc=100)
    _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)


@pytest.mark.parametrize("member,max_nobs", [(136, 40), (200, 300)])
def test_tiled_natural_radar(oracle, member, max_nobs):
    """MEMBER > 102 takes the tiled path by itself: (136, 40) -> dual form (p <= 80 < 0.6 k),
    (200, 300) -> primal form (p up to 600 > k)."""
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=member, max_nobs=max_nobs, nlon=20, nlat=20, nlev=4,
                                                  radius=4.0e3)
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0 and out["launches"] > 10


def test_tiled_k1000(oracle):
    """BASELINE config C4 shape: 1000 members, MAX_NOBS_PER_GRID(22) = 100 (p <= 200), a handful of points."""
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=1000, max_nobs=100, nlon=6, nlat=6, nlev=2, radius=1.5e3)
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0


def test_tiled_k1000_primal(oracle, monkeypatch):
    monkeypatch.setenv("LETKF_B200_TILED_FORM", "primal")
    cfg, rig1, rjg1, hgt1, obs, gues = radar_case(member=1000, max_nobs=100, nlon=4, nlat=4, nlev=2, radius=1.5e3)
    out, ref = _das_compare(cfg, rig1, rjg1, hgt1, obs, gues, oracle)
    assert out["nsolved"] > 0


# ---- letkf_core twin for ne > 128 (tiled) ------------------------------------------------------------
import numpy as np
import scale_letkf_b200 as sl
from scale_letkf_b200 import synth
from helpers import TOL, relerr


@pytest.fixture(scope="module")
def core_engine():
    cfg = sl.default_config(MEMBER=20, nlon=8, nlat=8, nlev=2)
    e = sl.LETKF(sl.resolve_config(cfg), device=0)
    yield e
    e.close()


@pytest.mark.parametrize("ne,nobs,npts", [(136, 200, 12), (200, 60, 8), (1000, 120, 3)])
def test_core_batch_tiled_sizes(core_engine, oracle, ne, nobs, npts):
    """letkf_core (common/common_letkf.f90:52-257) with more than 128 members: trans, transm, pao, transmd."""
    c = synth.make_core_batch(ne=ne, npts=npts, nobs=nobs, seed_no=300 + ne, det=True, infl=1.07)
    args = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"], c["rdiag"], c["rloc"], c["dep"], c["parm_infl"])
    ref = oracle.core_batch(*args, depd=c["depd"])
    r = core_engine.letkf_core(*args, depd=c["depd"])
    for key in ("trans", "transm", "pao", "transmd"):
        assert relerr(r[key], ref[key]) <= TOL, key
    if (c["nobsl"] == 0).any():
        i = int(np.argmax(c["nobsl"] == 0))
        assert np.array_equal(r["trans"][i], np.sqrt(1.07) * np.eye(ne))


def test_core_batch_tiled_options(core_engine, oracle):
    c = synth.make_core_batch(ne=150, npts=10, nobs=50, seed_no=33, infl=1.1)
    base = (c["ne"], c["nobs"], c["nobsl"], c["hdxb"])
    err2 = c["rdiag"] * c["rloc"]
    ref = oracle.core_batch(*base, err2, c["rloc"], c["dep"], c["parm_infl"], rdiag_wloc=False, infl_update=True)
    r = core_engine.letkf_core(*base, err2, c["rloc"], c["dep"], c["parm_infl"], rdiag_wloc=False, infl_update=True)
    for key in ("trans", "transm", "pao"):
        assert relerr(r[key], ref[key]) <= TOL, key
    sel = c["nobsl"] > 0
    assert relerr(r["parm_infl"][sel], ref["parm_infl"][sel]) <= TOL
    assert np.array_equal(r["parm_infl"][~sel], ref["parm_infl"][~sel])
    ref = oracle.core_batch(*base, c["rdiag"], c["rloc"], c["dep"], c["parm_infl"], want_transm=False, want_pao=False)
    r = core_engine.letkf_core(*base, c["rdiag"], c["rloc"], c["dep"], c["parm_infl"], want_transm=False, want_pao=False)
    assert relerr(r["trans"], ref["trans"]) <= TOL
