"""state_trans / state_trans_inv twins (scale/common/common_scale.f90:1181-1280).

PARITY UNPINNED BY THE REFERENCE: the oracle is pinned against an independent numpy restatement and the
round-trip property; the CUDA kernel against the oracle.  The transform contains a pow(): IEEE add / mul /
div are kept in the reference's order without FMA contraction, the power itself differs by <= 2 ulp between
CUDA and the host libm, hence the 1e-13 relative tolerance (the north_star bar for floating point is 1e-10)."""
import numpy as np
import pytest

import scale_letkf_b200 as sl
from scale_letkf_b200 import capi, synth

TOL = 1e-13


def thermo(pos_q=0, pos_qhyd=0):
    t = capi.Thermo()
    t.Rdry, t.Rvap, t.CVdry, t.PRE00 = 287.04, 461.46, 1004.64 - 287.04, 1.0e5
    for i, v in enumerate([1845.60 - 461.46, 4218.0, 4218.0, 2006.0, 2006.0, 2006.0]):
        t.TRACER_CV[i] = v
    t.POSITIVE_DEFINITE_Q, t.POSITIVE_DEFINITE_QHYD = pos_q, pos_qhyd
    return t


def restart_state(nlev=7, nlon=9, nlat=5, seed=55):
    """(rho, rho u, rho v, rho w, rho theta, qv, qc, qr, qi, qs, qg) member-major, Fortran order."""
    g = synth.rng(seed, 11)
    shp = (nlev, nlon, nlat)
    rho = g.uniform(0.3, 1.25, shp)
    v = np.zeros(shp + (11,), order="F")
    v[..., 0] = rho
    for n in (1, 2, 3):
        v[..., n] = rho * g.normal(0.0, 12.0, shp)
    v[..., 4] = rho * g.uniform(285.0, 420.0, shp)
    v[..., 5] = g.uniform(0.0, 0.02, shp)
    for n in range(6, 11):
        v[..., n] = g.normal(2.0e-4, 3.0e-4, shp)   # some negative hydrometeors: clamp test
    return v


def numpy_state_trans(t, v, inverse):
    v = v.copy(order="F")
    cv = np.array([t.TRACER_CV[i] for i in range(6)])
    if inverse:
        if t.POSITIVE_DEFINITE_Q:
            v[..., 5] = np.maximum(v[..., 5], 0.0)
        if t.POSITIVE_DEFINITE_QHYD:
            v[..., 6:11] = np.maximum(v[..., 6:11], 0.0)
    qdry = np.ones(v.shape[:3])
    cvtot = np.zeros(v.shape[:3])
    for n in range(5, 11):
        qdry = qdry - v[..., n]
        cvtot = cvtot + v[..., n] * cv[n - 5]
    cvtot = t.CVdry * qdry + cvtot
    rtot = t.Rdry * qdry + t.Rvap * v[..., 5]
    if not inverse:
        rho = v[..., 0].copy()
        pres = t.PRE00 * np.power(v[..., 4] * rtot / t.PRE00, (cvtot + rtot) / cvtot)
        temp = pres / (rho * rtot)
        v[..., 0] = v[..., 1] / rho
        v[..., 1] = v[..., 2] / rho
        v[..., 2] = v[..., 3] / rho
        v[..., 3] = temp
        v[..., 4] = pres
    else:
        rho = v[..., 4] / (rtot * v[..., 3])
        rhot = t.PRE00 / rtot * np.power(v[..., 4] / t.PRE00, cvtot / (cvtot + rtot))
        v[..., 4] = rhot
        v[..., 3] = v[..., 2] * rho
        v[..., 2] = v[..., 1] * rho
        v[..., 1] = v[..., 0] * rho
        v[..., 0] = rho
    return v


def rel(a, b):
    sc = np.maximum(np.abs(b).max(axis=(0, 1, 2), keepdims=True), 1e-300)
    return float((np.abs(a - b) / sc).max())


@pytest.mark.parametrize("pos", [(0, 0), (1, 1), (1, 0)])
def test_oracle_state_trans_vs_numpy_and_roundtrip(oracle, pos):
    t = thermo(*pos)
    x = restart_state()
    f = oracle.state_trans(t, x.copy(order="F"), inverse=False)
    assert rel(f, numpy_state_trans(t, x, False)) <= 1e-15
    assert 100.0 < f[..., 3].min() and f[..., 3].max() < 600.0 and f[..., 4].min() > 1.0e3   # plausible T, p
    b = oracle.state_trans(t, f.copy(order="F"), inverse=True)
    assert rel(b, numpy_state_trans(t, f, True)) <= 1e-15
    if pos == (0, 0):
        assert rel(b, x) <= 1e-12            # round trip
    else:
        assert b[..., 5].min() >= 0.0
        if pos[1]:
            assert b[..., 6:11].min() >= 0.0


def test_thermo_defaults_abi():
    import ctypes as C
    lib = capi.load_library()
    t = capi.Thermo()
    lib.letkf_b200_thermo_defaults(C.byref(t))
    d = thermo()
    assert (t.Rdry, t.Rvap, t.CVdry, t.PRE00) == (d.Rdry, d.Rvap, d.CVdry, d.PRE00)
    assert [t.TRACER_CV[i] for i in range(6)] == [d.TRACER_CV[i] for i in range(6)]


@pytest.mark.gpu
@pytest.mark.parametrize("pos", [(0, 0), (1, 1)])
@pytest.mark.parametrize("shape", [(7, 9, 5), (60, 64, 48)])
def test_cuda_state_trans(oracle, pos, shape):
    import torch
    t = thermo(*pos)
    cfg = sl.resolve_config(sl.default_config(MEMBER=4, nlon=shape[1], nlat=shape[2], nlev=shape[0]))
    e = sl.LETKF(cfg, device=0)
    x = restart_state(*shape)
    ref_f = oracle.state_trans(t, x.copy(order="F"), inverse=False)
    got_f = e.state_trans(x.copy(order="F"), t, inverse=False)                    # host buffers
    assert rel(got_f, ref_f) <= TOL
    ref_b = oracle.state_trans(t, ref_f.copy(order="F"), inverse=True)
    d = torch.from_numpy(np.ascontiguousarray(ref_f.transpose(3, 2, 1, 0))).cuda()  # same memory order, on device
    e.state_trans(d, t, inverse=True)
    got_b = d.cpu().numpy().transpose(3, 2, 1, 0)
    assert rel(got_b, ref_b) <= TOL
    if pos == (0, 0):
        assert rel(got_b, x) <= 1e-12
    e.close()


def test_oracle_enssprd_vs_numpy(oracle):
    """enssprd_grd (common_scale.f90:1557-1611) against numpy's two-pass standard deviation."""
    cfg = synth.config_c2(nlon=7, nlat=5, nlev=3, member=9)
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    g = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=8)
    s = oracle.enssprd_grd(9, g)
    ref = np.sqrt(((g[:, :, :9, :] - g[:, :, 9:10, :]) ** 2).sum(axis=2) / 8.0)
    assert np.abs(s - ref).max() / np.abs(ref).max() < 1e-14


@pytest.mark.gpu
def test_cuda_enssprd_bitexact(oracle):
    cfg = synth.config_c2(nlon=16, nlat=12, nlev=5, member=20)
    rig1, rjg1, hgt1 = synth.make_grid(cfg)
    g = synth.make_state(cfg, rig1, rjg1, hgt1, seed_no=9)
    e = sl.LETKF(cfg, device=0)
    assert np.array_equal(e.enssprd_grd(g), oracle.enssprd_grd(20, g))
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("np_", [1, 4])
def test_cuda_fused_pack_unpack(oracle, np_):
    """grd_to_buf with state_trans fused == state_trans then grd_to_buf (bit for bit: same arithmetic), and the pack
    itself equals the oracle's grd_to_buf (common_mpi_scale.f90:1428-1440); likewise buf_to_grd + state_trans_inv.
    Ragged shapes: nlev, nij1max not multiples of the 32 x 32 tile, unequal column counts per rank."""
    import ctypes as C
    import torch
    nlev, nlon, nlat = 37, 9, 7
    cfg = sl.resolve_config(sl.default_config(MEMBER=4, nlon=nlon, nlat=nlat, nlev=nlev))
    e = sl.LETKF(cfg, device=0)
    t = thermo(1, 1)
    x = restart_state(nlev, nlon, nlat, seed=91)
    _, nmax = e.nij1_of(np_, 0)
    nlevall = nlev * 11
    flat = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda()
    # pack
    d_x = flat(x)
    fused = torch.zeros(nmax * nlevall * np_, dtype=torch.float64, device="cuda")
    e.grd_to_buf(np_, d_x, None, fused, thermo=t)
    d_y = flat(x)
    e.state_trans(d_y, t, inverse=False)
    plain = torch.zeros_like(fused)
    e.grd_to_buf(np_, d_y, None, plain)
    assert torch.equal(fused, plain)
    ref = np.zeros((nmax, nlevall, np_), order="F")
    y = d_y.cpu().numpy().reshape(x.shape, order="F").copy(order="F")
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    oracle.lib().oracle_grd_to_buf(nlon, nlat, nlev, 11, 0, np_, P(y), None, P(ref))
    assert np.array_equal(plain.cpu().numpy(), ref.ravel(order="F"))
    # unpack
    back_f = torch.zeros_like(d_x)
    e.buf_to_grd(np_, plain, back_f, None, thermo=t)
    back_p = torch.zeros_like(d_x)
    e.buf_to_grd(np_, plain, back_p, None)
    assert torch.equal(back_p, d_y)
    e.state_trans(back_p, t, inverse=True)
    assert torch.equal(back_f, back_p)
    e.close()
