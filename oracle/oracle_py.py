"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by the product path.
PARITY UNPINNED by the reference (see oracle/letkf_oracle.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from scale_letkf_b200 import capi  # noqa: E402  (struct definitions of the C ABI only)

LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in
            ("oracle_core.cpp", "oracle_das.cpp", "letkf_oracle.h", "../include/letkf_b200.h")]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs)):
        return LIB
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        vp, i = C.c_void_p, C.c_int
        L.oracle_pythag.restype = C.c_double
        L.oracle_pythag.argtypes = [C.c_double, C.c_double]
        L.oracle_rs.argtypes = [i, i, vp, vp, vp]
        L.oracle_mtx_eigen.argtypes = [i, vp, vp, vp]
        L.oracle_letkf_core.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, i, i, vp, vp]
        L.oracle_core_batch.argtypes = [i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, vp, vp, i]
        L.oracle_quickselect_arg.restype = None
        L.oracle_quickselect_arg.argtypes = [vp, vp, i, i, i]
        L.oracle_quickselect_desc_arg.restype = None
        L.oracle_quickselect_desc_arg.argtypes = [vp, vp, i, i, i]
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [C.POINTER(capi.Config)]
        L.oracle_destroy.restype = None
        L.oracle_destroy.argtypes = [vp]
        L.oracle_set_quirks.restype = None
        L.oracle_set_quirks.argtypes = [vp, i]
        L.oracle_set_obs.argtypes = [vp, C.POINTER(capi.Obs)]
        L.oracle_set_grid.argtypes = [vp, i, vp, vp, vp]
        L.oracle_obs_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.oracle_get_ctype.argtypes = [vp, i, C.POINTER(capi.CtypeInfo)]
        L.oracle_get_ac_ext.argtypes = [vp, i, vp]
        L.oracle_get_sorted_index.argtypes = [vp, vp]
        L.oracle_obs_local.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, vp, vp, i, i]
        L.oracle_das_letkf.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i,
                                       C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.oracle_ensmean_grd.restype = None
        L.oracle_ensmean_grd.argtypes = [i, i, i, i, i, i, vp, vp]
        L.oracle_nij1.restype = None
        L.oracle_nij1.argtypes = [i, i, i, i, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        for name in ("oracle_grd_to_buf", "oracle_buf_to_grd"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [i, i, i, i, i, i, vp, vp, vp]
        for name in ("oracle_buf_to_ens", "oracle_ens_to_buf"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [i, i, i, i, i, i, i, i, i, i, vp, vp, vp]
        L.oracle_obs_departure_qc.restype = None
        L.oracle_obs_departure_qc.argtypes = [C.POINTER(capi.QcConfig), i, i, i, i, vp, vp, vp, vp, vp, vp]
        L.oracle_enssprd_grd.restype = None
        L.oracle_enssprd_grd.argtypes = [i, i, i, i, i, vp, vp]
        L.oracle_state_trans.restype = None
        L.oracle_state_trans.argtypes = [C.POINTER(capi.Thermo), i, i, i, i, i, i, vp]
        L.oracle_obsope_radar.restype = None
        L.oracle_obsope_radar.argtypes = [C.POINTER(capi.RadarConfig), i, vp, vp, vp, vp, vp, vp, vp, i, C.POINTER(vp), i, vp, vp]
        L.oracle_additive_inflation.restype = None
        d = C.c_double
        L.oracle_additive_inflation.argtypes = [i, i, i, i, i, i, d, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, i, vp, vp, d, d, d, d, vp]
        L.oracle_nobs_out.restype = None
        L.oracle_nobs_out.argtypes = [vp, i, vp, vp, vp, i]
        L.oracle_monit_dep.restype = None
        L.oracle_monit_dep.argtypes = [i, vp, vp, vp, vp, vp, vp]
        L.oracle_max_threads.restype = i
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rs(a):
    """EISPACK rs on a symmetric matrix -> (w ascending, z) ; a is (n,n)."""
    n = a.shape[0]
    af = np.asfortranarray(a, dtype=np.float64)
    w = np.zeros(n)
    z = np.zeros((n, n), order="F")
    ierr = lib().oracle_rs(n, n, _p(af), _p(w), _p(z))
    return ierr, w, z


def mtx_eigen(a):
    n = a.shape[0]
    af = np.asfortranarray(a, dtype=np.float64)
    w = np.zeros(n)
    z = np.zeros((n, n), order="F")
    nrank = lib().oracle_mtx_eigen(n, _p(af), _p(w), _p(z))
    return nrank, w, z


def letkf_core(hdxb, rdiag, rloc, dep, parm_infl, nobsl=None, rdiag_wloc=False, infl_update=False,
               depd=None, want_transm=True, want_pao=True):
    """One reference-style call.  hdxb is (nobs, ne) (any order; copied to Fortran order)."""
    nobs, ne = hdxb.shape
    if nobsl is None:
        nobsl = nobs
    h = np.asfortranarray(hdxb, dtype=np.float64)
    infl = np.array([parm_infl], dtype=np.float64)
    trans = np.zeros((ne, ne), order="F")
    transm = np.zeros(ne) if want_transm else None
    pao = np.zeros((ne, ne), order="F") if want_pao else None
    transmd = np.zeros(ne) if depd is not None else None
    r = lib().oracle_letkf_core(ne, max(nobs, 1), nobsl, _p(h), _p(_f64(rdiag)), _p(_f64(rloc)),
                                _p(_f64(dep)), _p(infl), _p(trans), _p(transm), _p(pao),
                                int(rdiag_wloc), int(infl_update),
                                _p(_f64(depd)) if depd is not None else None, _p(transmd))
    return dict(status=r, trans=trans, transm=transm, pao=pao, transmd=transmd, parm_infl=infl[0])


def core_batch(ne, nobs, nobsl, hdxb, rdiag, rloc, dep, parm_infl, rdiag_wloc=True,
               infl_update=False, depd=None, want_transm=True, want_pao=True, nthreads=0):
    """Batched layout of letkf_b200_core_batch: hdxb (npts, ne, nobs) C-order == per-point
    column-major (nobs, ne)."""
    npts = len(nobsl)
    trans = np.zeros((npts, ne, ne))
    transm = np.zeros((npts, ne)) if want_transm else None
    pao = np.zeros((npts, ne, ne)) if want_pao else None
    transmd = np.zeros((npts, ne)) if depd is not None else None
    infl = np.array(parm_infl, dtype=np.float64).copy()
    r = lib().oracle_core_batch(ne, nobs, npts, _p(np.ascontiguousarray(nobsl, dtype=np.int32)),
                                _p(_f64(hdxb)), _p(_f64(rdiag)), _p(_f64(rloc)), _p(_f64(dep)),
                                _p(infl), _p(trans), _p(transm), _p(pao), int(rdiag_wloc),
                                int(infl_update), _p(_f64(depd)) if depd is not None else None,
                                _p(transmd), nthreads)
    return dict(status=r, trans=trans, transm=transm, pao=pao, transmd=transmd, parm_infl=infl)


def quickselect_arg(A, X, left, right, K, desc=False):
    """A: float64 keys; X: int32 1-based indices into A (modified in place)."""
    fn = lib().oracle_quickselect_desc_arg if desc else lib().oracle_quickselect_arg
    fn(_p(A), _p(X), left, right, K)


def make_obs_struct(obs):
    """obs: dict of numpy arrays (elm, typ int32; ri, rj, lev, dat, err, val float64;
    ensval (nobs, nensobs) float64 C-order).  Returns (struct, keepalive)."""
    keep = {
        "elm": np.ascontiguousarray(obs["elm"], dtype=np.int32),
        "typ": np.ascontiguousarray(obs["typ"], dtype=np.int32),
    }
    for kf in ("ri", "rj", "lev", "dat", "err", "val", "ensval"):
        keep[kf] = _f64(obs[kf])
    o = capi.Obs()
    o.nobs = keep["elm"].shape[0]
    o.nensobs = keep["ensval"].shape[1] if keep["ensval"].ndim == 2 else 0
    for kf, arr in keep.items():
        setattr(o, kf, arr.ctypes.data)
    return o, keep


class Oracle:
    """Stateful twin of the module state das_letkf reads."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().oracle_create(C.byref(cfg))
        self._keep = None

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def set_quirks(self, ij_obsgrd_quirk):
        lib().oracle_set_quirks(self.h, int(ij_obsgrd_quirk))

    def set_obs(self, obs):
        o, keep = make_obs_struct(obs)
        r = lib().oracle_set_obs(self.h, C.byref(o))
        if r != 0:
            raise RuntimeError(f"oracle_set_obs failed: {r}")

    def set_grid(self, rig1, rjg1, hgt1):
        self.nij1 = len(rig1)
        hg = np.asfortranarray(hgt1, dtype=np.float64)   # (nij1, nlev)
        lib().oracle_set_grid(self.h, self.nij1, _p(_f64(rig1)), _p(_f64(rjg1)), _p(hg))

    def obs_info(self):
        a, b = C.c_int32(), C.c_int32()
        lib().oracle_obs_info(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def ctype(self, ic):
        info = capi.CtypeInfo()
        if lib().oracle_get_ctype(self.h, ic, C.byref(info)) != 0:
            raise IndexError(ic)
        return info

    def ac_ext(self, ic):
        info = self.ctype(ic)
        a = np.zeros((info.ngrdext_j, info.ngrdext_i + 1), dtype=np.int32)
        lib().oracle_get_ac_ext(self.h, ic, _p(a))
        return a

    def sorted_index(self):
        n, _ = self.obs_info()
        a = np.zeros(n, dtype=np.int32)
        lib().oracle_get_sorted_index(self.h, _p(a))
        return a

    def obs_local(self, ri, rj, rlev, rz, nvar, max_out, brute=False):
        npts = len(ri)
        nobsl = np.zeros(npts, dtype=np.int32)
        idx = np.full((npts, max_out), -1, dtype=np.int32)
        rdiag = np.zeros((npts, max_out))
        rloc = np.zeros((npts, max_out))
        r = lib().oracle_obs_local(self.h, npts, _p(_f64(ri)), _p(_f64(rj)), _p(_f64(rlev)),
                                   _p(_f64(rz)), nvar, _p(nobsl), _p(idx), _p(rdiag), _p(rloc),
                                   max_out, int(brute))
        if r != 0:
            raise RuntimeError("oracle_obs_local: max_out too small")
        return nobsl, idx, rdiag, rloc

    def nobs_out(self, nvar, pmean, nthreads=0):
        """NOBS_OUT fields of das_letkf (letkf_tools.f90:440-447, 767-778) for model variable nvar (the reference: iv3d_t).
        pmean: F-order (nij1, nlev) mean pressure.  -> out (nij1, nlev, 11) F-order, exact_hits (nij1, nlev)"""
        pm = np.asfortranarray(pmean, dtype=np.float64)
        nij1, nlev = pm.shape
        out = np.zeros((nij1, nlev, 11), order="F")
        hits = np.zeros((nij1, nlev), dtype=np.int32, order="F")
        lib().oracle_nobs_out(self.h, int(nvar), _p(pm), _p(out), _p(hits), int(nthreads))
        return out, hits

    def das_letkf(self, gues3d, gues2d=None, infl3d=None, want_rtps=False, want_nobsl=False,
                  point_mask=None, nthreads=0):
        """gues3d: Fortran-ordered (nij1, nlev, nens, nv3d), modified in place."""
        assert gues3d.flags.f_contiguous
        anal3d = np.zeros_like(gues3d, order="F")
        anal2d = np.zeros_like(gues2d, order="F") if gues2d is not None else None
        nij1, nlev = gues3d.shape[0], gues3d.shape[1]
        rtps = np.zeros((nij1, nlev, gues3d.shape[3]), order="F") if want_rtps else None
        nobsl = np.zeros((nij1, nlev), dtype=np.int32, order="F") if want_nobsl else None
        npts, nsolved = C.c_int64(), C.c_int64()
        pm = None
        if point_mask is not None:
            pm = np.asfortranarray(point_mask, dtype=np.uint8)
        r = lib().oracle_das_letkf(self.h, _p(gues3d), _p(gues2d), _p(anal3d), _p(anal2d),
                                   _p(infl3d), _p(rtps), _p(nobsl), _p(pm), nthreads,
                                   C.byref(npts), C.byref(nsolved))
        return dict(status=r, anal3d=anal3d, anal2d=anal2d, rtps=rtps, nobsl=nobsl,
                    npoints=npts.value, nsolved=nsolved.value)


def ensmean_grd(mem, v3d, v2d=None):
    nij, nlev, nens, nv3d = v3d.shape
    nv2d = v2d.shape[2] if v2d is not None else 0
    lib().oracle_ensmean_grd(mem, nens, nij, nlev, nv3d, nv2d, _p(v3d), _p(v2d))


def nij1(nlon, nlat, np_, rank):
    a, b = C.c_int32(), C.c_int32()
    lib().oracle_nij1(nlon, nlat, np_, rank, C.byref(a), C.byref(b))
    return a.value, b.value


def max_threads():
    return lib().oracle_max_threads()


def obs_departure_qc(qcfg, member, det, elm, dat, err, qc, ensval):
    """scale/letkf/letkf_obs.f90:355-560; returns (qc, val, ensval) as new arrays."""
    elm = np.ascontiguousarray(elm, dtype=np.int32)
    qc = np.array(qc, dtype=np.int32, order="C")
    ens = np.array(ensval, dtype=np.float64, order="C")
    val = np.zeros(len(elm))
    lib().oracle_obs_departure_qc(C.byref(qcfg), member, int(det), len(elm), ens.shape[1], _p(elm), _p(_f64(dat)),
                                  _p(_f64(err)), _p(qc), _p(ens), _p(val))
    return qc, val, ens


def state_trans(thermo, v3dg, inverse=False, iv3d_q=6):
    """In place on v3dg (nlev, nlon, nlat, nv3d) Fortran order (common_scale.f90:1181-1280)."""
    assert v3dg.flags.f_contiguous
    nlev, nlon, nlat, nv3d = v3dg.shape
    lib().oracle_state_trans(C.byref(thermo), int(bool(inverse)), nlev, nlon, nlat, nv3d, iv3d_q, _p(v3dg))
    return v3dg


def enssprd_grd(mem, v3d):
    nij, nlev, nens, nv3d = v3d.shape
    out = np.zeros((nij, nlev, nv3d), order="F")
    lib().oracle_enssprd_grd(mem, nens, nij, nlev, nv3d, _p(v3d), _p(out))
    return out


def additive_inflation(cfg, addi3d, anal3d, infl_add, gues3d=None, q_ratio=False, ref_only=False, ishuf=None, addi2d=None,
                       anal2d=None, rig1=None, rjg1=None, ref_ri=None, ref_rj=None, hloc=1.0):
    """letkf_tools.f90:804-929 on F-order arrays shaped like gues3d (nij, nlev, nens, nv3d); addi* become perturbations and
    anal* are updated in place.  ref_ri/ref_rj: positions of the (REF, PHARAD) observations.  Returns addinfl_weight."""
    nij, nlev, nens, nv3d = addi3d.shape
    nv2d = 0 if addi2d is None else addi2d.shape[2]
    sh = None if ishuf is None else np.ascontiguousarray(ishuf, dtype=np.int32)
    w = np.zeros(nij)
    rr = None if ref_ri is None else _f64(ref_ri)
    rj = None if ref_rj is None else _f64(ref_rj)
    lib().oracle_additive_inflation(cfg.MEMBER, nens, nij, nlev, nv3d, nv2d, float(infl_add), int(bool(q_ratio)), int(bool(ref_only)),
                                    cfg.iv3d_q, cfg.iv3d_qg, _p(sh), _p(addi3d), _p(addi2d), _p(gues3d), _p(anal3d), _p(anal2d),
                                    _p(None if rig1 is None else _f64(rig1)), _p(None if rjg1 is None else _f64(rjg1)),
                                    0 if rr is None else len(rr), _p(rr), _p(rj), float(hloc), float(cfg.DX), float(cfg.DY),
                                    float(cfg.dist_zero_fac_square), _p(w))
    return w


def monit_dep(elm, dep, qc):
    """common_obs_scale.f90:1851-1895 -> (nobs[16], bias[16], rmse[16])."""
    elm = np.ascontiguousarray(elm, dtype=np.int32)
    qc = np.ascontiguousarray(qc, dtype=np.int32)
    n, b, r = np.zeros(16, dtype=np.int32), np.zeros(16), np.zeros(16)
    lib().oracle_monit_dep(len(elm), _p(elm), _p(_f64(dep)), _p(qc), _p(n), _p(b), _p(r))
    return n, b, r


def obsope_conv(ccfg, elm, ril, rjl, lev, grids3, grids2, rotc=None):
    """oracle_conv.cpp: conventional observation operator (phys2ijk + Trans_XtoY) for all members; grids3 / grids2 = lists of
    F-order v3dgh(nlevh,nlonh,nlath,nv3dd) / v2dgh(nlonh,nlath,nv2dd)"""
    nobs, nmem = len(elm), len(grids3)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    elm = np.ascontiguousarray(elm, dtype=np.int32)
    ril, rjl, lev = f(ril), f(rjl), f(lev)
    rotc = None if rotc is None else f(rotc)
    p3 = (C.c_void_p * nmem)(*[g.ctypes.data for g in grids3])
    p2 = (C.c_void_p * nmem)(*[g.ctypes.data for g in grids2])
    y = np.zeros((nobs, nmem))
    q = np.zeros((nobs, nmem), dtype=np.int32)
    L = lib()
    L.oracle_obsope_conv.restype = None
    L.oracle_obsope_conv(C.byref(ccfg), nobs, _p(elm), _p(ril), _p(rjl), _p(lev), _p(rotc), nmem, p3, p2, nmem, _p(y), _p(q))
    return y, q


def monit_obs(sets, v3dgh, v2dgh, t_range=0.0):
    """monit_obs (common_obs_scale.f90:1370-1844) over a list of observation sets (dicts as scale_letkf_b200.LETKF.monit_obs takes
    them): per-set observation loop (oracle_conv.cpp) + monit_dep over the concatenation."""
    f = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    L = lib()
    L.oracle_monit_obs_set.restype = None
    L.oracle_monit_obs_set.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 9 + [C.c_double] + [C.c_void_p] * 4
    elms, ohxs, oqcs = [], [], []
    for st in sets:
        cfg = st["cfg"]
        conv = isinstance(cfg, capi.ConvConfig)
        elm = np.ascontiguousarray(st["elm"], dtype=np.int32)
        n = len(elm)
        ohx, oqc = np.zeros(n), np.zeros(n, dtype=np.int32)
        cp = C.cast(C.pointer(cfg), C.c_void_p)
        L.oracle_monit_obs_set(cp if conv else None, None if conv else cp, n, _p(elm), _p(f(st["ril"])), _p(f(st["rjl"])),
                               _p(f(st.get("lon"))), _p(f(st.get("lat"))), _p(f(st["lev"])), _p(f(st["dat"])), _p(f(st.get("dif"))),
                               _p(f(st.get("rotc"))), float(t_range), _p(v3dgh), _p(v2dgh), _p(ohx), _p(oqc))
        elms.append(elm); ohxs.append(ohx); oqcs.append(oqc)
    elm, ohx, oqc = np.concatenate(elms), np.concatenate(ohxs), np.concatenate(oqcs)
    nobs, bias, rmse = monit_dep(elm, ohx, oqc)
    return dict(nobs=nobs, bias=bias, rmse=rmse, ohx=ohx, oqc=oqc, elm=elm)


def obsope_radar(rcfg, elm, ril, rjl, lon, lat, lev, grids, rotc=None):
    """oracle_radar.cpp: radar observation operator for all members; grids = list of F-order v3dg(nlevh,nlonh,nlath,nv3dd)"""
    nobs, nmem = len(elm), len(grids)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    elm = np.ascontiguousarray(elm, dtype=np.int32)
    ril, rjl, lon, lat, lev = f(ril), f(rjl), f(lon), f(lat), f(lev)
    rotc = None if rotc is None else f(rotc)
    ptrs = (C.c_void_p * nmem)(*[g.ctypes.data for g in grids])
    y = np.zeros((nobs, nmem))
    q = np.zeros((nobs, nmem), dtype=np.int32)
    lib().oracle_obsope_radar(C.byref(rcfg), nobs, _p(elm), _p(ril), _p(rjl), _p(lon), _p(lat), _p(lev), _p(rotc), nmem, ptrs,
                              nmem, _p(y), _p(q))
    return y, q
