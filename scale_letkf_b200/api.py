"""Host-side mirror of the reference's operator interface for the analysis path.

`LETKF` carries the module state the Fortran `das_letkf` reads (grid coordinates,
bucket-sorted observation tables, namelist scalars) and exposes the reference's entry
points under their own names:

    letkf_core(...)      common/common_letkf.f90:52      (batched)
    set_letkf_obs(obs)   scale/letkf/letkf_obs.f90:78    (bucket-sort half)
    set_common_mpi_grid  scale/common/common_mpi_scale.f90:244 (rig1, rjg1, hgt1)
    obs_local(...)       scale/letkf/letkf_tools.f90:1325
    das_letkf(...)       scale/letkf/letkf_tools.f90:50
    ensmean_grd(...)     scale/common/common_scale.f90:1513

Every call goes through the C ABI of include/letkf_b200.h into CUDA kernels; numpy arrays
are host buffers (copied H2D/D2H inside the call), torch CUDA tensors are used in place.
There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import capi

_ERR = {capi.EINVAL: "EINVAL", capi.ECUDA: "ECUDA", capi.ESTATE: "ESTATE", capi.EEIGEN: "EEIGEN",
        capi.ENOMEM: "ENOMEM"}


class LetkfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"letkf_b200 error {_ERR.get(code, code)}: {msg}")
        self.code = code


def _is_torch(x):
    return x is not None and type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return x.ctypes.data_as(C.c_void_p)


class LETKF:
    def __init__(self, cfg, device=0):
        self.lib = capi.load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        r = self.lib.letkf_b200_create(C.byref(cfg), int(device), C.byref(self.h))
        if r != 0:
            self.h = None
            raise LetkfError(r, "letkf_b200_create failed (no usable CUDA device?)")
        self.device = device
        self.nij1 = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.letkf_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r, allow=()):
        if r != 0 and r not in allow:
            raise LetkfError(r, self.lib.letkf_b200_last_error(self.h).decode())
        return r

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.letkf_b200_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # ---- letkf_core ---------------------------------------------------------------------
    def letkf_core(self, ne, nobs, nobsl, hdxb, rdiag, rloc, dep, parm_infl, rdiag_wloc=True,
                   infl_update=False, depd=None, want_transm=True, want_pao=True):
        """Batched twin of letkf_core.  hdxb: (npts, ne, nobs) C-order (= per point the
        Fortran hdxb(nobs, ne)); returns dict(trans (npts, ne, ne) column-major per point, ...)."""
        npts = len(nobsl)
        nobsl = np.ascontiguousarray(nobsl, dtype=np.int32)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        hdxb, rdiag, rloc, dep = f(hdxb), f(rdiag), f(rloc), f(dep)
        infl = np.array(parm_infl, dtype=np.float64).copy()
        trans = np.zeros((npts, ne, ne))
        transm = np.zeros((npts, ne)) if want_transm else None
        pao = np.zeros((npts, ne, ne)) if want_pao else None
        depd_ = f(depd) if depd is not None else None
        transmd = np.zeros((npts, ne)) if depd is not None else None
        r = self.lib.letkf_b200_core_batch(self.h, ne, nobs, npts, _ptr(nobsl), _ptr(hdxb), _ptr(rdiag),
                                           _ptr(rloc), _ptr(dep), _ptr(infl), _ptr(trans), _ptr(transm),
                                           _ptr(pao), int(rdiag_wloc), int(infl_update), _ptr(depd_),
                                           _ptr(transmd), capi.MEM_HOST)
        self._ck(r)
        return dict(status=r, trans=trans, transm=transm, pao=pao, transmd=transmd, parm_infl=infl)

    # ---- module state ---------------------------------------------------------------------
    def set_common_mpi_grid(self, rig1, rjg1, hgt1):
        """rig1, rjg1 (nij1,), hgt1 (nij1, nlev) Fortran order -- numpy or torch CUDA."""
        self.nij1 = int(rig1.shape[0])
        if _is_torch(rig1):
            space = capi.MEM_DEVICE
        else:
            space = capi.MEM_HOST
            rig1 = np.ascontiguousarray(rig1, dtype=np.float64)
            rjg1 = np.ascontiguousarray(rjg1, dtype=np.float64)
            hgt1 = np.asfortranarray(hgt1, dtype=np.float64)
        self._ck(self.lib.letkf_b200_set_grid(self.h, self.nij1, _ptr(rig1), _ptr(rjg1), _ptr(hgt1), space))

    def set_letkf_obs(self, obs):
        """obs: dict(elm, typ, ri, rj, lev, dat, err, val, ensval (nobs, nensobs))."""
        keep = {"elm": np.ascontiguousarray(obs["elm"], dtype=np.int32),
                "typ": np.ascontiguousarray(obs["typ"], dtype=np.int32)}
        for kf in ("ri", "rj", "lev", "dat", "err", "val", "ensval"):
            keep[kf] = np.ascontiguousarray(obs[kf], dtype=np.float64)
        o = capi.Obs()
        o.nobs = keep["elm"].shape[0]
        o.nensobs = keep["ensval"].shape[1] if keep["ensval"].ndim == 2 else 0
        for kf, arr in keep.items():
            setattr(o, kf, arr.ctypes.data)
        self._ck(self.lib.letkf_b200_set_obs(self.h, C.byref(o)))

    def obs_departure_qc(self, elm, dat, err, qc, ensval, qcfg=None):
        """Departure + QC half of set_letkf_obs (scale/letkf/letkf_obs.f90:355-560) on the device.
        ensval (nobs, nensobs): H(x_m) per member on entry -> perturbations on exit; returns (qc, val,
        ensval) as new arrays.  Observations with qc == 0 afterwards go to set_letkf_obs."""
        if qcfg is None:
            qcfg = capi.QcConfig()
            self.lib.letkf_b200_qc_config_defaults(C.byref(qcfg))
        elm = np.ascontiguousarray(elm, dtype=np.int32)
        dat = np.ascontiguousarray(dat, dtype=np.float64)
        err = np.ascontiguousarray(err, dtype=np.float64)
        qc = np.array(qc, dtype=np.int32, order="C")
        ens = np.array(ensval, dtype=np.float64, order="C")
        val = np.zeros(len(elm))
        self._ck(self.lib.letkf_b200_obs_departure_qc(self.h, C.byref(qcfg), len(elm), ens.shape[1], _ptr(elm), _ptr(dat),
                                                      _ptr(err), _ptr(qc), _ptr(ens), _ptr(val), capi.MEM_HOST))
        return qc, val, ens

    def monit_dep(self, elm, dep, qc):
        """monit_dep (scale/common/common_obs_scale.f90:1851-1895): (nobs[16], bias[16], rmse[16]) per element uid."""
        elm = np.ascontiguousarray(elm, dtype=np.int32)
        dep = np.ascontiguousarray(dep, dtype=np.float64)
        qc = np.ascontiguousarray(qc, dtype=np.int32)
        n, b, r = np.zeros(16, dtype=np.int32), np.zeros(16), np.zeros(16)
        self._ck(self.lib.letkf_b200_monit_dep(self.h, len(elm), _ptr(elm), _ptr(dep), _ptr(qc), _ptr(n), _ptr(b), _ptr(r),
                                               capi.MEM_HOST))
        return n, b, r

    def set_letkf_obs_raw(self, raw, qcfg=None, qc_in=None):
        """The whole of set_letkf_obs (scale/letkf/letkf_obs.f90:78) for observations still carrying H(x_m):
        departure + QC on the device (:355-560), departure statistics of the accepted observations (monit_dep,
        :577-586), then the bucket sort of the qc == 0 observations (:660-976).  `raw`: dict(elm, typ, ri, rj, lev,
        dat, err, ensval (nobs, nensobs) = H(x_m)).  Returns dict(qc, val, nobs, bias, rmse, kept)."""
        n = len(raw["elm"])
        qc0 = np.zeros(n, dtype=np.int32) if qc_in is None else qc_in
        qc, val, ens = self.obs_departure_qc(raw["elm"], raw["dat"], raw["err"], qc0, raw["ensval"], qcfg)
        cnt, bias, rmse = self.monit_dep(raw["elm"], val, qc)
        keep = qc == 0
        obs = {kf: np.ascontiguousarray(np.asarray(raw[kf])[keep]) for kf in ("elm", "typ", "ri", "rj", "lev", "dat", "err")}
        obs["val"] = np.ascontiguousarray(val[keep])
        obs["ensval"] = np.ascontiguousarray(ens[keep])
        self._obs_keepalive = obs   # (set_obs may page-lock the ensemble table)
        self.set_letkf_obs(obs)
        return dict(qc=qc, val=val, nobs=cnt, bias=bias, rmse=rmse, kept=int(keep.sum()))

    def obs_info(self):
        a, b = C.c_int32(), C.c_int32()
        self._ck(self.lib.letkf_b200_obs_info(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def ctype(self, ic):
        info = capi.CtypeInfo()
        self._ck(self.lib.letkf_b200_get_ctype(self.h, ic, C.byref(info)))
        return info

    def ac_ext(self, ic):
        info = self.ctype(ic)
        a = np.zeros((info.ngrdext_j, info.ngrdext_i + 1), dtype=np.int32)
        self._ck(self.lib.letkf_b200_get_ac_ext(self.h, ic, _ptr(a)))
        return a

    def sorted_index(self):
        n, _ = self.obs_info()
        a = np.zeros(n, dtype=np.int32)
        self._ck(self.lib.letkf_b200_get_sorted_index(self.h, _ptr(a)))
        return a

    # ---- obs_local --------------------------------------------------------------------------
    def obs_local(self, ri, rj, rlev, rz, nvar, max_out):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        ri, rj, rlev, rz = f(ri), f(rj), f(rlev), f(rz)
        npts = len(ri)
        nobsl = np.zeros(npts, dtype=np.int32)
        idx = np.full((npts, max_out), -1, dtype=np.int32)
        rdiag = np.zeros((npts, max_out))
        rloc = np.zeros((npts, max_out))
        self._ck(self.lib.letkf_b200_obs_local(self.h, npts, _ptr(ri), _ptr(rj), _ptr(rlev), _ptr(rz), nvar,
                                               _ptr(nobsl), _ptr(idx), _ptr(rdiag), _ptr(rloc), max_out,
                                               capi.MEM_HOST))
        return nobsl, idx, rdiag, rloc

    # ---- das_letkf --------------------------------------------------------------------------
    def das_letkf(self, gues3d, anal3d=None, gues2d=None, anal2d=None, infl3d=None, want_rtps=False,
                  want_nobsl=False, logp=None, copy_back_gues=True, allow_eigen_fail=False):
        """gues3d (nij1, nlev, nens, nv3d) Fortran order (numpy) or the same memory as a torch
        CUDA tensor of shape (nv3d, nens, nlev, nij1).  INOUT: destroyed -> perturbations."""
        dev = _is_torch(gues3d)
        a = capi.DasArgs()
        nlev, nv3d = self.cfg.nlev, self.cfg.nv3d
        if dev:
            import torch
            assert gues3d.is_contiguous()
            if anal3d is None:
                anal3d = torch.empty_like(gues3d)
            rtps = torch.empty((nv3d, nlev, self.nij1), dtype=torch.float64, device=gues3d.device) if want_rtps else None
            nobsl = torch.empty((nlev, self.nij1), dtype=torch.int32, device=gues3d.device) if want_nobsl else None
            a.mem_space = capi.MEM_DEVICE
        else:
            assert gues3d.flags.f_contiguous and gues3d.dtype == np.float64
            if anal3d is None:
                anal3d = np.zeros_like(gues3d, order="F")
            if gues2d is not None and anal2d is None:
                anal2d = np.zeros_like(gues2d, order="F")
            rtps = np.zeros((self.nij1, nlev, nv3d), order="F") if want_rtps else None
            nobsl = np.zeros((self.nij1, nlev), dtype=np.int32, order="F") if want_nobsl else None
            a.mem_space = capi.MEM_HOST
        a.gues3d, a.anal3d = _ptr(gues3d), _ptr(anal3d)
        a.gues2d, a.anal2d = _ptr(gues2d), _ptr(anal2d)
        a.infl3d, a.rtps_infl_out, a.nobsl_out, a.logp = _ptr(infl3d), _ptr(rtps), _ptr(nobsl), _ptr(logp)
        a.reserved = 0 if copy_back_gues else 1
        r = self._ck(self.lib.letkf_b200_das_letkf(self.h, C.byref(a)),
                     allow=(capi.EEIGEN,) if allow_eigen_fail else ())
        st = [C.c_int64() for _ in range(4)]
        self.lib.letkf_b200_das_stats(self.h, *[C.byref(s) for s in st])
        ms, nl = C.c_float(), C.c_int()
        self.lib.letkf_b200_das_kernel_ms(self.h, C.byref(ms), C.byref(nl))
        ph = (C.c_int64 * 8)()
        it = C.c_int64()
        self.lib.letkf_b200_das_phase_clocks(self.h, ph, C.byref(it))
        nref = C.c_int64()
        self.lib.letkf_b200_das_refined(self.h, C.byref(nref))
        return dict(status=r, anal3d=anal3d, anal2d=anal2d, rtps=rtps, nobsl=nobsl, npoints=st[0].value,
                    nsolved=st[1].value, nfail=st[2].value, nobsl_sum=st[3].value, kernel_ms=ms.value,
                    launches=nl.value, phase_clocks=list(ph), solver_iterations=it.value, nrefined=nref.value)

    def ensmean_grd(self, v3d, v2d=None):
        """Fill slot MEMBER+1 with the member mean (in place)."""
        k = self.cfg.MEMBER
        nens = k + 2 if self.cfg.DET_RUN else k + 1
        if _is_torch(v3d):
            nij = v3d.shape[-1]
            space = capi.MEM_DEVICE
        else:
            nij = v3d.shape[0]
            space = capi.MEM_HOST
        self._ck(self.lib.letkf_b200_ensmean_grd(self.h, k, nens, nij, _ptr(v3d), _ptr(v2d), space))

    def enssprd_grd(self, v3d):
        """Ensemble spread around slot MEMBER+1 (common_scale.f90:1557-1611); v3d as in ensmean_grd."""
        k = self.cfg.MEMBER
        nens = k + 2 if self.cfg.DET_RUN else k + 1
        if _is_torch(v3d):
            import torch
            nij = v3d.shape[-1]
            out = torch.empty((self.cfg.nv3d, self.cfg.nlev, nij), dtype=torch.float64, device=v3d.device)
            space = capi.MEM_DEVICE
        else:
            nij = v3d.shape[0]
            out = np.zeros((nij, self.cfg.nlev, self.cfg.nv3d), order="F")
            space = capi.MEM_HOST
        self._ck(self.lib.letkf_b200_enssprd_grd(self.h, k, nens, nij, _ptr(v3d), None, _ptr(out), None, space))
        return out

    def additive_inflation(self, addi3d, anal3d, infl_add, gues3d=None, q_ratio=False, ref_only=False, ishuf=None,
                           addi2d=None, anal2d=None, want_weight=False):
        """Additive inflation block of das_letkf (letkf_tools.f90:804-929), in place on anal3d (anal2d): numpy F-order
        host arrays shaped like gues3d, or torch CUDA tensors with the same memory.  ishuf: permutation of 1..MEMBER."""
        dev = _is_torch(addi3d)
        sh = None if ishuf is None else np.ascontiguousarray(ishuf, dtype=np.int32)
        w = None
        if want_weight:
            if dev:
                import torch
                w = torch.zeros(self.nij1, dtype=torch.float64, device=addi3d.device)
            else:
                w = np.zeros(self.nij1)
        self._ck(self.lib.letkf_b200_additive_inflation(self.h, float(infl_add), int(bool(q_ratio)), int(bool(ref_only)), _ptr(sh),
                                                        _ptr(addi3d), _ptr(addi2d), _ptr(gues3d), _ptr(anal3d), _ptr(anal2d), _ptr(w),
                                                        capi.MEM_DEVICE if dev else capi.MEM_HOST))
        return w

    def nobs_out(self, nvar, pmean, logp=None):
        """NOBS_OUT fields of das_letkf (letkf_tools.f90:440-447, 767-778) for model variable nvar: pmean (nij1, nlev) F-order
        numpy (host) or a torch CUDA tensor with the same memory (nlev, nij1) -> (nij1, nlev, 11) / torch (11, nlev, nij1)."""
        if _is_torch(pmean):
            import torch
            out = torch.empty((11,) + tuple(pmean.shape), dtype=torch.float64, device=pmean.device)
            space = capi.MEM_DEVICE
        else:
            pmean = np.asfortranarray(pmean, dtype=np.float64)
            out = np.zeros(pmean.shape + (11,), order="F")
            space = capi.MEM_HOST
        self._ck(self.lib.letkf_b200_nobs_out(self.h, int(nvar), _ptr(pmean), _ptr(logp), _ptr(out), space))
        return out

    def thermo_defaults(self):
        t = capi.Thermo()
        self.lib.letkf_b200_thermo_defaults(C.byref(t))
        return t

    def state_trans(self, v3dg, thermo=None, inverse=False):
        """state_trans / state_trans_inv (scale/common/common_scale.f90:1181-1280), in place on one member-major
        grid v3dg(nlev,nlon,nlat,nv3d): numpy F-order (host) or a torch CUDA tensor with the same memory."""
        t = thermo if thermo is not None else self.thermo_defaults()
        space = capi.MEM_DEVICE if _is_torch(v3dg) else capi.MEM_HOST
        if space == capi.MEM_HOST:
            assert v3dg.flags.f_contiguous and v3dg.dtype == np.float64
        self._ck(self.lib.letkf_b200_state_trans(self.h, C.byref(t), int(bool(inverse)), _ptr(v3dg), space))
        return v3dg

    # ---- transposes (device pointers; the all-to-all between them is the caller's NCCL call) --
    def nij1_of(self, nprocs_e, myrank_e):
        a, b = C.c_int32(), C.c_int32()
        self._ck(self.lib.letkf_b200_nij1(self.h, nprocs_e, myrank_e, C.byref(a), C.byref(b)))
        return a.value, b.value

    def grd_to_buf(self, nprocs_e, v3dg, v2dg, bufs, thermo=None):
        """thermo given: state_trans is applied on the fly while packing (restart -> LETKF variables)."""
        t = C.byref(thermo) if thermo is not None else None
        self._ck(self.lib.letkf_b200_grd_to_buf_trans(self.h, nprocs_e, t, _ptr(v3dg), _ptr(v2dg), _ptr(bufs)))

    def buf_to_grd(self, nprocs_e, bufr, v3dg, v2dg, thermo=None):
        """thermo given: state_trans_inv is applied on the fly while unpacking."""
        t = C.byref(thermo) if thermo is not None else None
        self._ck(self.lib.letkf_b200_buf_to_grd_trans(self.h, nprocs_e, t, _ptr(bufr), _ptr(v3dg), _ptr(v2dg)))

    def buf_to_ens(self, nprocs_e, myrank_e, nens, mstart, mend, bufr, v3d, v2d):
        self._ck(self.lib.letkf_b200_buf_to_ens(self.h, nprocs_e, myrank_e, nens, mstart, mend, _ptr(bufr),
                                                _ptr(v3d), _ptr(v2d)))

    def ens_to_buf(self, nprocs_e, myrank_e, nens, mstart, mend, v3d, v2d, bufs):
        self._ck(self.lib.letkf_b200_ens_to_buf(self.h, nprocs_e, myrank_e, nens, mstart, mend, _ptr(v3d),
                                                _ptr(v2d), _ptr(bufs)))

    # ---- one-pass transposes over peer memory ------------------------------------------------------
    def peer_export(self, tensor):
        """(handle bytes, offset) of a CUDA tensor's memory, to be sent to the other ranks"""
        d = capi.Ipc()
        self._ck(self.lib.letkf_b200_peer_export(self.h, _ptr(tensor), C.byref(d)))
        return bytes(d.handle), int(d.offset)

    def peer_open(self, desc):
        """device address (int), valid on this rank's GPU, of a peer's exported tensor"""
        d = capi.Ipc()
        C.memmove(d.handle, desc[0], 64)
        d.offset = desc[1]
        out = C.c_void_p()
        self._ck(self.lib.letkf_b200_peer_open(self.h, C.byref(d), C.byref(out)))
        return out.value

    def scatter_grd_p2p(self, nprocs_e, myrank_e, nens, mslot, v3dg, v2dg, peer_v3d, peer_v2d=None, thermo=None):
        """peer_v3d: device addresses (ints) of v3d on ranks 0..np-1"""
        a3 = (C.c_void_p * nprocs_e)(*peer_v3d)
        a2 = (C.c_void_p * nprocs_e)(*peer_v2d) if peer_v2d else None
        t = C.byref(thermo) if thermo is not None else None
        self._ck(self.lib.letkf_b200_scatter_grd_p2p(self.h, nprocs_e, myrank_e, nens, mslot, t, _ptr(v3dg), _ptr(v2dg), a3, a2))

    def gather_grd_p2p(self, nprocs_e, myrank_e, nens, mstart, mend, v3d, v2d, peer_v3dg, peer_v2dg=None, thermo=None):
        """peer_v3dg: device addresses of the member-major grids of members mstart..mend (on ranks 0..mend-mstart)"""
        n = mend - mstart + 1
        a3 = (C.c_void_p * n)(*peer_v3dg)
        a2 = (C.c_void_p * n)(*peer_v2dg) if peer_v2dg else None
        t = C.byref(thermo) if thermo is not None else None
        self._ck(self.lib.letkf_b200_gather_grd_p2p(self.h, nprocs_e, myrank_e, nens, mstart, mend, t, _ptr(v3d), _ptr(v2d), a3, a2))

    # ---- radar observation operator ------------------------------------------------------------------
    def radar_config_defaults(self):
        r = capi.RadarConfig()
        self.lib.letkf_b200_radar_config_defaults(C.byref(r))
        return r

    def obsope_radar(self, rcfg, elm, ril, rjl, lon, lat, lev, grids, rotc=None):
        """H(x_m) of radar observations for all members (obsope_tools.f90:476-494 twin).  grids: list of member history
        grids v3dg(nlevh,nlonh,nlath,nv3dd) -- numpy F-order (host) or torch CUDA tensors.  Returns (yobs, qc), shape
        (nobs, nmem), as numpy arrays (host) or torch tensors (device)."""
        nobs, nmem = len(elm), len(grids)
        dev = _is_torch(grids[0])
        ptrs = (C.c_void_p * nmem)(*[_ptr(g) for g in grids])
        if dev:
            import torch
            d = grids[0].device
            t = lambda a, dt: a if _is_torch(a) else torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=d)
            elm, ril, rjl, lon, lat, lev = (t(elm, np.int32), t(ril, np.float64), t(rjl, np.float64), t(lon, np.float64),
                                            t(lat, np.float64), t(lev, np.float64))
            rotc = None if rotc is None else t(rotc, np.float64)
            y = torch.empty((nobs, nmem), dtype=torch.float64, device=d)
            q = torch.empty((nobs, nmem), dtype=torch.int32, device=d)
            space = capi.MEM_DEVICE
        else:
            f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
            elm = np.ascontiguousarray(elm, dtype=np.int32)
            ril, rjl, lon, lat, lev = f(ril), f(rjl), f(lon), f(lat), f(lev)
            rotc = None if rotc is None else f(rotc)
            y = np.zeros((nobs, nmem))
            q = np.zeros((nobs, nmem), dtype=np.int32)
            space = capi.MEM_HOST
        self._ck(self.lib.letkf_b200_obsope_radar(self.h, C.byref(rcfg), nobs, _ptr(elm), _ptr(ril), _ptr(rjl), _ptr(lon),
                                                  _ptr(lat), _ptr(lev), _ptr(rotc), nmem, ptrs, nmem, _ptr(y), _ptr(q), space))
        return y, q

    # ---- conventional observation operator, monit_obs ---------------------------------------------------
    def conv_config_defaults(self):
        r = capi.ConvConfig()
        self.lib.letkf_b200_conv_config_defaults(C.byref(r))
        return r

    def obsope_conv(self, ccfg, elm, ril, rjl, lev, grids3, grids2, rotc=None):
        """H(x_m) of conventional (prepbufr) observations for all members (obsope_tools.f90:466-473 twin: phys2ijk +
        Trans_XtoY).  grids3 / grids2: per member v3dgh(nlevh,nlonh,nlath,nv3dd) / v2dgh(nlonh,nlath,nv2dd), numpy F-order
        (host) or torch CUDA tensors.  Returns (yobs, qc), shape (nobs, nmem)."""
        nobs, nmem = len(elm), len(grids3)
        dev = _is_torch(grids3[0])
        p3 = (C.c_void_p * nmem)(*[_ptr(g) for g in grids3])
        p2 = (C.c_void_p * nmem)(*[_ptr(g) for g in grids2])
        if dev:
            import torch
            d = grids3[0].device
            t = lambda a, dt: a if _is_torch(a) else torch.as_tensor(np.ascontiguousarray(a, dtype=dt), device=d)
            elm, ril, rjl, lev = t(elm, np.int32), t(ril, np.float64), t(rjl, np.float64), t(lev, np.float64)
            rotc = None if rotc is None else t(rotc, np.float64)
            y = torch.empty((nobs, nmem), dtype=torch.float64, device=d)
            q = torch.empty((nobs, nmem), dtype=torch.int32, device=d)
            space = capi.MEM_DEVICE
        else:
            f = lambda x: np.ascontiguousarray(x, dtype=np.float64)
            elm = np.ascontiguousarray(elm, dtype=np.int32)
            ril, rjl, lev = f(ril), f(rjl), f(lev)
            rotc = None if rotc is None else f(rotc)
            y = np.zeros((nobs, nmem))
            q = np.zeros((nobs, nmem), dtype=np.int32)
            space = capi.MEM_HOST
        self._ck(self.lib.letkf_b200_obsope_conv(self.h, C.byref(ccfg), nobs, _ptr(elm), _ptr(ril), _ptr(rjl), _ptr(lev), _ptr(rotc),
                                                 nmem, p3, p2, nmem, _ptr(y), _ptr(q), space))
        return y, q

    def monit_obs(self, sets, v3dgh, v2dgh, t_range=0.0):
        """monit_obs (common_obs_scale.f90:1370-1844): departure statistics of ONE state (history variables v3dgh, v2dgh: numpy
        F-order host arrays) against the observations of every input set.  sets: list of dicts with `cfg` (ConvConfig for the
        prepbufr format, RadarConfig for the radar format), elm, ril, rjl, lev, dat and, optionally, lon, lat (radar), dif, rotc.
        Returns dict(nobs[16], bias[16], rmse[16], ohx, oqc, elm): the per-observation arrays concatenated in set order."""
        f = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        elms, ohxs, oqcs = [], [], []
        for st in sets:
            cfg = st["cfg"]
            conv = isinstance(cfg, capi.ConvConfig)
            elm = np.ascontiguousarray(st["elm"], dtype=np.int32)
            n = len(elm)
            ohx, oqc = np.zeros(n), np.zeros(n, dtype=np.int32)
            self._ck(self.lib.letkf_b200_monit_obs_set(
                self.h, C.byref(cfg) if conv else None, None if conv else C.byref(cfg), n, _ptr(elm), _ptr(f(st["ril"])),
                _ptr(f(st["rjl"])), _ptr(f(st.get("lon"))), _ptr(f(st.get("lat"))), _ptr(f(st["lev"])), _ptr(f(st["dat"])),
                _ptr(f(st.get("dif"))), _ptr(f(st.get("rotc"))), float(t_range), _ptr(v3dgh), _ptr(v2dgh), _ptr(ohx), _ptr(oqc),
                capi.MEM_HOST))
            elms.append(elm); ohxs.append(ohx); oqcs.append(oqc)
        elm, ohx, oqc = np.concatenate(elms), np.concatenate(ohxs), np.concatenate(oqcs)
        nobs, bias, rmse = self.monit_dep(elm, ohx, oqc)
        return dict(nobs=nobs, bias=bias, rmse=rmse, ohx=ohx, oqc=oqc, elm=elm)

    # ---- device-resident observation chain --------------------------------------------------------------
    def obs_departure_qc_device(self, elm, dat, err, qc, ensval, qcfg=None):
        """departure + QC (letkf_obs.f90:355-560) in place on torch CUDA tensors: ensval (nobs, nensobs) H(x_m) ->
        perturbations, qc updated; returns val (nobs)"""
        import torch
        if qcfg is None:
            qcfg = capi.QcConfig()
            self.lib.letkf_b200_qc_config_defaults(C.byref(qcfg))
        val = torch.zeros(elm.shape[0], dtype=torch.float64, device=ensval.device)
        self._ck(self.lib.letkf_b200_obs_departure_qc(self.h, C.byref(qcfg), elm.shape[0], ensval.shape[1], _ptr(elm), _ptr(dat),
                                                      _ptr(err), _ptr(qc), _ptr(ensval), _ptr(val), capi.MEM_DEVICE))
        return val

    def set_letkf_obs_device(self, obs, qc=None):
        """set_letkf_obs on torch CUDA tensors of ALL observations + their QC flags (int32, 0 = accepted): filter, combined
        types, vertical coordinate and bucket sort on the device.  Returns the number of accepted observations."""
        o = capi.Obs()
        o.nobs = obs["elm"].shape[0]
        o.nensobs = obs["ensval"].shape[1]
        for kf in ("elm", "typ", "ri", "rj", "lev", "dat", "err", "val", "ensval"):
            assert obs[kf].is_contiguous()
            setattr(o, kf, obs[kf].data_ptr())
        nk = C.c_int32()
        self._obs_keepalive = obs
        self._ck(self.lib.letkf_b200_set_obs_device(self.h, C.byref(o), _ptr(qc), C.byref(nk)))
        return nk.value

    def kept_index(self):
        n, _ = self.obs_info()
        a = np.zeros(n, dtype=np.int32)
        self._ck(self.lib.letkf_b200_get_kept_index(self.h, _ptr(a)))
        return a
