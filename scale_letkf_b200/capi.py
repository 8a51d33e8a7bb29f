"""ctypes view of include/letkf_b200.h (structs, constants, prototypes, loader).

The structures mirror `letkf_b200_config`, `letkf_b200_obs`, `letkf_b200_ctype_info` and
`letkf_b200_das_args` field for field; the names are the reference's namelist / module
names (scale/common/common_nml.f90:19-330, scale/letkf/letkf_obs.f90:33-65).

There is no CPU fallback: `load_library()` raises if the CUDA shared library has not been
built (run `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os

NOBTYPE = 24
NID_OBS = 16
NID_VARLOCAL = 9
MAX_NV = 16
MAX_MEMBER = 4096

OK, EINVAL, ECUDA, ESTATE, EEIGEN, ENOMEM = 0, -1, -2, -3, -4, -5
MEM_HOST, MEM_DEVICE = 0, 1

# raw observation element ids (scale/common/common_obs_scale.f90:45-77)
ID_U, ID_V, ID_T, ID_TV, ID_Q, ID_RH = 2819, 2820, 3073, 3074, 3330, 3331
ID_PS, ID_RAIN = 14593, 19999
ID_RADAR_REF, ID_RADAR_REF_ZERO, ID_RADAR_VR, ID_RADAR_PRH = 4001, 4004, 4002, 4003
# report types (1-based index into obtypelist, common_obs_scale.f90:84-89)
TYP_ADPUPA, TYP_ADPSFC, TYP_PHARAD = 1, 8, 22


class Config(C.Structure):
    _fields_ = [
        ("MEMBER", C.c_int32), ("DET_RUN", C.c_int32),
        ("nlon", C.c_int32), ("nlat", C.c_int32), ("nlev", C.c_int32),
        ("nv3d", C.c_int32), ("nv2d", C.c_int32),
        ("IHALO", C.c_int32), ("JHALO", C.c_int32),
        ("DX", C.c_double), ("DY", C.c_double),
        ("iv3d_p", C.c_int32), ("iv3d_q", C.c_int32), ("iv3d_qg", C.c_int32),
        ("INFL_MUL", C.c_double), ("INFL_MUL_MIN", C.c_double),
        ("INFL_MUL_ADAPTIVE", C.c_int32), ("RELAX_TO_INFLATED_PRIOR", C.c_int32),
        ("RELAX_ALPHA", C.c_double), ("RELAX_ALPHA_SPREAD", C.c_double),
        ("Q_UPDATE_TOP", C.c_double), ("Q_SPRD_MAX", C.c_double),
        ("BOUNDARY_BUFFER_WIDTH", C.c_double),
        ("HORI_LOCAL", C.c_double * NOBTYPE), ("VERT_LOCAL", C.c_double * NOBTYPE),
        ("HORI_LOCAL_RADAR_OBSNOREF", C.c_double), ("HORI_LOCAL_RADAR_VR", C.c_double),
        ("VERT_LOCAL_RADAR_VR", C.c_double), ("VERT_LOCAL_RAIN_BASE", C.c_double),
        ("MAX_NOBS_PER_GRID", C.c_int32 * NOBTYPE), ("MAX_NOBS_PER_GRID_CRITERION", C.c_int32),
        ("OBS_MIN_SPACING", C.c_double * NOBTYPE), ("OBS_SORT_GRID_SPACING", C.c_double * NOBTYPE),
        ("VAR_LOCAL", (C.c_double * MAX_NV) * NID_VARLOCAL),
        ("RADAR_ZMAX", C.c_double),
        ("dist_zero_fac", C.c_double), ("dist_zero_fac_square", C.c_double),
        ("reserved", C.c_int32 * 8),
    ]


class CtypeInfo(C.Structure):
    _fields_ = [
        ("elm", C.c_int32), ("elm_u", C.c_int32), ("typ", C.c_int32),
        ("ngrd_i", C.c_int32), ("ngrd_j", C.c_int32),
        ("ngrdsch_i", C.c_int32), ("ngrdsch_j", C.c_int32),
        ("ngrdext_i", C.c_int32), ("ngrdext_j", C.c_int32),
        ("tot_ext", C.c_int32), ("ac_begin", C.c_int32), ("n_merge", C.c_int32),
        ("hori_loc", C.c_double), ("vert_loc", C.c_double),
        ("grdspc_i", C.c_double), ("grdspc_j", C.c_double),
    ]


class Obs(C.Structure):
    _fields_ = [
        ("nobs", C.c_int32), ("nensobs", C.c_int32),
        ("elm", C.c_void_p), ("typ", C.c_void_p),
        ("ri", C.c_void_p), ("rj", C.c_void_p), ("lev", C.c_void_p),
        ("dat", C.c_void_p), ("err", C.c_void_p), ("val", C.c_void_p),
        ("ensval", C.c_void_p),
    ]


class QcConfig(C.Structure):
    """letkf_b200_qc_config: PARAM_LETKF gross-error factors and PARAM_LETKF_RADAR switches
    (scale/common/common_nml.f90:129-137, 248-259)."""
    _fields_ = [
        ("GROSS_ERROR", C.c_double), ("GROSS_ERROR_RAIN", C.c_double), ("GROSS_ERROR_RADAR_REF", C.c_double),
        ("GROSS_ERROR_RADAR_VR", C.c_double), ("GROSS_ERROR_RADAR_PRH", C.c_double),
        ("GROSS_ERROR_TCX", C.c_double), ("GROSS_ERROR_TCY", C.c_double), ("GROSS_ERROR_TCP", C.c_double),
        ("RADAR_REF_THRES_DBZ", C.c_double),
        ("USE_RADAR_REF", C.c_int32), ("USE_RADAR_VR", C.c_int32),
        ("MIN_RADAR_REF_MEMBER", C.c_int32), ("MIN_RADAR_REF_MEMBER_OBSREF", C.c_int32),
    ]


class Thermo(C.Structure):
    """letkf_b200_thermo: constants of state_trans / state_trans_inv (scale/common/common_scale.f90:1181-1280)."""
    _fields_ = [
        ("Rdry", C.c_double), ("Rvap", C.c_double), ("CVdry", C.c_double), ("PRE00", C.c_double),
        ("TRACER_CV", C.c_double * MAX_NV),
        ("POSITIVE_DEFINITE_Q", C.c_int32), ("POSITIVE_DEFINITE_QHYD", C.c_int32),
    ]


class RadarConfig(C.Structure):
    """letkf_b200_radar_config (radar observation operator, common_nml.f90:257-272 + grid sizes + radar position)"""
    _fields_ = [
        ("METHOD_REF_CALC", C.c_int32), ("USE_TERMINAL_VELOCITY", C.c_int32),
        ("nlevh", C.c_int32), ("nlonh", C.c_int32), ("nlath", C.c_int32), ("nlev", C.c_int32), ("KHALO", C.c_int32),
        ("nv3dd", C.c_int32),
        ("MIN_RADAR_REF_DBZ", C.c_double), ("LOW_REF_SHIFT", C.c_double), ("RADAR_ZMAX", C.c_double),
        ("radar_lon", C.c_double), ("radar_lat", C.c_double), ("radar_z", C.c_double),
    ]


class ConvConfig(C.Structure):
    """letkf_b200_conv_config (conventional observation operator: grid sizes + PS_ADJUST_THRES, common_nml.f90:148)"""
    _fields_ = [
        ("nlevh", C.c_int32), ("nlonh", C.c_int32), ("nlath", C.c_int32), ("nlev", C.c_int32), ("KHALO", C.c_int32),
        ("nv3dd", C.c_int32), ("nv2dd", C.c_int32), ("stggrd", C.c_int32),
        ("PS_ADJUST_THRES", C.c_double),
    ]


class Ipc(C.Structure):
    """letkf_b200_ipc: CUDA IPC handle + offset of a device pointer (one-pass transposes over peer memory)."""
    _fields_ = [("handle", C.c_ubyte * 64), ("offset", C.c_uint64)]


# QC codes (scale/common/common_obs_scale.f90:139-151)
IQC_GOOD, IQC_GROSS_ERR, IQC_REF_MEM, IQC_OBS_BAD, IQC_OTYPE = 0, 5, 12, 50, 90
UNDEF = -9.99e33   # common/common.f90:38


class DasArgs(C.Structure):
    _fields_ = [
        ("gues3d", C.c_void_p), ("gues2d", C.c_void_p),
        ("anal3d", C.c_void_p), ("anal2d", C.c_void_p),
        ("infl3d", C.c_void_p), ("rtps_infl_out", C.c_void_p),
        ("nobsl_out", C.c_void_p), ("logp", C.c_void_p),
        ("mem_space", C.c_int32), ("reserved", C.c_int32),
    ]


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libletkf_b200.so")
_lib = None

_vp, _i, _ip = C.c_void_p, C.c_int, C.POINTER(C.c_int32)
PROTOTYPES = {
    # name: (restype, argtypes) -- one entry per symbol declared in include/letkf_b200.h
    "letkf_b200_config_defaults": (None, [C.POINTER(Config)]),
    "letkf_b200_config_resolve": (None, [C.POINTER(Config)]),
    "letkf_b200_create": (_i, [C.POINTER(Config), _i, C.POINTER(_vp)]),
    "letkf_b200_destroy": (_i, [_vp]),
    "letkf_b200_last_error": (C.c_char_p, [_vp]),
    "letkf_b200_set_stream": (_i, [_vp, _vp]),
    "letkf_b200_core_batch": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _i, _i, _vp, _vp, _i]),
    "letkf_b200_set_grid": (_i, [_vp, _i, _vp, _vp, _vp, _i]),
    "letkf_b200_set_obs": (_i, [_vp, C.POINTER(Obs)]),
    "letkf_b200_obs_info": (_i, [_vp, _ip, _ip]),
    "letkf_b200_get_ctype": (_i, [_vp, _i, C.POINTER(CtypeInfo)]),
    "letkf_b200_get_ac_ext": (_i, [_vp, _i, _vp]),
    "letkf_b200_get_sorted_index": (_i, [_vp, _vp]),
    "letkf_b200_qc_config_defaults": (None, [C.POINTER(QcConfig)]),
    "letkf_b200_obs_departure_qc": (_i, [_vp, C.POINTER(QcConfig), _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "letkf_b200_abi_size_qc": (_i, []),
    "letkf_b200_monit_dep": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "letkf_b200_obs_local": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i]),
    "letkf_b200_das_letkf": (_i, [_vp, C.POINTER(DasArgs)]),
    "letkf_b200_das_stats": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "letkf_b200_das_refined": (_i, [_vp, C.POINTER(C.c_int64)]),
    "letkf_b200_das_kernel_ms": (_i, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "letkf_b200_das_phase_clocks": (_i, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "letkf_b200_ensmean_grd": (_i, [_vp, _i, _i, _i, _vp, _vp, _i]),
    "letkf_b200_nobs_out": (_i, [_vp, _i, _vp, _vp, _vp, _i]),
    "letkf_b200_additive_inflation": (_i, [_vp, C.c_double, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "letkf_b200_enssprd_grd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i]),
    "letkf_b200_thermo_defaults": (None, [C.POINTER(Thermo)]),
    "letkf_b200_state_trans": (_i, [_vp, C.POINTER(Thermo), _i, _vp, _i]),
    "letkf_b200_grd_to_buf": (_i, [_vp, _i, _vp, _vp, _vp]),
    "letkf_b200_grd_to_buf_trans": (_i, [_vp, _i, C.POINTER(Thermo), _vp, _vp, _vp]),
    "letkf_b200_buf_to_grd_trans": (_i, [_vp, _i, C.POINTER(Thermo), _vp, _vp, _vp]),
    "letkf_b200_buf_to_ens": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "letkf_b200_ens_to_buf": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "letkf_b200_buf_to_grd": (_i, [_vp, _i, _vp, _vp, _vp]),
    "letkf_b200_nij1": (_i, [_vp, _i, _i, _ip, _ip]),
    "letkf_b200_set_obs_device": (_i, [_vp, C.POINTER(Obs), _vp, _ip]),
    "letkf_b200_get_kept_index": (_i, [_vp, _vp]),
    "letkf_b200_radar_config_defaults": (None, [C.POINTER(RadarConfig)]),
    "letkf_b200_conv_config_defaults": (None, [C.POINTER(ConvConfig)]),
    "letkf_b200_obsope_conv": (_i, [_vp, C.POINTER(ConvConfig), _i, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_vp), C.POINTER(_vp), _i,
                                    _vp, _vp, _i]),
    "letkf_b200_monit_obs_set": (_i, [_vp, C.POINTER(ConvConfig), C.POINTER(RadarConfig), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp, C.c_double, _vp, _vp, _vp, _vp, _i]),
    "letkf_b200_obsope_radar": (_i, [_vp, C.POINTER(RadarConfig), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, C.POINTER(_vp), _i,
                                     _vp, _vp, _i]),
    "letkf_b200_peer_export": (_i, [_vp, _vp, C.POINTER(Ipc)]),
    "letkf_b200_peer_open": (_i, [_vp, C.POINTER(Ipc), C.POINTER(_vp)]),
    "letkf_b200_scatter_grd_p2p": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "letkf_b200_gather_grd_p2p": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "letkf_b200_abi_sizes": (None, [C.POINTER(C.c_int32 * 4)]),
    "letkf_b200_build_info": (C.c_char_p, []),
}


def load_library(path=None):
    """dlopen libletkf_b200.so and attach prototypes.  Raises if the library is missing:
    the product has no CPU path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # LETKF_B200_LIB: another build of the same library (A/B measurements of kernel variants)
    p = path or os.environ.get("LETKF_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root.")
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError => header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
