"""Namelist defaults and fallback rules of the reference, as a `capi.Config` factory.

Mirrors scale/common/common_nml.f90: PARAM_ENSEMBLE :40-46, PARAM_LETKF :109-142,
PARAM_LETKF_OBS :160-218 (+ "negative inherits element 1" rules :741-775),
PARAM_LETKF_VAR_LOCAL :221-229, PARAM_LETKF_RADAR :264, and the two default-REAL
constants of scale/letkf/letkf_obs.f90:27-28.
"""
import numpy as np

from . import capi

# (double)3.651483717f and (double)13.33333333f: gfortran widens the default-REAL literals
# (letkf_obs.f90:27-28), so these -- not 2*sqrt(10/3) and 40/3 -- are the reference values.
DIST_ZERO_FAC = float(np.float32(3.651483717))
DIST_ZERO_FAC_SQUARE = float(np.float32(13.33333333))

_OBS_MIN_SPACING = [300.0e3, 100.0e3, 100.0e3, 150.0e3, 300.0e3, 150.0e3, 150.0e3, 100.0e3,
                    150.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3,
                    300.0e3, 150.0e3, 150.0e3, 150.0e3, 150.0e3, 1.0e3, 15.0e3, 1000.0e3]


def default_config(**kw):
    """Reference defaults (unresolved: negative entries still mean 'inherit')."""
    c = capi.Config()
    c.MEMBER, c.DET_RUN = 3, 0
    c.nlon = c.nlat = c.nlev = 0
    c.nv3d, c.nv2d = 11, 0
    c.IHALO = c.JHALO = 2
    c.DX = c.DY = 1.0
    c.iv3d_p, c.iv3d_q, c.iv3d_qg = 5, 6, 11
    c.INFL_MUL, c.INFL_MUL_MIN = 1.0, -1.0
    c.INFL_MUL_ADAPTIVE = 0
    c.RELAX_TO_INFLATED_PRIOR = 0
    c.RELAX_ALPHA = c.RELAX_ALPHA_SPREAD = 0.0
    c.Q_UPDATE_TOP, c.Q_SPRD_MAX, c.BOUNDARY_BUFFER_WIDTH = 0.0, -1.0, 0.0
    for t in range(capi.NOBTYPE):
        c.HORI_LOCAL[t] = -1.0
        c.VERT_LOCAL[t] = -1.0
        c.MAX_NOBS_PER_GRID[t] = -1
        c.OBS_MIN_SPACING[t] = _OBS_MIN_SPACING[t]
        c.OBS_SORT_GRID_SPACING[t] = -1.0
    c.HORI_LOCAL[0] = 500.0e3
    c.VERT_LOCAL[0] = 0.4
    c.VERT_LOCAL[21] = 1000.0
    c.MAX_NOBS_PER_GRID[0] = 0
    c.OBS_SORT_GRID_SPACING[0] = 0.0
    c.HORI_LOCAL_RADAR_OBSNOREF = c.HORI_LOCAL_RADAR_VR = c.VERT_LOCAL_RADAR_VR = -1.0
    c.VERT_LOCAL_RAIN_BASE = 85000.0
    c.MAX_NOBS_PER_GRID_CRITERION = 1
    for iv in range(capi.NID_VARLOCAL):
        for n in range(capi.MAX_NV):
            c.VAR_LOCAL[iv][n] = 1.0
    c.RADAR_ZMAX = 99.0e3
    c.dist_zero_fac = DIST_ZERO_FAC
    c.dist_zero_fac_square = DIST_ZERO_FAC_SQUARE
    for key, val in kw.items():
        set_field(c, key, val)
    return c


def set_field(c, key, val):
    """Set a scalar field, or an array field from a dict {index0: value} / sequence."""
    cur = getattr(c, key)
    if hasattr(cur, "__len__"):
        if isinstance(val, dict):
            for i, v in val.items():
                cur[i] = v
        else:
            for i, v in enumerate(val):
                cur[i] = v
    else:
        setattr(c, key, val)


def resolve_config(c):
    """read_nml_letkf_obs fallback rules (common_nml.f90:741-775), in place."""
    for t in range(1, capi.NOBTYPE):
        if c.HORI_LOCAL[t] < 0.0:
            c.HORI_LOCAL[t] = c.HORI_LOCAL[0]
        if c.VERT_LOCAL[t] < 0.0:
            c.VERT_LOCAL[t] = c.VERT_LOCAL[0]
        if c.MAX_NOBS_PER_GRID[t] < 0:
            c.MAX_NOBS_PER_GRID[t] = c.MAX_NOBS_PER_GRID[0]
        if c.OBS_MIN_SPACING[t] <= 0.0:
            c.OBS_MIN_SPACING[t] = c.OBS_MIN_SPACING[0]
        if c.OBS_SORT_GRID_SPACING[t] < 0.0:
            c.OBS_SORT_GRID_SPACING[t] = c.OBS_SORT_GRID_SPACING[0]
    if c.HORI_LOCAL_RADAR_OBSNOREF < 0.0:
        c.HORI_LOCAL_RADAR_OBSNOREF = c.HORI_LOCAL[21]
    if c.HORI_LOCAL_RADAR_VR < 0.0:
        c.HORI_LOCAL_RADAR_VR = c.HORI_LOCAL[21]
    if c.VERT_LOCAL_RADAR_VR < 0.0:
        c.VERT_LOCAL_RADAR_VR = c.VERT_LOCAL[21]
    if not 1 <= c.MAX_NOBS_PER_GRID_CRITERION <= 3:
        raise ValueError("Unsupported MAX_NOBS_PER_GRID_CRITERION")
    return c


def config_bytes(c):
    import ctypes
    return bytes(ctypes.string_at(ctypes.addressof(c), ctypes.sizeof(c)))
