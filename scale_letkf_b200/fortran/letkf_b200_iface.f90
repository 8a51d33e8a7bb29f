!===============================================================================
! letkf_b200_iface.f90 -- ISO_C_BINDING view of include/letkf_b200.h and the
! drop-in replacements of the two reference entry points of the analysis path:
!
!     letkf_core(ne,nobs,nobsl,hdxb,rdiag,rloc,dep,parm_infl,trans,transm,pao,
!                rdiag_wloc,infl_update,depd,transmd)     common/common_letkf.f90:52
!     das_letkf(gues3d,gues2d,anal3d,anal2d)              scale/letkf/letkf_tools.f90:50
!
! The argument lists are the reference's, so a caller switches by
!     USE letkf_tools   ->  USE letkf_b200_iface, ONLY: das_letkf => das_letkf_b200
!     USE common_letkf  ->  USE letkf_b200_iface, ONLY: letkf_core => letkf_core_b200
! and links libletkf_b200.so.  Module state the reference's das_letkf reads through
! USE association (namelist scalars, rig1/rjg1/hgt1, obs(:), obsda_sort) is handed
! over once per cycle by letkf_b200_setup (after set_common_mpi_grid and the QC
! half of set_letkf_obs; the bucket-sort half of set_letkf_obs is done on device).
!
! NOT COMPILED IN THE BUILD CONTAINER (no Fortran front-end there): the struct
! layouts below are the same field-for-field tables as scale_letkf_b200/capi.py,
! whose ctypes mirror is checked against sizeof/offsetof by tests/test_capi_cpu.py.
! Compile with the reference: gfortran -cpp -c letkf_b200_iface.f90 (needs the
! reference's common_nml / common_scale / common_mpi_scale / common_obs_scale .mod).
!===============================================================================
module letkf_b200_iface
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: letkf_b200_config, letkf_b200_obs, letkf_b200_das_args
  public :: letkf_b200_setup, letkf_b200_finalize
  public :: das_letkf_b200, letkf_core_b200
  public :: scatter_grd_b200_pack, gather_grd_b200_unpack
  public :: letkf_b200_ipc, letkf_b200_radar_config, letkf_b200_thermo, letkf_b200_qc_config
  public :: scatter_grd_b200_p2p, gather_grd_b200_p2p, c_peer_export, c_peer_open, c_set_obs_device, c_obsope_radar
  public :: additive_inflation_b200, nobs_out_b200, letkf_b200_conv_config, c_obsope_conv, c_monit_obs_set, c_monit_dep

  integer(c_int), parameter, public :: LETKF_B200_NOBTYPE = 24
  integer(c_int), parameter, public :: LETKF_B200_NID_VARLOCAL = 9
  integer(c_int), parameter, public :: LETKF_B200_MAX_NV = 16
  integer(c_int), parameter, public :: LETKF_B200_MEM_HOST = 0, LETKF_B200_MEM_DEVICE = 1
  integer(c_int), parameter, public :: LETKF_B200_OK = 0, LETKF_B200_EEIGEN = -4

  ! struct letkf_b200_config (include/letkf_b200.h:60-92)
  type, bind(C) :: letkf_b200_config
    integer(c_int32_t) :: MEMBER, DET_RUN
    integer(c_int32_t) :: nlon, nlat, nlev
    integer(c_int32_t) :: nv3d, nv2d
    integer(c_int32_t) :: IHALO, JHALO
    real(c_double)     :: DX, DY
    integer(c_int32_t) :: iv3d_p, iv3d_q, iv3d_qg
    real(c_double)     :: INFL_MUL, INFL_MUL_MIN
    integer(c_int32_t) :: INFL_MUL_ADAPTIVE, RELAX_TO_INFLATED_PRIOR
    real(c_double)     :: RELAX_ALPHA, RELAX_ALPHA_SPREAD
    real(c_double)     :: Q_UPDATE_TOP, Q_SPRD_MAX, BOUNDARY_BUFFER_WIDTH
    real(c_double)     :: HORI_LOCAL(LETKF_B200_NOBTYPE), VERT_LOCAL(LETKF_B200_NOBTYPE)
    real(c_double)     :: HORI_LOCAL_RADAR_OBSNOREF, HORI_LOCAL_RADAR_VR, VERT_LOCAL_RADAR_VR
    real(c_double)     :: VERT_LOCAL_RAIN_BASE
    integer(c_int32_t) :: MAX_NOBS_PER_GRID(LETKF_B200_NOBTYPE)
    integer(c_int32_t) :: MAX_NOBS_PER_GRID_CRITERION
    real(c_double)     :: OBS_MIN_SPACING(LETKF_B200_NOBTYPE), OBS_SORT_GRID_SPACING(LETKF_B200_NOBTYPE)
    ! C: VAR_LOCAL[iv][n]  ==  Fortran VAR_LOCAL(n, iv)
    real(c_double)     :: VAR_LOCAL(LETKF_B200_MAX_NV, LETKF_B200_NID_VARLOCAL)
    real(c_double)     :: RADAR_ZMAX
    real(c_double)     :: dist_zero_fac, dist_zero_fac_square
    integer(c_int32_t) :: reserved(8)
  end type

  ! struct letkf_b200_obs (include/letkf_b200.h:109-114)
  type, bind(C) :: letkf_b200_obs
    integer(c_int32_t) :: nobs, nensobs
    type(c_ptr) :: elm, typ
    type(c_ptr) :: ri, rj, lev, dat, err, val
    type(c_ptr) :: ensval
  end type

  ! struct letkf_b200_das_args (include/letkf_b200.h:165-178)
  type, bind(C) :: letkf_b200_das_args
    type(c_ptr) :: gues3d, gues2d, anal3d, anal2d
    type(c_ptr) :: infl3d, rtps_infl_out, nobsl_out, logp
    integer(c_int32_t) :: mem_space, reserved
  end type

  ! struct letkf_b200_qc_config (include/letkf_b200.h): gross-error factors and radar switches
  type, bind(C), public :: letkf_b200_qc_config
    real(c_double)     :: GROSS_ERROR, GROSS_ERROR_RAIN, GROSS_ERROR_RADAR_REF, GROSS_ERROR_RADAR_VR, GROSS_ERROR_RADAR_PRH
    real(c_double)     :: GROSS_ERROR_TCX, GROSS_ERROR_TCY, GROSS_ERROR_TCP
    real(c_double)     :: RADAR_REF_THRES_DBZ
    integer(c_int32_t) :: USE_RADAR_REF, USE_RADAR_VR
    integer(c_int32_t) :: MIN_RADAR_REF_MEMBER, MIN_RADAR_REF_MEMBER_OBSREF
  end type

  ! struct letkf_b200_thermo (include/letkf_b200.h): constants of state_trans / state_trans_inv
  type, bind(C), public :: letkf_b200_thermo
    real(c_double)     :: Rdry, Rvap, CVdry, PRE00
    real(c_double)     :: TRACER_CV(LETKF_B200_MAX_NV)
    integer(c_int32_t) :: POSITIVE_DEFINITE_Q, POSITIVE_DEFINITE_QHYD
  end type

  ! struct letkf_b200_ipc (include/letkf_b200.h): CUDA IPC handle + offset of a device pointer, plain bytes to send to the peers
  type, bind(C), public :: letkf_b200_ipc
    integer(c_signed_char) :: handle(64)
    integer(c_int64_t)     :: offset
  end type

  ! struct letkf_b200_radar_config (include/letkf_b200.h): PARAM_LETKF_RADAR switches + grid sizes + obs(iof)%meta(1:3)
  type, bind(C), public :: letkf_b200_radar_config
    integer(c_int32_t) :: METHOD_REF_CALC, USE_TERMINAL_VELOCITY
    integer(c_int32_t) :: nlevh, nlonh, nlath, nlev, KHALO, nv3dd
    real(c_double)     :: MIN_RADAR_REF_DBZ, LOW_REF_SHIFT, RADAR_ZMAX
    real(c_double)     :: radar_lon, radar_lat, radar_z
  end type

  ! struct letkf_b200_conv_config (include/letkf_b200.h): grid sizes of the history variables + PS_ADJUST_THRES
  type, bind(C), public :: letkf_b200_conv_config
    integer(c_int32_t) :: nlevh, nlonh, nlath, nlev, KHALO, nv3dd, nv2dd, stggrd
    real(c_double)     :: PS_ADJUST_THRES
  end type

  type(c_ptr), save :: handle = c_null_ptr

  interface
    subroutine c_config_defaults(cfg) bind(C, name='letkf_b200_config_defaults')
      import :: letkf_b200_config
      type(letkf_b200_config), intent(out) :: cfg
    end subroutine
    subroutine c_config_resolve(cfg) bind(C, name='letkf_b200_config_resolve')
      import :: letkf_b200_config
      type(letkf_b200_config), intent(inout) :: cfg
    end subroutine
    integer(c_int) function c_create(cfg, device, h) bind(C, name='letkf_b200_create')
      import :: letkf_b200_config, c_int, c_ptr
      type(letkf_b200_config), intent(in) :: cfg
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: h
    end function
    integer(c_int) function c_destroy(h) bind(C, name='letkf_b200_destroy')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function
    function c_last_error(h) result(msg) bind(C, name='letkf_b200_last_error')
      import :: c_ptr
      type(c_ptr), value :: h
      type(c_ptr) :: msg
    end function
    integer(c_int) function c_core_batch(h, ne, nobs, npts, nobsl, hdxb, rdiag, rloc, dep, parm_infl, &
                                         trans, transm, pao, rdiag_wloc, infl_update, depd, transmd, &
                                         mem_space) bind(C, name='letkf_b200_core_batch')
      import :: c_int, c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: ne, nobs, npts, rdiag_wloc, infl_update, mem_space
      integer(c_int32_t), intent(in) :: nobsl(*)
      real(c_double), intent(in) :: hdxb(*), rdiag(*), rloc(*), dep(*)
      real(c_double), intent(inout) :: parm_infl(*)
      real(c_double), intent(out) :: trans(*)
      type(c_ptr), value :: transm, pao, depd, transmd   ! OPTIONAL absent -> c_null_ptr
    end function
    ! departure + QC half of set_letkf_obs (scale/letkf/letkf_obs.f90:355-560): call with obsda%qc, obsda%ensval,
    ! obsda%val and the obs(set)%{elm,dat,err}(idx) gathered per obsda entry, before the bucket sort
    integer(c_int) function c_obs_departure_qc(h, q, nobs, nensobs, elm, dat, err, qc, ensval, val, mem_space) &
        bind(C, name='letkf_b200_obs_departure_qc')
      import :: c_int, c_ptr, c_int32_t, c_double, letkf_b200_qc_config
      type(c_ptr), value :: h
      type(letkf_b200_qc_config), intent(in) :: q
      integer(c_int), value :: nobs, nensobs, mem_space
      integer(c_int32_t), intent(in) :: elm(*)
      real(c_double), intent(in) :: dat(*), err(*)
      integer(c_int32_t), intent(inout) :: qc(*)
      real(c_double), intent(inout) :: ensval(*), val(*)
    end function
    ! state_trans (inverse = 0) / state_trans_inv (inverse = 1), scale/common/common_scale.f90:1181-1280:
    ! fill letkf_b200_thermo from scale_const (Rdry, Rvap, CVdry, PRE00) and scale_tracer (TRACER_CV)
    integer(c_int) function c_state_trans(h, t, inverse, v3dg, mem_space) bind(C, name='letkf_b200_state_trans')
      import :: c_int, c_ptr, c_double, letkf_b200_thermo
      type(c_ptr), value :: h
      type(letkf_b200_thermo), intent(in) :: t
      integer(c_int), value :: inverse, mem_space
      real(c_double), intent(inout) :: v3dg(*)
    end function
    ! monit_dep (scale/common/common_obs_scale.f90:1851): nobs(nid_obs), bias(nid_obs), rmse(nid_obs)
    integer(c_int) function c_monit_dep(h, nn, elm, dep, qc, nobs, bias, rmse, mem_space) bind(C, name='letkf_b200_monit_dep')
      import :: c_int, c_ptr, c_int32_t, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: nn, mem_space
      integer(c_int32_t), intent(in) :: elm(*), qc(*)     ! elm: NINT(obs%elm)
      real(c_double), intent(in) :: dep(*)
      integer(c_int32_t), intent(out) :: nobs(*)
      real(c_double), intent(out) :: bias(*), rmse(*)
    end function
    integer(c_int) function c_set_grid(h, nij1, rig1, rjg1, hgt1, mem_space) bind(C, name='letkf_b200_set_grid')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: nij1, mem_space
      real(c_double), intent(in) :: rig1(*), rjg1(*), hgt1(*)
    end function
    integer(c_int) function c_set_obs(h, obs) bind(C, name='letkf_b200_set_obs')
      import :: c_int, c_ptr, letkf_b200_obs
      type(c_ptr), value :: h
      type(letkf_b200_obs), intent(in) :: obs
    end function
    integer(c_int) function c_das_letkf(h, args) bind(C, name='letkf_b200_das_letkf')
      import :: c_int, c_ptr, letkf_b200_das_args
      type(c_ptr), value :: h
      type(letkf_b200_das_args), intent(in) :: args
    end function
    integer(c_int) function c_ensmean_grd(h, mem, nens, nij, v3d, v2d, mem_space) bind(C, name='letkf_b200_ensmean_grd')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: mem, nens, nij, mem_space
      real(c_double), intent(inout) :: v3d(*), v2d(*)
    end function
    integer(c_int) function c_grd_to_buf(h, np, v3dg, v2dg, bufs) bind(C, name='letkf_b200_grd_to_buf')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, v3dg, v2dg, bufs      ! device pointers
      integer(c_int), value :: np
    end function
    ! pack / unpack with state_trans / state_trans_inv fused in (t = c_null_ptr: plain pack / unpack)
    integer(c_int) function c_grd_to_buf_trans(h, np, t, v3dg, v2dg, bufs) bind(C, name='letkf_b200_grd_to_buf_trans')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, t, v3dg, v2dg, bufs
      integer(c_int), value :: np
    end function
    integer(c_int) function c_buf_to_grd_trans(h, np, t, bufr, v3dg, v2dg) bind(C, name='letkf_b200_buf_to_grd_trans')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, t, bufr, v3dg, v2dg
      integer(c_int), value :: np
    end function
    integer(c_int) function c_buf_to_ens(h, np, myrank_e, nens, mstart, mend, bufr, v3d, v2d) &
        bind(C, name='letkf_b200_buf_to_ens')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, bufr, v3d, v2d
      integer(c_int), value :: np, myrank_e, nens, mstart, mend
    end function
    integer(c_int) function c_ens_to_buf(h, np, myrank_e, nens, mstart, mend, v3d, v2d, bufs) &
        bind(C, name='letkf_b200_ens_to_buf')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, v3d, v2d, bufs
      integer(c_int), value :: np, myrank_e, nens, mstart, mend
    end function
    integer(c_int) function c_buf_to_grd(h, np, bufr, v3dg, v2dg) bind(C, name='letkf_b200_buf_to_grd')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, bufr, v3dg, v2dg
      integer(c_int), value :: np
    end function
    ! ---- one-pass transposes with the exchange inside (scatter/gather_grd_mpi_alltoall, common_mpi_scale.f90:1279-1396) ----
    integer(c_int) function c_peer_export(h, devptr, ipc) bind(C, name='letkf_b200_peer_export')
      import :: c_int, c_ptr, letkf_b200_ipc
      type(c_ptr), value :: h, devptr
      type(letkf_b200_ipc), intent(out) :: ipc
    end function
    integer(c_int) function c_peer_open(h, ipc, mapped) bind(C, name='letkf_b200_peer_open')
      import :: c_int, c_ptr, letkf_b200_ipc
      type(c_ptr), value :: h
      type(letkf_b200_ipc), intent(in) :: ipc
      type(c_ptr), intent(out) :: mapped
    end function
    integer(c_int) function c_scatter_grd_p2p(h, np, myrank_e, nens, mslot, t, v3dg, v2dg, peer_v3d, peer_v2d) &
        bind(C, name='letkf_b200_scatter_grd_p2p')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, t, v3dg, v2dg          ! t: letkf_b200_thermo or c_null_ptr (no state_trans); device pointers
      integer(c_int), value :: np, myrank_e, nens, mslot
      type(c_ptr), intent(in) :: peer_v3d(*), peer_v2d(*)   ! device address of v3d / v2d on every rank (own + peer_open)
    end function
    integer(c_int) function c_gather_grd_p2p(h, np, myrank_e, nens, mstart, mend, t, v3d, v2d, peer_v3dg, peer_v2dg) &
        bind(C, name='letkf_b200_gather_grd_p2p')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, t, v3d, v2d
      integer(c_int), value :: np, myrank_e, nens, mstart, mend
      type(c_ptr), intent(in) :: peer_v3dg(*), peer_v2dg(*)  ! member-major grids of members mstart..mend on their owners
    end function
    ! ---- device-resident observation chain (letkf_obs.f90:308-342, 747-805) and the radar operator (obsope_tools.f90:476-494) ----
    integer(c_int) function c_set_obs_device(h, obs, qc, nkept) bind(C, name='letkf_b200_set_obs_device')
      import :: c_int, c_ptr, c_int32_t, letkf_b200_obs
      type(c_ptr), value :: h, qc                      ! qc: device int32 array or c_null_ptr
      type(letkf_b200_obs), intent(in) :: obs          ! every pointer a DEVICE array
      integer(c_int32_t), intent(out) :: nkept
    end function
    integer(c_int) function c_obsope_radar(h, r, nobs, elm, ril, rjl, lon, lat, lev, rotc, nmem, v3dgh, ld_out, yobs, qc, &
                                           mem_space) bind(C, name='letkf_b200_obsope_radar')
      import :: c_int, c_ptr, letkf_b200_radar_config
      type(c_ptr), value :: h, elm, ril, rjl, lon, lat, lev, rotc, yobs, qc
      type(letkf_b200_radar_config), intent(in) :: r
      integer(c_int), value :: nobs, nmem, ld_out, mem_space
      type(c_ptr), intent(in) :: v3dgh(*)              ! v3dgh(nlevh,nlonh,nlath,nv3dd) of every member
    end function
    ! ---- conventional operator (obsope_tools.f90:466-473) and the observation loop of monit_obs (common_obs_scale.f90:1516-1572) ----
    integer(c_int) function c_obsope_conv(h, r, nobs, elm, ril, rjl, lev, rotc, nmem, v3dgh, v2dgh, ld_out, yobs, qc, mem_space) &
        bind(C, name='letkf_b200_obsope_conv')
      import :: c_int, c_ptr, letkf_b200_conv_config
      type(c_ptr), value :: h, elm, ril, rjl, lev, rotc, yobs, qc
      type(letkf_b200_conv_config), intent(in) :: r
      integer(c_int), value :: nobs, nmem, ld_out, mem_space
      type(c_ptr), intent(in) :: v3dgh(*), v2dgh(*)    ! v3dgh(nlevh,nlonh,nlath,nv3dd), v2dgh(nlonh,nlath,nv2dd) of every member
    end function
    ! conv / radar: exactly one is a c_loc(config), the other c_null_ptr; then c_monit_dep over the concatenated (elm, ohx, oqc)
    integer(c_int) function c_monit_obs_set(h, conv, radar, nobs, elm, ril, rjl, lon, lat, lev, dat, dif, rotc, t_range, v3dgh, v2dgh, &
                                            ohx, oqc, mem_space) bind(C, name='letkf_b200_monit_obs_set')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h, conv, radar, elm, ril, rjl, lon, lat, lev, dat, dif, rotc, v3dgh, v2dgh, ohx, oqc
      integer(c_int), value :: nobs, mem_space
      real(c_double), value :: t_range
    end function
    ! ---- post-loop blocks of das_letkf ----
    integer(c_int) function c_additive_inflation(h, infl_add, q_ratio, ref_only, ishuf, addi3d, addi2d, gues3d, anal3d, anal2d, &
                                                 weight_out, mem_space) bind(C, name='letkf_b200_additive_inflation')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h, ishuf, addi3d, addi2d, gues3d, anal3d, anal2d, weight_out
      real(c_double), value :: infl_add
      integer(c_int), value :: q_ratio, ref_only, mem_space
    end function
    integer(c_int) function c_nobs_out(h, nvar, pmean, logp, out, mem_space) bind(C, name='letkf_b200_nobs_out')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, pmean, logp, out
      integer(c_int), value :: nvar, mem_space
    end function
  end interface

contains

  !-----------------------------------------------------------------------------
  ! Hand the module state das_letkf reads to the library.  Call once per cycle,
  ! after set_common_mpi_grid (rig1, rjg1, hgt1: common_mpi_scale.f90:303-308)
  ! and after the QC / departure half of set_letkf_obs (letkf_obs.f90:268-622):
  ! the arguments are the QC-passed observations in arrival order; the ctype
  ! table, sorting mesh, counting sort and extended prefix sums
  ! (letkf_obs.f90:308-342, 660-976) are rebuilt on the device.
  !-----------------------------------------------------------------------------
  subroutine letkf_b200_setup(device, nobs, nensobs, elm, typ, ri, rj, lev, dat, err, val, ensval)
    use common_nml          ! MEMBER, DET_RUN, INFL_MUL, ..., HORI_LOCAL(:), VAR_LOCAL_*(:)
    use common_scale, only: nlong, nlatg, nlev, nv3d, nv2d, iv3d_p, iv3d_q, iv3d_qg
    use common_mpi_scale, only: nij1, rig1, rjg1, hgt1
    use letkf_obs, only: dist_zero_fac, dist_zero_fac_square
    use scale_grid, only: DX, DY
    use scale_grid_index, only: IHALO, JHALO
    integer, intent(in) :: device, nobs, nensobs
    integer(c_int32_t), intent(in), target :: elm(nobs), typ(nobs)
    real(c_double), intent(in), target :: ri(nobs), rj(nobs), lev(nobs), dat(nobs), err(nobs), val(nobs)
    real(c_double), intent(in), target :: ensval(nensobs, nobs)
    type(letkf_b200_config) :: cfg
    type(letkf_b200_obs) :: o
    integer :: n

    call c_config_defaults(cfg)
    cfg%MEMBER = MEMBER
    cfg%DET_RUN = merge(1, 0, DET_RUN)
    ! The library works in GLOBAL grid-index space ("PRC 1x1 view"): rig1/rjg1 (common_mpi_scale.f90:303-308) and
    ! obs%ri/rj (phys2ij) are global indices, so the sorting mesh and relax_beta take the global domain size.  On a
    ! decomposed domain (PRC_NUM_X*PRC_NUM_Y > 1) hand letkf_b200_setup AT LEAST the observations of this rank's
    ! extended subdomain (obsda_ext, letkf_obs.f90:918-1138); the local lists then hold the same observations as the
    ! reference's (bucket scan order, i.e. summation order, may differ: 1e-10 class).  Passing the subdomain
    ! nlon/nlat here would push every observation of ranks with iproc/jproc > 0 into the last bucket column/row.
    cfg%nlon = nlong;  cfg%nlat = nlatg;  cfg%nlev = nlev
    cfg%nv3d = nv3d;  cfg%nv2d = nv2d
    cfg%IHALO = IHALO; cfg%JHALO = JHALO
    cfg%DX = DX; cfg%DY = DY
    cfg%iv3d_p = iv3d_p; cfg%iv3d_q = iv3d_q; cfg%iv3d_qg = iv3d_qg
    cfg%INFL_MUL = INFL_MUL; cfg%INFL_MUL_MIN = INFL_MUL_MIN
    cfg%INFL_MUL_ADAPTIVE = merge(1, 0, INFL_MUL_ADAPTIVE)
    cfg%RELAX_TO_INFLATED_PRIOR = merge(1, 0, RELAX_TO_INFLATED_PRIOR)
    cfg%RELAX_ALPHA = RELAX_ALPHA; cfg%RELAX_ALPHA_SPREAD = RELAX_ALPHA_SPREAD
    cfg%Q_UPDATE_TOP = Q_UPDATE_TOP; cfg%Q_SPRD_MAX = Q_SPRD_MAX
    cfg%BOUNDARY_BUFFER_WIDTH = BOUNDARY_BUFFER_WIDTH
    cfg%HORI_LOCAL = HORI_LOCAL; cfg%VERT_LOCAL = VERT_LOCAL
    cfg%HORI_LOCAL_RADAR_OBSNOREF = HORI_LOCAL_RADAR_OBSNOREF
    cfg%HORI_LOCAL_RADAR_VR = HORI_LOCAL_RADAR_VR
    cfg%VERT_LOCAL_RADAR_VR = VERT_LOCAL_RADAR_VR
    cfg%VERT_LOCAL_RAIN_BASE = VERT_LOCAL_RAIN_BASE
    cfg%MAX_NOBS_PER_GRID = MAX_NOBS_PER_GRID
    cfg%MAX_NOBS_PER_GRID_CRITERION = MAX_NOBS_PER_GRID_CRITERION
    cfg%OBS_MIN_SPACING = OBS_MIN_SPACING
    cfg%OBS_SORT_GRID_SPACING = OBS_SORT_GRID_SPACING
    do n = 1, nv3d + nv2d       ! var_local(n, iv), letkf_tools.f90:130-141
      cfg%VAR_LOCAL(n, 1) = VAR_LOCAL_UV(n)
      cfg%VAR_LOCAL(n, 2) = VAR_LOCAL_T(n)
      cfg%VAR_LOCAL(n, 3) = VAR_LOCAL_Q(n)
      cfg%VAR_LOCAL(n, 4) = VAR_LOCAL_PS(n)
      cfg%VAR_LOCAL(n, 5) = VAR_LOCAL_RAIN(n)
      cfg%VAR_LOCAL(n, 6) = VAR_LOCAL_TC(n)
      cfg%VAR_LOCAL(n, 7) = VAR_LOCAL_RADAR_REF(n)
      cfg%VAR_LOCAL(n, 8) = VAR_LOCAL_RADAR_VR(n)
      cfg%VAR_LOCAL(n, 9) = VAR_LOCAL_H08(n)
    end do
    cfg%RADAR_ZMAX = RADAR_ZMAX
    ! default-REAL literals widened to double by the compiler that built the host program:
    ! carried as data so that the cut-off tests are bit-identical (letkf_obs.f90:27-28)
    cfg%dist_zero_fac = dist_zero_fac
    cfg%dist_zero_fac_square = dist_zero_fac_square
    call c_config_resolve(cfg)

    ! namelist switches whose post-processing (letkf_tools.f90:693-932) is outside this wrapper: refuse loudly rather
    ! than return a different analysis or silently skip an output file
    ! INFL_ADD > 0: call additive_inflation_b200 after das_letkf_b200 with the ensemble read_ens_mpi_addiinfl returns
    if (nv2d > 0 .and. (INFL_MUL <= 0.0d0 .or. INFL_MUL_ADAPTIVE)) &
      call unsupported('2-D variables with INFL_MUL <= 0 or INFL_MUL_ADAPTIVE (no 2-D inflation field in the interface)')

    if (c_associated(handle)) call check(c_destroy(handle), 'destroy')
    call check(c_create(cfg, int(device, c_int), handle), 'create')
    call check(c_set_grid(handle, int(nij1, c_int), rig1, rjg1, hgt1, LETKF_B200_MEM_HOST), 'set_grid')
    o%nobs = nobs; o%nensobs = nensobs
    o%elm = c_loc(elm); o%typ = c_loc(typ)
    o%ri = c_loc(ri); o%rj = c_loc(rj); o%lev = c_loc(lev); o%dat = c_loc(dat)
    o%err = c_loc(err); o%val = c_loc(val); o%ensval = c_loc(ensval)
    call check(c_set_obs(handle, o), 'set_obs')
  end subroutine letkf_b200_setup

  subroutine letkf_b200_finalize()
    if (c_associated(handle)) call check(c_destroy(handle), 'destroy')
    handle = c_null_ptr
  end subroutine

  !-----------------------------------------------------------------------------
  ! Drop-in for das_letkf (letkf_tools.f90:50): same arguments, same INTENTs.
  ! gues3d/gues2d are INTENT(INOUT) "destroyed" in the reference; PROGRAM letkf never reads them again
  ! (letkf.f90:196-236), so the perturbations are NOT copied back from the device (reserved = 1) -- pass
  ! copy_back = .true. if a caller does rely on slots 1..MEMBER holding dX on return.
  ! Optional fields, all (nij1,nlev,nv3d) unless noted, replace the module-level work arrays of the reference:
  !   infl3d   INOUT  work3d: multiplicative inflation.  Required when INFL_MUL <= 0 (the caller reads INFL_MUL_IN_BASENAME
  !                   into it first, letkf_tools.f90:240-262) or INFL_MUL_ADAPTIVE (returns the adapted field for
  !                   INFL_MUL_OUT_BASENAME, :693-720)
  !   rtps3d   OUT    work3da: RTPS relaxation factors for RELAX_SPREAD_OUT (:722-750)
  !   nobs2d   OUT    (nij1,nlev) number of local observations used, for NOBS_OUT (:752-802; per-type counts are not returned)
  ! The gather / write_restart of these fields stays with the caller, exactly as in the reference.
  !-----------------------------------------------------------------------------
  subroutine das_letkf_b200(gues3d, gues2d, anal3d, anal2d, infl3d, rtps3d, nobs2d, copy_back)
    use common_nml, only: INFL_MUL, INFL_MUL_ADAPTIVE, RELAX_SPREAD_OUT, NOBS_OUT
    use common_scale, only: nlev, nv3d, nv2d, iv3d_p
    use common_mpi_scale, only: nij1, nens, mmean
    real(c_double), intent(inout), target :: gues3d(nij1, nlev, nens, nv3d)
    real(c_double), intent(inout), target :: gues2d(nij1, nens, nv2d)
    real(c_double), intent(out), target :: anal3d(nij1, nlev, nens, nv3d)
    real(c_double), intent(out), target :: anal2d(nij1, nens, nv2d)
    real(c_double), intent(inout), optional, target :: infl3d(nij1, nlev, nv3d)
    real(c_double), intent(out), optional, target :: rtps3d(nij1, nlev, nv3d)
    integer(c_int32_t), intent(out), optional, target :: nobs2d(nij1, nlev)
    logical, intent(in), optional :: copy_back
    real(c_double), allocatable, target :: logp(:, :)
    type(letkf_b200_das_args) :: a
    if ((INFL_MUL <= 0.0d0 .or. INFL_MUL_ADAPTIVE) .and. .not. present(infl3d)) &
      call unsupported('INFL_MUL <= 0 / INFL_MUL_ADAPTIVE without the infl3d argument of das_letkf_b200')
    if (RELAX_SPREAD_OUT .and. .not. present(rtps3d)) &
      call unsupported('RELAX_SPREAD_OUT without the rtps3d argument of das_letkf_b200')
    if (NOBS_OUT .and. .not. present(nobs2d)) call unsupported('NOBS_OUT without the nobs2d argument of das_letkf_b200')
    a%gues3d = c_loc(gues3d); a%anal3d = c_loc(anal3d)
    a%gues2d = c_null_ptr;    a%anal2d = c_null_ptr
    if (nv2d > 0) then
      a%gues2d = c_loc(gues2d); a%anal2d = c_loc(anal2d)
    end if
    a%infl3d = c_null_ptr; a%rtps_infl_out = c_null_ptr; a%nobsl_out = c_null_ptr
    if (present(infl3d)) a%infl3d = c_loc(infl3d)
    if (present(rtps3d)) a%rtps_infl_out = c_loc(rtps3d)
    if (present(nobs2d)) a%nobsl_out = c_loc(nobs2d)
    ! ln(mean pressure) with the HOST's log (the same libm as the reference's obs_local_cal, letkf_tools.f90:1852-1866):
    ! local-observation selection is then bit-identical to a CPU run on this machine.  (Left null, the library
    ! computes the same table on the host for host buffers; spelled out here so that the contract is visible.)
    allocate (logp(nij1, nlev))
    logp = log(gues3d(:, :, mmean, iv3d_p))
    a%logp = c_loc(logp)
    a%mem_space = LETKF_B200_MEM_HOST
    a%reserved = 1
    if (present(copy_back)) then
      if (copy_back) a%reserved = 0
    end if
    call check(c_das_letkf(handle, a), 'das_letkf')   ! the reference STOPs on eigensolver failure
    deallocate (logp)
  end subroutine das_letkf_b200

  !-----------------------------------------------------------------------------
  ! Drop-in for letkf_core (common_letkf.f90:52): one point per call (npts = 1).
  ! Callers that can batch points should call letkf_b200_core_batch directly.
  !-----------------------------------------------------------------------------
  subroutine letkf_core_b200(ne, nobs, nobsl, hdxb, rdiag, rloc, dep, parm_infl, trans, transm, pao, &
                             rdiag_wloc, infl_update, depd, transmd)
    integer, intent(in) :: ne, nobs, nobsl
    real(c_double), intent(in) :: hdxb(nobs, ne), rdiag(nobs), rloc(nobs), dep(nobs)
    real(c_double), intent(inout) :: parm_infl
    real(c_double), intent(out) :: trans(ne, ne)
    real(c_double), intent(out), optional, target :: transm(ne), pao(ne, ne)
    logical, intent(in), optional :: rdiag_wloc, infl_update
    real(c_double), intent(in), optional, target :: depd(nobs)
    real(c_double), intent(out), optional, target :: transmd(ne)
    integer(c_int32_t) :: nl(1)
    real(c_double) :: infl(1)
    integer(c_int) :: wl, iu
    type(c_ptr) :: p_transm, p_pao, p_depd, p_transmd
    nl(1) = nobsl; infl(1) = parm_infl
    wl = 0; iu = 0
    if (present(rdiag_wloc)) wl = merge(1, 0, rdiag_wloc)
    if (present(infl_update)) iu = merge(1, 0, infl_update)
    p_transm = c_null_ptr; p_pao = c_null_ptr; p_depd = c_null_ptr; p_transmd = c_null_ptr
    if (present(transm)) p_transm = c_loc(transm)
    if (present(pao)) p_pao = c_loc(pao)
    if (present(depd) .and. present(transmd)) then    ! common_letkf.f90:196
      p_depd = c_loc(depd); p_transmd = c_loc(transmd)
    end if
    call check(c_core_batch(handle, int(ne, c_int), int(nobs, c_int), 1_c_int, nl, hdxb, rdiag, rloc, dep, &
                            infl, trans, p_transm, p_pao, wl, iu, p_depd, p_transmd, LETKF_B200_MEM_HOST), &
               'letkf_core')
    parm_infl = infl(1)
  end subroutine letkf_core_b200

  !-----------------------------------------------------------------------------
  ! Pack / unpack halves of scatter_grd_mpi_alltoall / gather_grd_mpi_alltoall
  ! (common_mpi_scale.f90:1279-1396) on DEVICE buffers; the exchange between them
  ! is the caller's all-to-all (NCCL send/recv group, or CUDA-aware MPI_ALLTOALLV on
  ! the same device pointers with the counts of set_alltoall_counts).
  !-----------------------------------------------------------------------------
  subroutine scatter_grd_b200_pack(np, d_v3dg, d_v2dg, d_bufs)
    integer, intent(in) :: np
    type(c_ptr), intent(in) :: d_v3dg, d_v2dg, d_bufs
    call check(c_grd_to_buf(handle, int(np, c_int), d_v3dg, d_v2dg, d_bufs), 'grd_to_buf')
  end subroutine
  subroutine gather_grd_b200_unpack(np, d_bufr, d_v3dg, d_v2dg)
    integer, intent(in) :: np
    type(c_ptr), intent(in) :: d_bufr, d_v3dg, d_v2dg
    call check(c_buf_to_grd(handle, int(np, c_int), d_bufr, d_v3dg, d_v2dg), 'buf_to_grd')
  end subroutine

  !-----------------------------------------------------------------------------
  ! scatter_grd_mpi_alltoall / gather_grd_mpi_alltoall (common_mpi_scale.f90:1279-1396) in ONE pass with the exchange
  ! inside the kernel: every rank reads its own DEVICE arrays and stores into the receiving ranks' arrays over NVLink.
  ! peer_*(1:np): the device address of the destination array on every e-rank as seen from this GPU -- this rank's own
  ! pointer at index myrank_e + 1, the others mapped ONCE with letkf_b200_peer_export on the owner (send the 72-byte
  ! letkf_b200_ipc with MPI_ALLGATHER over MPI_COMM_e) and letkf_b200_peer_open here.  The MPI_BARRIER(MPI_COMM_e) the
  ! caller places before (nobody still reads the destination) and after the call stands where MPI_ALLTOALL blocked.
  ! thermo present: state_trans / state_trans_inv (common_scale.f90:1181-1280) are applied on the way.
  !-----------------------------------------------------------------------------
  subroutine scatter_grd_b200_p2p(np, myrank_e, nens, mslot, d_v3dg, d_v2dg, peer_v3d, peer_v2d, thermo)
    integer, intent(in) :: np, myrank_e, nens, mslot          ! mslot: the member slot (1-based) this rank's grid fills
    type(c_ptr), intent(in) :: d_v3dg, d_v2dg, peer_v3d(np), peer_v2d(np)
    type(letkf_b200_thermo), intent(in), target, optional :: thermo
    type(c_ptr) :: pt
    pt = c_null_ptr
    if (present(thermo)) pt = c_loc(thermo)
    call check(c_scatter_grd_p2p(handle, int(np, c_int), int(myrank_e, c_int), int(nens, c_int), int(mslot, c_int), pt, &
                                 d_v3dg, d_v2dg, peer_v3d, peer_v2d), 'scatter_grd_p2p')
  end subroutine
  subroutine gather_grd_b200_p2p(np, myrank_e, nens, mstart, mend, d_v3d, d_v2d, peer_v3dg, peer_v2dg, thermo)
    integer, intent(in) :: np, myrank_e, nens, mstart, mend
    type(c_ptr), intent(in) :: d_v3d, d_v2d, peer_v3dg(np), peer_v2dg(np)
    type(letkf_b200_thermo), intent(in), target, optional :: thermo
    type(c_ptr) :: pt
    pt = c_null_ptr
    if (present(thermo)) pt = c_loc(thermo)
    call check(c_gather_grd_p2p(handle, int(np, c_int), int(myrank_e, c_int), int(nens, c_int), int(mstart, c_int), &
                                int(mend, c_int), pt, d_v3d, d_v2d, peer_v3dg, peer_v2dg), 'gather_grd_p2p')
  end subroutine

  !-----------------------------------------------------------------------------
  ! Additive inflation block of das_letkf (letkf_tools.f90:804-929): call after das_letkf_b200 when INFL_ADD > 0 with the
  ! ensemble read_ens_mpi_addiinfl delivered (host arrays laid out like gues3d / gues2d; NOT modified), the background
  ! array (its mean slot is read when INFL_ADD_Q_RATIO) and, with INFL_ADD_SHUFFLE, the permutation the caller drew with
  ! Knuth_Shuffle and broadcast (:861-866).
  !-----------------------------------------------------------------------------
  subroutine additive_inflation_b200(addi3d, addi2d, gues3d, anal3d, anal2d, ishuf)
    use common_nml, only: INFL_ADD, INFL_ADD_Q_RATIO, INFL_ADD_REF_ONLY, MEMBER
    use common_scale, only: nlev, nv3d, nv2d
    use common_mpi_scale, only: nij1, nens
    real(c_double), intent(in), target :: addi3d(nij1,nlev,nens,nv3d), addi2d(nij1,nens,nv2d), gues3d(nij1,nlev,nens,nv3d)
    real(c_double), intent(inout), target :: anal3d(nij1,nlev,nens,nv3d), anal2d(nij1,nens,nv2d)
    integer(c_int32_t), intent(in), target, optional :: ishuf(MEMBER)
    type(c_ptr) :: p_sh, p_a2, p_n2
    p_sh = c_null_ptr; p_a2 = c_null_ptr; p_n2 = c_null_ptr
    if (present(ishuf)) p_sh = c_loc(ishuf)
    if (nv2d > 0) then
      p_a2 = c_loc(addi2d); p_n2 = c_loc(anal2d)
    end if
    call check(c_additive_inflation(handle, INFL_ADD, merge(1_c_int, 0_c_int, INFL_ADD_Q_RATIO), &
                                    merge(1_c_int, 0_c_int, INFL_ADD_REF_ONLY), p_sh, c_loc(addi3d), p_a2, c_loc(gues3d), &
                                    c_loc(anal3d), p_n2, c_null_ptr, LETKF_B200_MEM_HOST), 'additive_inflation')
  end subroutine

  !-----------------------------------------------------------------------------
  ! NOBS_OUT fields (letkf_tools.f90:440-447, 767-778): work3d(:,:,1:11) as the reference fills it before write_restart --
  ! local observation counts of report types 1, 3, 4, 8, 22, of (REF | RE0 | VR, PHARAD), and their cut-off distances --
  ! for the variable-localisation group of iv3d_t.  pmean = gues3d(:,:,mmean,iv3d_p) BEFORE das_letkf_b200 (which destroys it).
  !-----------------------------------------------------------------------------
  subroutine nobs_out_b200(pmean, work3d)
    use common_scale, only: nlev, iv3d_t
    use common_mpi_scale, only: nij1
    real(c_double), intent(in), target :: pmean(nij1,nlev)
    real(c_double), intent(out), target :: work3d(nij1,nlev,11)
    call check(c_nobs_out(handle, int(iv3d_t, c_int), c_loc(pmean), c_null_ptr, c_loc(work3d), LETKF_B200_MEM_HOST), 'nobs_out')
  end subroutine

  subroutine unsupported(what)
    character(len=*), intent(in) :: what
    write (6, '(2A)') '[Error] letkf_b200: not supported by this binding: ', what
    stop 99
  end subroutine unsupported

  subroutine check(status, what)
    integer(c_int), intent(in) :: status
    character(len=*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    integer :: i
    if (status == LETKF_B200_OK) return
    write (6, '(3A,I4)') '[Error] letkf_b200 ', what, ' failed, status ', status
    if (c_associated(handle)) then
      call c_f_pointer(c_last_error(handle), msg, [256])
      do i = 1, 256
        if (msg(i) == c_null_char) exit
        write (6, '(A)', advance='no') msg(i)
      end do
      write (6, *)
    end if
    stop 2        ! the reference's convention (common_mtx.f90:61-64)
  end subroutine check

end module letkf_b200_iface
