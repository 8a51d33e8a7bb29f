/*
 * letkf_b200.h -- C ABI of the B200-native LETKF analysis hot path.
 *
 * Drop-in boundary for the SCALE-LETKF analysis path.  The reference has no FFI:
 * the boundary today is two Fortran module procedures,
 *     letkf_core(ne,nobs,nobsl,hdxb,rdiag,rloc,dep,parm_infl,trans,transm,pao,
 *                rdiag_wloc,infl_update,depd,transmd)      common/common_letkf.f90:52
 *     das_letkf(gues3d,gues2d,anal3d,anal2d)               scale/letkf/letkf_tools.f90:50
 * plus the module state das_letkf reads (grid coordinates, namelist scalars, the
 * bucket-sorted observation tables built by set_letkf_obs, scale/letkf/letkf_obs.f90:78)
 * and the member<->grid transposes scatter/gather_grd_mpi_alltoall
 * (scale/common/common_mpi_scale.f90:1279,1340).
 *
 * Every entry point below replaces one of those; the Fortran ISO_C_BINDING stub
 * a maintainer adds is in scale_letkf_b200/fortran/letkf_b200_iface.f90 and
 * INTEGRATION.md.  All arrays are Fortran column-major, fp64 (REAL(r_size),
 * common/common.f90:18-24) and default 32-bit INTEGER.  Functions return 0 on
 * success and a negative LETKF_B200_E* code otherwise; the library never calls
 * exit()/abort() (the reference STOPs, common/common_mtx.f90:61-64).
 *
 * There is no CPU fallback: every compute entry point runs CUDA kernels built
 * for sm_100a and returns LETKF_B200_ECUDA when no device is usable.
 */
#ifndef LETKF_B200_H
#define LETKF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LETKF_B200_NOBTYPE 24       /* nobtype, scale/common/common_nml.f90:22 */
#define LETKF_B200_NID_OBS 16       /* nid_obs, scale/common/common_nml.f90:21 */
#define LETKF_B200_NID_VARLOCAL 9   /* nid_obs_varlocal, common_obs_scale.f90:43 */
#define LETKF_B200_MAX_NV 16        /* upper bound on nv3d+nv2d (11+0 in the reference) */
#define LETKF_B200_MAX_MEMBER 4096  /* k <= 102: one CTA per point; larger: tiled whole-GPU path */

/* status codes */
#define LETKF_B200_OK 0
#define LETKF_B200_EINVAL (-1)   /* bad argument / unsupported configuration */
#define LETKF_B200_ECUDA (-2)    /* CUDA runtime error or no device          */
#define LETKF_B200_ESTATE (-3)   /* call order (grid/obs not set)            */
#define LETKF_B200_EEIGEN (-4)   /* eigensolve failed at >=1 point (reference: STOP 2) */
#define LETKF_B200_ENOMEM (-5)

/* where the caller's array arguments live */
#define LETKF_B200_MEM_HOST 0
#define LETKF_B200_MEM_DEVICE 1

/*
 * Namelist + module scalars das_letkf reads, names 1:1 with
 * scale/common/common_nml.f90 (PARAM_ENSEMBLE :40-46, PARAM_LETKF :109-142,
 * PARAM_LETKF_OBS :160-218, PARAM_LETKF_VAR_LOCAL :221-229, PARAM_LETKF_RADAR :264).
 * Arrays indexed by report type are 0-based here (type 22 'PHARAD' = index 21).
 * letkf_b200_config_defaults() fills the reference defaults;
 * letkf_b200_config_resolve() applies the "negative inherits element 1" rules of
 * read_nml_letkf_obs (common_nml.f90:741-775).
 */
typedef struct letkf_b200_config {
  /* ensemble */
  int32_t MEMBER;                /* k */
  int32_t DET_RUN;               /* 0/1: deterministic member in slot MEMBER+2 */
  /* grid (single sorting mesh over the whole horizontal plane, PRC 1x1 view) */
  int32_t nlon, nlat, nlev;      /* = nlong, nlatg, nlev */
  int32_t nv3d, nv2d;
  int32_t IHALO, JHALO;
  double DX, DY;
  int32_t iv3d_p, iv3d_q, iv3d_qg; /* 1-based variable indices, common_scale.f90:45-51 */
  /* PARAM_LETKF */
  double INFL_MUL, INFL_MUL_MIN;
  int32_t INFL_MUL_ADAPTIVE;
  int32_t RELAX_TO_INFLATED_PRIOR;
  double RELAX_ALPHA, RELAX_ALPHA_SPREAD;
  double Q_UPDATE_TOP, Q_SPRD_MAX, BOUNDARY_BUFFER_WIDTH;
  /* PARAM_LETKF_OBS */
  double HORI_LOCAL[LETKF_B200_NOBTYPE];
  double VERT_LOCAL[LETKF_B200_NOBTYPE];
  double HORI_LOCAL_RADAR_OBSNOREF, HORI_LOCAL_RADAR_VR, VERT_LOCAL_RADAR_VR;
  double VERT_LOCAL_RAIN_BASE;
  int32_t MAX_NOBS_PER_GRID[LETKF_B200_NOBTYPE];
  int32_t MAX_NOBS_PER_GRID_CRITERION;
  double OBS_MIN_SPACING[LETKF_B200_NOBTYPE];
  double OBS_SORT_GRID_SPACING[LETKF_B200_NOBTYPE];
  /* PARAM_LETKF_VAR_LOCAL: VAR_LOCAL[iv][n]; iv: 0 UV,1 T,2 Q,3 PS,4 RAIN,5 TC,6 RADAR_REF,7 RADAR_VR,8 H08 */
  double VAR_LOCAL[LETKF_B200_NID_VARLOCAL][LETKF_B200_MAX_NV];
  /* PARAM_LETKF_RADAR */
  double RADAR_ZMAX;
  /* letkf_obs.f90:27-28 -- default-REAL literals widened to double; carried as data */
  double dist_zero_fac, dist_zero_fac_square;
  int32_t reserved[8];
} letkf_b200_config;

/* One combined observation type (elm_u, typ) and its sorting mesh:
 * letkf_obs.f90:33-41 (ctype tables) and :47-65 (obs_grid_type). */
typedef struct letkf_b200_ctype_info {
  int32_t elm, elm_u, typ;       /* raw element id, uid_obs(elm) 1..16, report type 1..24 */
  int32_t ngrd_i, ngrd_j, ngrdsch_i, ngrdsch_j, ngrdext_i, ngrdext_j;
  int32_t tot_ext;               /* obs of this ctype                                   */
  int32_t ac_begin;              /* ac_ext(0,1): sorted index of this ctype's first obs  */
  int32_t n_merge;               /* letkf_tools.f90:167-192; 0 = merged into a master    */
  double hori_loc, vert_loc, grdspc_i, grdspc_j;
} letkf_b200_ctype_info;

/* QC-passed observations in their original (unsorted) order: the fields of
 * obs(:)%{elm,typ,lev,dat,err,ri,rj} and obsda%{val,ensval} (common_obs_scale.f90:98-133)
 * that das_letkf/obs_local read.  ensval is ensval(nensobs,nobs), member fastest,
 * already in perturbation form; row MEMBER+1 (DET_RUN) holds y - H(x_det). */
typedef struct letkf_b200_obs {
  int32_t nobs, nensobs;
  const int32_t *elm, *typ;
  const double *ri, *rj, *lev, *dat, *err, *val;
  const double *ensval;
} letkf_b200_obs;

typedef struct letkf_b200_handle letkf_b200_handle;

/* ---- lifetime / configuration -------------------------------------------- */
void letkf_b200_config_defaults(letkf_b200_config *cfg);
void letkf_b200_config_resolve(letkf_b200_config *cfg);
int letkf_b200_create(const letkf_b200_config *cfg, int device, letkf_b200_handle **out);
int letkf_b200_destroy(letkf_b200_handle *h);
const char *letkf_b200_last_error(const letkf_b200_handle *h);
/* run all subsequent work of this handle on a caller-owned cudaStream_t (NULL = default) */
int letkf_b200_set_stream(letkf_b200_handle *h, void *cuda_stream);

/* ---- letkf_core twin (common/common_letkf.f90:52) -------------------------
 * Batched: point i uses hdxb + i*nobs*ne (column-major (nobs,ne), rows 1..nobsl[i]),
 * rdiag/rloc/dep/depd + i*nobs, parm_infl[i]; writes trans + i*ne*ne (column-major),
 * transm/transmd + i*ne, pao + i*ne*ne.  transm, pao, depd, transmd may be NULL
 * (Fortran OPTIONAL absent); when transm is NULL the mean weight is added to every
 * column of trans (:218-226).  npts = 1 reproduces one reference call. */
int letkf_b200_core_batch(letkf_b200_handle *h, int ne, int nobs, int npts,
                          const int32_t *nobsl, const double *hdxb, const double *rdiag,
                          const double *rloc, const double *dep, double *parm_infl,
                          double *trans, double *transm, double *pao, int rdiag_wloc,
                          int infl_update, const double *depd, double *transmd,
                          int mem_space);

/* ---- module state das_letkf reads ----------------------------------------- */
/* rig1(nij1), rjg1(nij1), hgt1(nij1,nlev): common_mpi_scale.f90:40-42,303-308 */
int letkf_b200_set_grid(letkf_b200_handle *h, int nij1, const double *rig1,
                        const double *rjg1, const double *hgt1, int mem_space);
/* Twin of the bucket-sort half of set_letkf_obs (letkf_obs.f90:308-342 ctype table,
 * :660-695 mesh, :747-805 counting sort, :922-976 extended prefix sums).  Host arrays. */
int letkf_b200_set_obs(letkf_b200_handle *h, const letkf_b200_obs *obs);
int letkf_b200_obs_info(const letkf_b200_handle *h, int32_t *nobstotal, int32_t *nctype);
int letkf_b200_get_ctype(const letkf_b200_handle *h, int ic, letkf_b200_ctype_info *out);
/* ac_ext(0:ngrdext_i, 1:ngrdext_j) of ctype ic, column-major, chained across ctypes */
int letkf_b200_get_ac_ext(const letkf_b200_handle *h, int ic, int32_t *ac_ext);
/* sorted position -> original obs index (0-based): the order of obsda_sort */
int letkf_b200_get_sorted_index(const letkf_b200_handle *h, int32_t *sorted_to_orig);

/* ---- departure + QC half of set_letkf_obs (scale/letkf/letkf_obs.f90:355-560) -----------------
 * Per observation n with qc[n] == 0 on entry:  radar acceptance rules (USE_RADAR_REF / USE_RADAR_VR,
 * undef data, number of members above RADAR_REF_THRES_DBZ, :370-421), ensemble mean of H(x) (sequential
 * sum m = 1..MEMBER then divide, :474-478), perturbations ensval(m,n) -= mean (:486-488), departure
 * val[n] = dat - mean (:489), deterministic departure ensval(MEMBER+1,n) = dat - ensval(MEMBER+1,n) (:490-492),
 * gross-error check |val| > GROSS_ERROR_x * err (:503-549).  qc is rewritten with the reference's codes
 * (iqc_good 0, iqc_gross_err 5, iqc_ref_mem 12, iqc_obs_bad 50, iqc_otype 90, common_obs_scale.f90:139-151).
 * Observations that fail a check before the departure step keep their ensval/val untouched, as in the
 * reference (`cycle`).  The Himawari-8 branch (#ifdef H08) is not built.  ensval is (nensobs, nobs),
 * member fastest.  The QC-passed observations (qc == 0) are what letkf_b200_set_obs takes. */
typedef struct letkf_b200_qc_config {
  double GROSS_ERROR, GROSS_ERROR_RAIN, GROSS_ERROR_RADAR_REF, GROSS_ERROR_RADAR_VR, GROSS_ERROR_RADAR_PRH;
  double GROSS_ERROR_TCX, GROSS_ERROR_TCY, GROSS_ERROR_TCP;      /* common_nml.f90:129-137; < 0: GROSS_ERROR */
  double RADAR_REF_THRES_DBZ;                                     /* common_nml.f90:257 */
  int32_t USE_RADAR_REF, USE_RADAR_VR;                            /* common_nml.f90:248-249 */
  int32_t MIN_RADAR_REF_MEMBER, MIN_RADAR_REF_MEMBER_OBSREF;      /* common_nml.f90:258-259 */
} letkf_b200_qc_config;
void letkf_b200_qc_config_defaults(letkf_b200_qc_config *q);
int letkf_b200_obs_departure_qc(letkf_b200_handle *h, const letkf_b200_qc_config *q, int nobs, int nensobs,
                                const int32_t *elm, const double *dat, const double *err, int32_t *qc,
                                double *ensval, double *val, int mem_space);

/* ---- monit_dep twin (scale/common/common_obs_scale.f90:1851-1895): departure statistics per observed
 * element uid (1..16, Tv counted as T, RE0 as REF) over the observations with qc == 0: count, bias = mean(dep),
 * rmse = sqrt(mean(dep^2)); undef (-9.99e33) where the count is zero.  The reference sums serially in
 * observation order; here the sums are a fixed-shape tree (deterministic, rounding differs: tested at 1e-12). */
int letkf_b200_monit_dep(letkf_b200_handle *h, int nobs, const int32_t *elm, const double *dep, const int32_t *qc,
                         int32_t *nobs_out, double *bias, double *rmse, int mem_space);

/* ---- obs_local twin (letkf_tools.f90:1325) ---------------------------------
 * For npts points (ri,rj,rlev=mean pressure,rz=height) and model variable nvar
 * (1-based, 0 = no variable localisation): writes nobsl[i] and, when not NULL, the
 * selected sorted-obs indices (0-based) idx + i*max_out, rdiag/rloc + i*max_out.
 * Order of the list is unspecified (the reference's is quickselect order). */
int letkf_b200_obs_local(letkf_b200_handle *h, int npts, const double *ri, const double *rj,
                         const double *rlev, const double *rz, int nvar, int32_t *nobsl,
                         int32_t *idx, double *rdiag, double *rloc, int max_out,
                         int mem_space);

/* ---- NOBS_OUT fields of das_letkf (letkf_tools.f90:281-284, 399-401, 440-447, 767-778) ------
 * obs_local's optional outputs nobsl_t / cutd_t (:1342-1343) at every analysis point (set_grid), for the variable-localisation
 * group of the 3-D model variable nvar (1-based; the reference writes those of iv3d_t), reduced to the eleven fields the
 * reference writes to NOBS_OUT_BASENAME, out(nij1,nlev,11):
 *   1..5  number of local observations of report types 1 (ADPUPA), 3, 4, 8 and 22 (PHARAD), summed over the elements
 *   6..8  nobsl_t(REF | RE0 | VR, PHARAD)        9..11  cutd_t(REF | RE0 | VR, PHARAD)
 * Points with relax_beta = 0 keep zeros.  pmean = gues3d(:,:,mmean,iv3d_p) (nij1,nlev); logp optional as in das_letkf (host
 * buffers: ln p is always taken by the host libm).  Counts are identical to the reference's, including its quirk that an
 * unlimited merged group (REF + RE0) reports cumulative counts.  cutd_t is the criterion value of the worst selected
 * observation of a group that filled MAX_NOBS_PER_GRID -- what the reference's quickselect leaves in the last slot -- and
 * differs from the reference only when its last incremental search pass found EXACTLY the limit (no quickselect: the
 * reference then reports the last SCANNED observation, which depends on the level-to-level search_q0 history). */
int letkf_b200_nobs_out(letkf_b200_handle *h, int nvar, const double *pmean, const double *logp, double *out, int mem_space);

/* ---- das_letkf twin (letkf_tools.f90:50) ----------------------------------- */
typedef struct letkf_b200_das_args {
  double *gues3d;        /* (nij1,nlev,nens,nv3d) INOUT: destroyed -> perturbations, slot MEMBER+1 = mean */
  double *gues2d;        /* (nij1,nens,nv2d) or NULL when nv2d == 0 */
  double *anal3d;        /* (nij1,nlev,nens,nv3d) OUT: slots 1..MEMBER (+ mmdet) */
  double *anal2d;
  double *infl3d;        /* optional (nij1,nlev,nv3d) work3d: in = inflation field when INFL_MUL<=0,
                            out = adaptively updated field when INFL_MUL_ADAPTIVE; NULL = INFL_MUL */
  double *rtps_infl_out; /* optional (nij1,nlev,nv3d): work3da, RELAX_SPREAD_OUT (letkf_tools.f90:271-276) */
  int32_t *nobsl_out;    /* optional (nij1,nlev): local obs count of variable group 1 (NOBS_OUT) */
  const double *logp;    /* optional (nij1,nlev): log(mean pressure) precomputed by the host libm so that
                            selection thresholds are bit-identical to a CPU run; NULL = device log() */
  int32_t mem_space;     /* LETKF_B200_MEM_HOST: arrays are host memory (H2D/D2H inside the call) */
  int32_t reserved;
} letkf_b200_das_args;
int letkf_b200_das_letkf(letkf_b200_handle *h, const letkf_b200_das_args *args);
/* counters of the last das call: points analysed, points solved (nobsl>0), eigensolve failures */
int letkf_b200_das_stats(const letkf_b200_handle *h, int64_t *npoints, int64_t *nsolved,
                         int64_t *nfail, int64_t *nobsl_sum);
/* solves of the last das call that took the ill-conditioned path (lambda_max(A) / c0 above ~1e4: explicit A^-1/2 and one
 * step of iterative refinement of the mean weight against the original matrix; MEMBER <= 102 only) */
int letkf_b200_das_refined(const letkf_b200_handle *h, int64_t *nrefined);
/* CUDA-event time (ms) of the dominant kernel in the last das call (bench roofline) */
int letkf_b200_das_kernel_ms(const letkf_b200_handle *h, float *analysis_ms, int *launches);

/* profiling aid: SM-clock cycles summed over CTAs (thread 0) per phase of the last das call --
 * [0] load+perturbation, [1] local-obs search, [2] Gram, [3] factorisation, [4] eigen/f(A) iteration,
 * [5] apply (G^T x, scalars, W dx), [6] relaxation+store, [7] scheduling -- and the total number of
 * solver sweeps/iterations.  (The reference's equivalent is mpi_timer, common_mpi_scale.f90:1971.) */
int letkf_b200_das_phase_clocks(const letkf_b200_handle *h, int64_t *clocks, int64_t *solver_iterations);

/* ---- ensmean_grd twin (common_scale.f90:1513) ------------------------------ */
int letkf_b200_ensmean_grd(letkf_b200_handle *h, int mem, int nens, int nij, double *v3d,
                           double *v2d, int mem_space);

/* ---- enssprd_grd twin (common_scale.f90:1557-1611): spread over members 1..mem around slot mem+1 (which
 * must already hold the mean): v3ds(nij,nlev,nv3d), v2ds(nij,nv2d) = sqrt(sum_m (x_m - mean)^2 / (mem-1)),
 * summed in member order. */
int letkf_b200_enssprd_grd(letkf_b200_handle *h, int mem, int nens, int nij, const double *v3d,
                           const double *v2d, double *v3ds, double *v2ds, int mem_space);

/* ---- state_trans / state_trans_inv twins (scale/common/common_scale.f90:1181-1280) -------------
 * In place on ONE member-major grid v3dg(nlev,nlon,nlat,nv3d): SCALE restart variables
 * (rho, rho u, rho v, rho w, rho theta, q...) <-> LETKF state variables (u, v, w, T, p, q...), including
 * the POSITIVE_DEFINITE_Q / _QHYD clamps of the inverse (:1243-1250).  The thermodynamic constants come
 * from the SCALE-RM library (scale_const, scale_tracer -- not vendored in the reference tree), so they
 * are carried as data; letkf_b200_thermo_defaults() fills SCALE-RM's usual values. */
typedef struct letkf_b200_thermo {
  double Rdry, Rvap, CVdry, PRE00;
  double TRACER_CV[LETKF_B200_MAX_NV];   /* CV of moisture variable iv3d_q + i (vapour, cloud, rain, ice, snow, graupel) */
  int32_t POSITIVE_DEFINITE_Q, POSITIVE_DEFINITE_QHYD;   /* common_nml.f90 PARAM_LETKF */
} letkf_b200_thermo;
void letkf_b200_thermo_defaults(letkf_b200_thermo *t);
int letkf_b200_state_trans(letkf_b200_handle *h, const letkf_b200_thermo *t, int inverse, double *v3dg,
                           int mem_space);

/* ---- member<->grid transposes (common_mpi_scale.f90:1279-1476) --------------
 * pack:   v3dg(nlev,nlon,nlat,nv3d), v2dg(nlon,nlat,nv2d) of ONE member ->
 *         bufs(nij1max,nlevall,np) with the cyclic column deal of grd_to_buf (:1428)
 * unpack: bufr(nij1max,nlevall,mcount) received from the member owners ->
 *         v3d(1:nij1,:,mstart:mend,:), v2d(1:nij1,mstart:mend,:)
 * and the two reverse directions for gather_grd_mpi_alltoall.  The exchange between
 * pack and unpack is an NCCL all-to-all issued by the caller on the same stream.
 * Device pointers only. */
int letkf_b200_grd_to_buf(letkf_b200_handle *h, int np, const double *v3dg, const double *v2dg,
                          double *bufs);
int letkf_b200_buf_to_ens(letkf_b200_handle *h, int np, int myrank_e, int nens, int mstart,
                          int mend, const double *bufr, double *v3d, double *v2d);
int letkf_b200_ens_to_buf(letkf_b200_handle *h, int np, int myrank_e, int nens, int mstart,
                          int mend, const double *v3d, const double *v2d, double *bufs);
int letkf_b200_buf_to_grd(letkf_b200_handle *h, int np, const double *bufr, double *v3dg,
                          double *v2dg);
/* pack with state_trans fused in front (restart variables -> LETKF variables while packing), unpack with
 * state_trans_inv fused behind (common_scale.f90:1181-1280 inside common_mpi_scale.f90:1428-1476); t == NULL: plain */
int letkf_b200_grd_to_buf_trans(letkf_b200_handle *h, int np, const letkf_b200_thermo *t, const double *v3dg,
                                const double *v2dg, double *bufs);
int letkf_b200_buf_to_grd_trans(letkf_b200_handle *h, int np, const letkf_b200_thermo *t, const double *bufr,
                                double *v3dg, double *v2dg);
/* nij1 / nij1max of rank myrank_e among np ranks (set_common_mpi_grid :264-283) */
int letkf_b200_nij1(const letkf_b200_handle *h, int np, int myrank_e, int32_t *nij1,
                    int32_t *nij1max);

/* sizeof() of the four public structs, in declaration order (config, ctype_info, obs, das_args;
 * letkf_b200_abi_size_qc() returns sizeof(letkf_b200_qc_config)):
 * lets a foreign-language binding (ctypes, ISO_C_BINDING) check its mirror of the layouts */
void letkf_b200_abi_sizes(int32_t sizes[4]);
int letkf_b200_abi_size_qc(void);

/* library build info (arch string, e.g. "sm_100a") */
const char *letkf_b200_build_info(void);

/* ---- one-pass member<->grid transposes over peer memory ---------------------
 * Twins of scatter_grd_mpi_alltoall / gather_grd_mpi_alltoall (common_mpi_scale.f90:1279-1396) with the exchange
 * INSIDE the kernel: every rank reads its own arrays and writes straight into the receiving rank's array -- its own
 * memory or a peer GPU's over NVLink -- so grd_to_buf, MPI_ALLTOALL(V) and buf_to_grd (:1428-1476) are one pass.
 * Ranks of one node, one process per GPU.  Peer arrays are made addressable with CUDA IPC:
 *   peer_export  on the owner: 64-byte handle + offset of a device pointer (send it to the other ranks: MPI_Allgather,
 *                torch.distributed.all_gather_object, ...)
 *   peer_open    on every other rank: map it, get a device pointer valid on this rank's GPU (cached per handle)
 * Synchronisation is the caller's, exactly as around the reference's blocking collective: all ranks must have
 * finished the consumers of the destination arrays before the call (barrier), and the destination is complete
 * once every rank's call has completed on its stream (stream synchronise, then barrier).
 *   scatter: this rank holds the member that becomes slot `mslot` (1-based) as v3dg(nlev,nlon,nlat,nv3d) [v2dg(nlon,
 *            nlat,nv2d)]; peer_v3d[m] = v3d(nij1_m,nlev,nens,nv3d) of rank m, m = 0..np-1 (own array included).  Ranks
 *            without a member in this round do not call.
 *   gather:  members mstart..mend (1-based; member mstart+q lives on rank q) leave this rank's v3d for
 *            peer_v3dg[q] = v3dg of rank q, q = 0..mend-mstart.  Every rank calls.
 * t != NULL applies state_trans (scatter) / state_trans_inv (gather) on the fly (common_scale.f90:1181-1280). */
typedef struct letkf_b200_ipc {
  unsigned char handle[64];
  uint64_t offset;
} letkf_b200_ipc;
int letkf_b200_peer_export(letkf_b200_handle *h, const void *devptr, letkf_b200_ipc *out);
int letkf_b200_peer_open(letkf_b200_handle *h, const letkf_b200_ipc *in, void **mapped);
int letkf_b200_scatter_grd_p2p(letkf_b200_handle *h, int np, int myrank_e, int nens, int mslot,
                               const letkf_b200_thermo *t, const double *v3dg, const double *v2dg,
                               double *const *peer_v3d, double *const *peer_v2d);
int letkf_b200_gather_grd_p2p(letkf_b200_handle *h, int np, int myrank_e, int nens, int mstart, int mend,
                              const letkf_b200_thermo *t, const double *v3d, const double *v2d,
                              double *const *peer_v3dg, double *const *peer_v2dg);

/* ---- device-resident observation chain -------------------
 * set_obs_device: like set_obs, but every array of `obs` (and qc) is a DEVICE array of ALL observations as the
 * observation operator (letkf_b200_obsope_radar) and the departure/QC kernel (letkf_b200_obs_departure_qc) leave them;
 * the qc == iqc_good filter of the counting sort (letkf_obs.f90:752, 791), the combined-type lookup (:300-342) and
 * the vertical coordinate are done on the device.  qc may be NULL (all accepted).  *nkept = accepted observations;
 * get_kept_index returns their indices into the input arrays (arrival order), so that sorted_index can be mapped
 * back.  ln(lev) of observations localised in ln p is CUDA's log() here (<= 1 ulp of the host's): radar
 * observations (localised in height) are unaffected. */
int letkf_b200_set_obs_device(letkf_b200_handle *h, const letkf_b200_obs *obs, const int32_t *qc, int32_t *nkept);
int letkf_b200_get_kept_index(const letkf_b200_handle *h, int32_t *kept);

/* ---- additive inflation block of das_letkf (scale/letkf/letkf_tools.f90:804-929) --------
 * anal(i,k,m,n) += (addi(i,k,ishuf(m),n) - mean_m addi(i,k,:,n)) * INFL_ADD * addinfl_weight(i)
 *                  [* gues3d(i,k,mmean,n) for the moisture variables iv3d_q..iv3d_qg when INFL_ADD_Q_RATIO]
 *   addi3d/2d  the additive-inflation ensemble as read_ens_mpi_addiinfl delivers it, laid out like gues3d/gues2d
 *              (members 1..MEMBER are read; the arrays are NOT modified: the reference turns them into perturbations in place,
 *              nothing reads them afterwards)
 *   gues3d     background array of das_letkf: only its mean slot is read, only with q_ratio (may be NULL otherwise)
 *   ishuf      INFL_ADD_SHUFFLE: permutation of 1..MEMBER drawn by the caller (Knuth_Shuffle + MPI_BCAST in the reference);
 *              NULL = identity
 *   ref_only   INFL_ADD_REF_ONLY: addinfl_weight(i) = exp(-d^2/2) of the nearest radar-reflectivity observation (combined
 *              type REF/PHARAD of the tables of the last set_obs) in units of its localisation scale, 0 beyond dist_zero_fac;
 *              otherwise 1.  weight_out (nij1, optional) receives the weights.
 * Arithmetic in the reference's order without FMA contraction: identical to the CPU except for exp() (<= 1 ulp). */
int letkf_b200_additive_inflation(letkf_b200_handle *h, double infl_add, int q_ratio, int ref_only, const int32_t *ishuf,
                                  const double *addi3d, const double *addi2d, const double *gues3d, double *anal3d,
                                  double *anal2d, double *weight_out, int mem_space);

/* ---- radar observation operator ----------------------
 * Twin of the obsfmt_radar branch of obsope_cal (scale/obs/obsope_tools.f90:476-494): phys2ijkz (scale/common/
 * common_obs_scale.f90:1116-1237), Trans_XtoY_radar (:342-493) and calc_ref_vr (:626-990, METHOD_REF_CALC 1/2/3), for ALL
 * members at once, so that H(x_m) of a 30-second radar volume never leaves the GPU.
 *   v3dgh[m]   member m's history grid WITH halos, v3dg(nlevh,nlonh,nlath,nv3dd) as read_ens_history_iter delivers
 *              it (common_scale.f90:66-78: u,v,w,t,p,q,qc,qr,qi,qs,qg,rh,hgt), level fastest
 *   ril, rjl   observation position in this subdomain's halo'ed index space (rij_g2l), lon/lat/lev as in obs(iof)
 *   rotc       [2][nobs] MPRJ_rotcoef(lon,lat) of the SCALE-RM map projection (host library, un-vendored); NULL = (1,0)
 *   yobs, qc   [nobs][ld_out], member fastest (the ensval(nensobs,nobs) layout of obs_da_value): H(x_m) and its QC flag
 *              (iqc_good = 0, iqc_radar_vhi = 19, iqc_out_vhi = 20, iqc_out_vlo = 21, iqc_otype = 90, iqc_out_h = 98;
 *              iqc_ref_low is reset to good exactly as obsope_cal does, :489) */
typedef struct letkf_b200_radar_config {
  int32_t METHOD_REF_CALC;        /* common_nml.f90:270 (default 3) */
  int32_t USE_TERMINAL_VELOCITY;  /* :272 */
  int32_t nlevh, nlonh, nlath, nlev, KHALO, nv3dd;
  double MIN_RADAR_REF_DBZ;       /* :261 */
  double LOW_REF_SHIFT;           /* :262 */
  double RADAR_ZMAX;              /* common_nml.f90 PARAM_LETKF_RADAR */
  double radar_lon, radar_lat, radar_z;   /* obs(iof)%meta(1:3) */
} letkf_b200_radar_config;
void letkf_b200_radar_config_defaults(letkf_b200_radar_config *r);
int letkf_b200_obsope_radar(letkf_b200_handle *h, const letkf_b200_radar_config *r, int nobs, const int32_t *elm,
                            const double *ril, const double *rjl, const double *lon, const double *lat,
                            const double *lev, const double *rotc, int nmem, const double *const *v3dgh,
                            int ld_out, double *yobs, int32_t *qc, int mem_space);

/* ---- conventional (prepbufr) observation operator ----------------------
 * Twin of the obsfmt_prepbufr branch of obsope_cal (scale/obs/obsope_tools.f90:466-473) and of monit_obs (scale/common/
 * common_obs_scale.f90:1530-1540): phys2ijk (:999-1110, pressure -> level index by linear interpolation in ln p; surface
 * observations keep their height) and Trans_XtoY (:264-337: U, V with the map-projection rotation, T, Tv, Q, RH by tri-linear
 * interpolation; PS from the 2-D fields with the height adjustment prsadj :600-617 and the PS_ADJUST_THRES test), for ALL
 * members in one launch.  v3dgh[m] = v3dgh(nlevh,nlonh,nlath,nv3dd) of member m in the reference's history-variable order
 * (u v w t p q qc qr qi qs qg rh hgt), v2dgh[m] = v2dgh(nlonh,nlath,nv2dd) (topo ps rain u10m v10m t2m q2m); ril/rjl: local
 * grid coordinates (rij_g2l); rotc(nobs,2) = MPRJ_rotcoef of SCALE-RM's map projection at (lon, lat) -- carried as data,
 * NULL = (1, 0).  yobs/qc (ld_out, nobs): member fastest.  qc codes of the reference (iqc_good 0, iqc_ps_ter 10,
 * iqc_out_vhi 20, iqc_out_vlo 21, iqc_otype 90, iqc_out_h 98).  ln() is CUDA's (<= 1 ulp of the host's): tested at 1e-12. */
typedef struct letkf_b200_conv_config {
  int32_t nlevh, nlonh, nlath, nlev, KHALO, nv3dd, nv2dd;
  int32_t stggrd;                 /* 1: u, v on the staggered grid (monit_obs calls Trans_XtoY with stggrd = 1) */
  double PS_ADJUST_THRES;         /* common_nml.f90:148 */
} letkf_b200_conv_config;
void letkf_b200_conv_config_defaults(letkf_b200_conv_config *c);
int letkf_b200_obsope_conv(letkf_b200_handle *h, const letkf_b200_conv_config *r, int nobs, const int32_t *elm, const double *ril,
                           const double *rjl, const double *lev, const double *rotc, int nmem, const double *const *v3dgh,
                           const double *const *v2dgh, int ld_out, double *yobs, int32_t *qc, int mem_space);

/* ---- monit_obs twin (scale/common/common_obs_scale.f90:1370-1844), one observation set (file) per call ----------------
 * H(x) of ONE state (the mean background / analysis as history variables v3dgh, v2dgh) at the observations of one input
 * file -- conv != NULL: prepbufr format (phys2ijk + Trans_XtoY, the caller sets conv->stggrd = 1 like monit_obs);
 * radar != NULL: radar format (phys2ijkz + Trans_XtoY_radar, iqc_ref_low counts as good, no RADAR_ZMAX test; skip the call
 * when DEPARTURE_STAT_RADAR is off) -- then ohx = dat - H(x) where the operator's qc is good, undef elsewhere (:1566-1570).
 * Observations with |dif| > t_range (DEPARTURE_STAT_T_RANGE > 0; dif may be NULL) keep oqc = -1.  The statistics over all
 * sets are letkf_b200_monit_dep on the concatenated (elm, ohx, oqc) -- exactly the reference's last step (:1808). */
int letkf_b200_monit_obs_set(letkf_b200_handle *h, const letkf_b200_conv_config *conv, const letkf_b200_radar_config *radar, int nobs,
                             const int32_t *elm, const double *ril, const double *rjl, const double *lon, const double *lat,
                             const double *lev, const double *dat, const double *dif, const double *rotc, double t_range,
                             const double *v3dgh, const double *v2dgh, double *ohx, int32_t *oqc, int mem_space);

#ifdef __cplusplus
}
#endif
#endif /* LETKF_B200_H */
