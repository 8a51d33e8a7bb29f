/*
 * letkf_oracle.h -- CPU restatement of the SCALE-LETKF analysis hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and there only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (Fortran 90 + MPI + SCALE-RM) cannot be compiled in
 * this environment (no Fortran front-end, no MPI, SCALE-RM/NetCDF not vendored) and its
 * tree holds no golden vectors, KATs or fixtures for this path (its only test diffs log
 * statistics against files outside the repository, scale/run/test.sh:274-302).  The
 * oracle is therefore pinned by analytic known answers, algebraic invariants and an
 * independent LAPACK (scipy.linalg.eigh) cross-check -- see tests/ and DESIGN.md.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  Arrays are Fortran column-major.
 */
#ifndef LETKF_ORACLE_H
#define LETKF_ORACLE_H

#include "../include/letkf_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* EISPACK restatements, common/netlibrs.f */
double oracle_pythag(double a, double b);                                   /* :1-20   */
void oracle_tred2(int nm, int n, const double *a, double *d, double *e, double *z); /* :520-683 */
int oracle_tql2(int nm, int n, double *d, double *e, double *z);            /* :215-384 */
int oracle_rs(int nm, int n, const double *a, double *w, double *z);        /* :21-79 (matz != 0) */
/* common/common_mtx.f90:41-99; returns nrank_eff, or -1 (rs failed) / -2 (no positive eigenvalue) */
int oracle_mtx_eigen(int n, const double *a, double *eival, double *eivec);
/* common/common_letkf.f90:52-257; optional arguments may be NULL; returns 0 or the mtx_eigen error */
int oracle_letkf_core(int ne, int nobs, int nobsl, const double *hdxb, const double *rdiag,
                      const double *rloc, const double *dep, double *parm_infl, double *trans,
                      double *transm, double *pao, int rdiag_wloc, int infl_update,
                      const double *depd, double *transmd);
/* batched driver over oracle_letkf_core with the layout of letkf_b200_core_batch */
int oracle_core_batch(int ne, int nobs, int npts, const int32_t *nobsl, const double *hdxb,
                      const double *rdiag, const double *rloc, const double *dep,
                      double *parm_infl, double *trans, double *transm, double *pao,
                      int rdiag_wloc, int infl_update, const double *depd, double *transmd,
                      int nthreads);
/* common/common_sort.f90:341-369 / :404-432.  A is 1-based through X (X holds 1-based
 * indices into A, as in Fortran); left/right/K are 1-based positions in X. */
void oracle_quickselect_arg(const double *A, int32_t *X, int left, int right, int K);
void oracle_quickselect_desc_arg(const double *A, int32_t *X, int left, int right, int K);

/* stateful twin of the module state (set_letkf_obs / set_common_mpi_grid / das_letkf) */
typedef struct oracle_state oracle_state;
oracle_state *oracle_create(const letkf_b200_config *cfg);
void oracle_destroy(oracle_state *s);
/* quirk != 0 reproduces ij_obsgrd's use of ngrd_i for the j index (letkf_obs.f90:1200) */
void oracle_set_quirks(oracle_state *s, int ij_obsgrd_quirk);
int oracle_set_obs(oracle_state *s, const letkf_b200_obs *obs);   /* letkf_obs.f90:308-342,660-976 */
int oracle_set_grid(oracle_state *s, int nij1, const double *rig1, const double *rjg1,
                    const double *hgt1);
int oracle_obs_info(const oracle_state *s, int32_t *nobstotal, int32_t *nctype);
int oracle_get_ctype(const oracle_state *s, int ic, letkf_b200_ctype_info *out);
int oracle_get_ac_ext(const oracle_state *s, int ic, int32_t *ac_ext);
int oracle_get_sorted_index(const oracle_state *s, int32_t *sorted_to_orig);
/* letkf_tools.f90:1325-1759 for a batch of points (srch_q0 starts at 1 for each point);
 * brute != 0 replaces the bucket search by an O(nobs) scan of every observation
 * (semantic definition used by test_search_semantic_vs_bruteforce). */
int oracle_obs_local(oracle_state *s, int npts, const double *ri, const double *rj,
                     const double *rlev, const double *rz, int nvar, int32_t *nobsl,
                     int32_t *idx, double *rdiag, double *rloc, int max_out, int brute);
/* letkf_tools.f90:50-932 (live path :130-230, :289-693).  point_mask (nij1,nlev) optional:
 * only points with mask != 0 are analysed (bounded CPU-baseline samples); others untouched. */
int oracle_das_letkf(oracle_state *s, double *gues3d, double *gues2d, double *anal3d,
                     double *anal2d, double *infl3d, double *rtps_infl_out,
                     int32_t *nobsl_out, const uint8_t *point_mask, int nthreads,
                     int64_t *npoints, int64_t *nsolved);
/* NOBS_OUT fields of das_letkf (letkf_tools.f90:281-284, 399-401, 440-447, 767-778) for model variable nvar: out (nij1,nlev,11),
 * exact_hits (nij1,nlev) optional */
void oracle_nobs_out(oracle_state *s, int nvar, const double *pmean, double *out, int32_t *exact_hits, int nthreads);
/* scale/common/common_scale.f90:1513-1552 */
void oracle_ensmean_grd(int mem, int nens, int nij, int nlev, int nv3d, int nv2d, double *v3d,
                        double *v2d);
/* scale/common/common_mpi_scale.f90:264-283, 1428-1476, 1279-1396 (pack/unpack halves) */
void oracle_nij1(int nlon, int nlat, int np, int myrank_e, int32_t *nij1, int32_t *nij1max);
void oracle_grd_to_buf(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np,
                       const double *v3dg, const double *v2dg, double *bufs);
void oracle_buf_to_ens(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np, int myrank_e,
                       int nens, int mstart, int mend, const double *bufr, double *v3d,
                       double *v2d);
void oracle_ens_to_buf(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np, int myrank_e,
                       int nens, int mstart, int mend, const double *v3d, const double *v2d,
                       double *bufs);
void oracle_buf_to_grd(int nlon, int nlat, int nlev, int nv3d, int nv2d, int np,
                       const double *bufr, double *v3dg, double *v2dg);
/* scale/letkf/letkf_obs.f90:355-560: departure + QC half of set_letkf_obs (H08 branch not built) */
void oracle_obs_departure_qc(const letkf_b200_qc_config *q, int member, int det, int nobs, int nensobs,
                             const int32_t *elm, const double *dat, const double *err, int32_t *qc,
                             double *ensval, double *val);
/* scale/common/common_scale.f90:1557-1611 (3-D part) */
void oracle_enssprd_grd(int mem, int nens, int nij, int nlev, int nv3d, const double *v3d, double *v3ds);
/* scale/common/common_scale.f90:1181-1280 */
void oracle_state_trans(const letkf_b200_thermo *t, int inverse, int nlev, int nlon, int nlat, int nv3d,
                        int iv3d_q, double *v3dg);
/* scale/common/common_obs_scale.f90:1851-1895 (serial sums in observation order) */
void oracle_additive_inflation(int mem, int nens, int nij, int nlev, int nv3d, int nv2d, double infl_add, int q_ratio,
                               int ref_only, int iv3d_q, int iv3d_qg, const int32_t *ishuf, double *addi3d, double *addi2d,
                               const double *gues3d, double *anal3d, double *anal2d, const double *rig1, const double *rjg1,
                               int nref, const double *obs_ri, const double *obs_rj, double hloc, double DX, double DY,
                               double dist_zero_fac_square, double *weight);
void oracle_monit_dep(int nn, const int32_t *elm, const double *dep, const int32_t *qc, int32_t *nobs,
                      double *bias, double *rmse);
int oracle_max_threads(void);

/* conventional observation operator for all members and the observation loop of monit_obs (oracle_conv.cpp): obsope_tools.f90:
 * 466-473, common_obs_scale.f90:264-337, 600-617, 999-1110, 1516-1572; arguments as letkf_b200_obsope_conv /
 * letkf_b200_monit_obs_set (host pointers) */
void oracle_obsope_conv(const letkf_b200_conv_config *r, int nobs, const int32_t *elm, const double *ril, const double *rjl,
                        const double *lev, const double *rotc, int nmem, const double *const *v3dgh, const double *const *v2dgh,
                        int ld_out, double *yobs, int32_t *qc);
void oracle_monit_obs_set(const letkf_b200_conv_config *conv, const letkf_b200_radar_config *radar, int nobs, const int32_t *elm,
                          const double *ril, const double *rjl, const double *lon, const double *lat, const double *lev,
                          const double *dat, const double *dif, const double *rotc, double t_range, const double *v3dgh,
                          const double *v2dgh, double *ohx, int32_t *oqc);
/* radar observation operator for all members (oracle_radar.cpp): obsope_tools.f90:476-494, common_obs_scale.f90:342-493,
 * 626-990, 1116-1237; arguments as letkf_b200_obsope_radar (host pointers) */
void oracle_obsope_radar(const letkf_b200_radar_config *r, int nobs, const int32_t *elm, const double *ril, const double *rjl,
                         const double *lon, const double *lat, const double *lev, const double *rotc, int nmem,
                         const double *const *v3dgh, int ld_out, double *yobs, int32_t *qc);

#ifdef __cplusplus
}
#endif
#endif
