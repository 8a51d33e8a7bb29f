/*
 * oracle_conv.cpp -- CPU restatement of the conventional (prepbufr) observation operator and of monit_obs
 * (TEST INFRASTRUCTURE ONLY, see letkf_oracle.h).
 *
 * Follows, statement by statement (paths relative to the reference root):
 *   scale/obs/obsope_tools.f90:466-473            obsfmt_prepbufr branch of obsope_cal
 *   scale/common/common_obs_scale.f90:999-1110    phys2ijk
 *   scale/common/common_obs_scale.f90:1295-1366   itpl_2d, itpl_2d_column, itpl_3d
 *   scale/common/common_obs_scale.f90:264-337     Trans_XtoY
 *   scale/common/common_obs_scale.f90:600-617     prsadj
 *   scale/common/common_obs_scale.f90:1516-1572   the observation loop of monit_obs (prepbufr and radar formats)
 * MPRJ_rotcoef and state_to_history belong to the un-vendored SCALE-RM library: their results are inputs (rotc, v3dgh / v2dgh).
 */
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "letkf_oracle.h"

namespace {

const double gg = 9.81, rd = 287.05, rv = 461.50, undef = -9.99e33;   /* common/common.f90:29-38 */
const double fvirt = rv / rd - 1.0;                                   /* common/common.f90:34 */
const int iqc_good = 0, iqc_ps_ter = 10, iqc_ref_low = 11, iqc_out_vhi = 20, iqc_out_vlo = 21, iqc_otype = 90, iqc_out_h = 98;
const int id_u_obs = 2819, id_v_obs = 2820, id_t_obs = 3073, id_tv_obs = 3074, id_q_obs = 3330, id_rh_obs = 3331,
          id_ps_obs = 14593;
/* common_scale.f90:66-85 */
const int iv3dd_u = 1, iv3dd_v = 2, iv3dd_t = 4, iv3dd_p = 5, iv3dd_q = 6, iv3dd_rh = 12;
const int iv2dd_topo = 1, iv2dd_ps = 2, iv2dd_t2m = 6, iv2dd_q2m = 7;

struct V3 {   /* var(nlevh, nlonh, nlath), 1-based */
  const double *p;
  int n1, n2;
  double operator()(int a, int b, int c) const { return p[(size_t)(a - 1) + (size_t)n1 * ((size_t)(b - 1) + (size_t)n2 * (size_t)(c - 1))]; }
};
struct V2 {   /* var(nlonh, nlath), 1-based */
  const double *p;
  int n1;
  double operator()(int a, int b) const { return p[(size_t)(a - 1) + (size_t)n1 * (size_t)(b - 1)]; }
};

/* common_obs_scale.f90:1295-1315 */
double itpl_2d(const V2 &var, double ri, double rj) {
  const int i = (int)std::ceil(ri);
  const double ai = ri - (double)(i - 1);
  const int j = (int)std::ceil(rj);
  const double aj = rj - (double)(j - 1);
  return var(i - 1, j - 1) * (1 - ai) * (1 - aj) + var(i, j - 1) * ai * (1 - aj) + var(i - 1, j) * (1 - ai) * aj + var(i, j) * ai * aj;
}

/* common_obs_scale.f90:1339-1366 (called as itpl_3d(var, rk, ri, rj): first argument along the first dimension) */
double itpl_3d(const V3 &var, double ri, double rj, double rk) {
  const int i = (int)std::ceil(ri);
  const double ai = ri - (double)(i - 1);
  const int j = (int)std::ceil(rj);
  const double aj = rj - (double)(j - 1);
  const int k = (int)std::ceil(rk);
  const double ak = rk - (double)(k - 1);
  return var(i - 1, j - 1, k - 1) * (1 - ai) * (1 - aj) * (1 - ak) + var(i, j - 1, k - 1) * ai * (1 - aj) * (1 - ak) +
         var(i - 1, j, k - 1) * (1 - ai) * aj * (1 - ak) + var(i, j, k - 1) * ai * aj * (1 - ak) +
         var(i - 1, j - 1, k) * (1 - ai) * (1 - aj) * ak + var(i, j - 1, k) * ai * (1 - aj) * ak +
         var(i - 1, j, k) * (1 - ai) * aj * ak + var(i, j, k) * ai * aj * ak;
}

/* common_obs_scale.f90:999-1110 */
void phys2ijk(const V3 &p_full, int nlevh, int nlonh, int nlath, int nlev, int KHALO, int elem, double ri, double rj, double rlev,
              double &rk, int &qc) {
  qc = iqc_good;
  if (ri < 1.0 || ri > nlonh || rj < 1.0 || rj > nlath) {   /* :1024-1031 */
    rk = undef;
    qc = iqc_out_h;
    return;
  }
  if (elem > 9999) {   /* surface observation (:1033-1034) */
    rk = rlev;
    return;
  }
  const int i = (int)std::ceil(ri), j = (int)std::ceil(rj);
  int ks = 1 + KHALO;   /* the lowest valid level (:1044-1054) */
  for (int jj = j - 1; jj <= j; ++jj)
    for (int ii = i - 1; ii <= i; ++ii) {
      int k;
      for (k = 1 + KHALO; k <= nlev + KHALO; ++k)
        if (p_full(k, ii, jj) >= 0.0) break;
      if (k > ks) ks = k;
    }
  /* lnps(:,i-1:i,j-1:j) = LOG(p_full(:,i-1:i,j-1:j)); call itpl_2d_column(lnps,ri,rj,plev) (:1067-1068) */
  std::vector<double> plev(nlevh + 1);
  {
    const double ai = ri - (double)(i - 1), aj = rj - (double)(j - 1);
    for (int k = 1; k <= nlevh; ++k)
      plev[k] = std::log(p_full(k, i - 1, j - 1)) * (1 - ai) * (1 - aj) + std::log(p_full(k, i, j - 1)) * ai * (1 - aj) +
                std::log(p_full(k, i - 1, j)) * (1 - ai) * aj + std::log(p_full(k, i, j)) * ai * aj;
  }
  rk = std::log(rlev);   /* :1073 */
  if (rk < plev[nlev + KHALO]) {   /* :1077-1085 */
    rk = undef;
    qc = iqc_out_vhi;
    return;
  }
  if (rk > plev[ks]) {   /* :1086-1099 */
    rk = undef;
    qc = iqc_out_vlo;
    return;
  }
  int k;
  for (k = ks + 1; k <= nlev + KHALO; ++k)   /* :1103-1105 */
    if (plev[k] < rk) break;
  if (k > nlev + KHALO) k = nlev + KHALO;   /* rk == plev(top): the Fortran index runs one past the end (reads a halo level) */
  const double ak = (rk - plev[k - 1]) / (plev[k] - plev[k - 1]);
  rk = (double)(k - 1) + ak;
}

/* common_obs_scale.f90:600-617 */
void prsadj(double &p, double dz, double t, double q) {
  const double gamma = 5.0e-3;
  if (dz != 0) {
    const double tv = t * (1.0 + 0.608 * q);
    p = p * std::pow((-gamma * dz + tv) / tv, gg / (gamma * rd));
  }
}

/* common_obs_scale.f90:264-337 */
void Trans_XtoY(int elm, double ri, double rj, double rk, double rotc1, double rotc2, const double *v3d, const double *v2d, int nlevh,
                int nlonh, int nlath, int stggrd_, double PS_ADJUST_THRES, double &yobs, int &qc) {
  const size_t vsz = (size_t)nlevh * nlonh * nlath, v2sz = (size_t)nlonh * nlath;
  auto var = [&](int iv3dd) { return V3{v3d + vsz * (size_t)(iv3dd - 1), nlevh, nlonh}; };
  auto var2 = [&](int iv2dd) { return V2{v2d + v2sz * (size_t)(iv2dd - 1), nlonh}; };
  yobs = undef;
  qc = iqc_good;
  if (elm == id_u_obs || elm == id_v_obs) {
    double u, v;
    if (stggrd_ == 1) {
      u = itpl_3d(var(iv3dd_u), rk, ri - 0.5, rj);
      v = itpl_3d(var(iv3dd_v), rk, ri, rj - 0.5);
    } else {
      u = itpl_3d(var(iv3dd_u), rk, ri, rj);
      v = itpl_3d(var(iv3dd_v), rk, ri, rj);
    }
    if (elm == id_u_obs) yobs = u * rotc1 - v * rotc2;
    else yobs = u * rotc2 + v * rotc1;
  } else if (elm == id_t_obs) {
    yobs = itpl_3d(var(iv3dd_t), rk, ri, rj);
  } else if (elm == id_tv_obs) {
    yobs = itpl_3d(var(iv3dd_t), rk, ri, rj);
    const double q = itpl_3d(var(iv3dd_q), rk, ri, rj);
    yobs = yobs * (1.0 + fvirt * q);
  } else if (elm == id_q_obs) {
    yobs = itpl_3d(var(iv3dd_q), rk, ri, rj);
  } else if (elm == id_ps_obs) {
    const double t = itpl_2d(var2(iv2dd_t2m), ri, rj);
    const double q = itpl_2d(var2(iv2dd_q2m), ri, rj);
    const double topo = itpl_2d(var2(iv2dd_topo), ri, rj);
    yobs = itpl_2d(var2(iv2dd_ps), ri, rj);
    prsadj(yobs, rk - topo, t, q);
    if (std::fabs(rk - topo) > PS_ADJUST_THRES) qc = iqc_ps_ter;
  } else if (elm == id_rh_obs) {
    yobs = itpl_3d(var(iv3dd_rh), rk, ri, rj);
  } else {
    qc = iqc_otype;
  }
}

}  // namespace

/* obsfmt_prepbufr branch of obsope_cal (obsope_tools.f90:466-473) for all members; arguments as letkf_b200_obsope_conv */
extern "C" void oracle_obsope_conv(const letkf_b200_conv_config *r, int nobs, const int32_t *elm, const double *ril, const double *rjl,
                                   const double *lev, const double *rotc, int nmem, const double *const *v3dgh,
                                   const double *const *v2dgh, int ld_out, double *yobs, int32_t *qc_out) {
  const size_t vsz = (size_t)r->nlevh * r->nlonh * r->nlath;
#pragma omp parallel for schedule(dynamic, 5) collapse(2)
  for (int m = 0; m < nmem; ++m)
    for (int nn = 0; nn < nobs; ++nn) {
      double rk, yo = undef;
      int qc;
      phys2ijk(V3{v3dgh[m] + vsz * (size_t)(iv3dd_p - 1), r->nlevh, r->nlonh}, r->nlevh, r->nlonh, r->nlath, r->nlev, r->KHALO, elm[nn],
               ril[nn], rjl[nn], lev[nn], rk, qc);
      if (qc == iqc_good)
        Trans_XtoY(elm[nn], ril[nn], rjl[nn], rk, rotc ? rotc[nn] : 1.0, rotc ? rotc[(size_t)nobs + nn] : 0.0, v3dgh[m], v2dgh[m],
                   r->nlevh, r->nlonh, r->nlath, r->stggrd, r->PS_ADJUST_THRES, yo, qc);
      yobs[(size_t)nn * ld_out + m] = yo;
      qc_out[(size_t)nn * ld_out + m] = qc;
    }
}

/* the observation loop of monit_obs (common_obs_scale.f90:1516-1572) for one observation set; arguments as
 * letkf_b200_monit_obs_set (host pointers).  The radar operator is oracle_obsope_radar on one state without the RADAR_ZMAX
 * test (monit_obs has none). */
extern "C" void oracle_monit_obs_set(const letkf_b200_conv_config *conv, const letkf_b200_radar_config *radar, int nobs,
                                     const int32_t *elm, const double *ril, const double *rjl, const double *lon, const double *lat,
                                     const double *lev, const double *dat, const double *dif, const double *rotc, double t_range,
                                     const double *v3dgh, const double *v2dgh, double *ohx, int32_t *oqc) {
  std::vector<double> hx((size_t)nobs);
  std::vector<int32_t> q((size_t)nobs);
  if (conv) {
    oracle_obsope_conv(conv, nobs, elm, ril, rjl, lev, rotc, 1, &v3dgh, &v2dgh, 1, hx.data(), q.data());
  } else {
    letkf_b200_radar_config rr = *radar;
    rr.RADAR_ZMAX = 1.0e300;
    oracle_obsope_radar(&rr, nobs, elm, ril, rjl, lon, lat, lev, rotc, 1, &v3dgh, 1, hx.data(), q.data());
    for (int n = 0; n < nobs; ++n)
      if (q[n] == iqc_ref_low) q[n] = iqc_good;   /* :1555 (oracle_obsope_radar already does it like obsope_cal) */
  }
  for (int n = 0; n < nobs; ++n) {
    oqc[n] = -1;   /* :1460 */
    ohx[n] = undef;
    if (t_range <= 0.0 || !dif || std::fabs(dif[n]) <= t_range) {   /* :1516-1517 */
      oqc[n] = q[n];
      if (oqc[n] == iqc_good) ohx[n] = dat[n] - hx[n];   /* :1566-1570 */
      else ohx[n] = undef;
    }
  }
}
