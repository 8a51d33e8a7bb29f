/*
 * oracle_radar.cpp -- CPU restatement of the radar observation operator (TEST INFRASTRUCTURE ONLY, see letkf_oracle.h).
 *
 * Follows, statement by statement (paths relative to the reference root):
 *   scale/obs/obsope_tools.f90:476-494            obsfmt_radar branch of obsope_cal
 *   scale/common/common_obs_scale.f90:1116-1237   phys2ijkz
 *   scale/common/common_obs_scale.f90:1317-1366   itpl_2d_column, itpl_3d
 *   scale/common/common_obs_scale.f90:342-493     Trans_XtoY_radar
 *   scale/common/common_obs_scale.f90:626-990     calc_ref_vr (METHOD_REF_CALC 1, 2, 3)
 *   common/common.f90:401-424, 861-912            com_distll_1, com_gamma
 * MPRJ_rotcoef belongs to the un-vendored SCALE-RM library: its result is an input (rotc).
 * Default-REAL literals of the Fortran are written as (double)<literal>f.
 */
#include <cmath>
#include <cstddef>
#include <cstdint>

#include "letkf_oracle.h"

namespace {

const double pi = 3.1415926535, gg = 9.81, rd = 287.05, re = 6371.3e3, undef = -9.99e33;   /* common/common.f90:28-38 */
const double deg2rad = pi / 180.0, rad2deg = 180.0 / pi;
const int iqc_good = 0, iqc_ref_low = 11, iqc_radar_vhi = 19, iqc_out_vhi = 20, iqc_out_vlo = 21, iqc_otype = 90,
          iqc_out_h = 98;

struct Var {   /* var(nlevh, nlonh, nlath), 1-based */
  const double *p;
  int n1, n2;
  double operator()(int a, int b, int c) const { return p[(size_t)(a - 1) + (size_t)n1 * ((size_t)(b - 1) + (size_t)n2 * (size_t)(c - 1))]; }
};

/* common_obs_scale.f90:1339-1366 */
double itpl_3d(const Var &var, double ri, double rj, double rk) {
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

/* common/common.f90:861-912 */
double com_gamma(double x) {
  static const double G[26] = {1.0, 0.5772156649015329, -0.6558780715202538, -0.420026350340952e-1, 0.1665386113822915,
                               -.421977345555443e-1, -.96219715278770e-2, .72189432466630e-2, -.11651675918591e-2,
                               -.2152416741149e-3, .1280502823882e-3, -.201348547807e-4, -.12504934821e-5,
                               .11330272320e-5, -.2056338417e-6, .61160950e-8, .50020075e-8, -.11812746e-8, .1043427e-9,
                               .77823e-11, -.36968e-11, .51e-12, -.206e-13, -.54e-14, .14e-14, .1e-15};
  double ga;
  if (x == (double)(int)x) {
    if (x > 0.0) {
      ga = 1.0;
      const int m1 = (int)(x - 1);
      for (int k = 2; k <= m1; ++k) ga = ga * k;
    } else {
      ga = 1.0e300;
    }
    return ga;
  }
  double z, r = 1.0;
  int m = 0;
  if (std::fabs(x) > 1.0) {
    z = std::fabs(x);
    m = (int)z;
    for (int k = 1; k <= m; ++k) r = r * (z - k);
    z = z - m;
  } else {
    z = x;
  }
  double gr = G[25];
  for (int k = 25; k >= 1; --k) gr = gr * z + G[k - 1];
  ga = 1.0 / (gr * z);
  if (std::fabs(x) > 1.0) {
    ga = ga * r;
    if (x < 0.0) ga = -pi / (x * ga * std::sin(pi * x));
  }
  return ga;
}

/* common_obs_scale.f90:626-990 */
void calc_ref_vr(int METHOD_REF_CALC, int USE_TERMINAL_VELOCITY, double qv, double qc, double qr, double qci, double qs, double qg,
                 double u, double v, double w, double t, double p, double az, double elev, double &ref, double &vr) {
  (void)qv; (void)qc; (void)qci;
  double zr = 0.0, zs = 0.0, zg = 0.0, zms = 0.0, zmg = 0.0, wt = 0.0;
  ref = 0.0;
  double ro = p / (rd * t);
  if (METHOD_REF_CALC == 1) {
    const double nor = 8.0e6, ror = 1000.0;
    const double pip = std::pow(pi, 1.75);
    const double cf = 10.0e18 * 72;
    const double p0 = 1.0e5;
    const double qt = qr + qs + qg;
    if (qt > 0.0) {
      ref = cf * std::pow(ro * qt, 1.75);
      ref = ref / (pip * std::pow(nor, 0.75) * std::pow(ror, 1.75));
    } else {
      ref = 0.0;
    }
    if (qt > 0.0) {
      const double a = std::pow(p0 / p, (double)0.4f);
      wt = 5.40 * a * std::pow(qt, 0.125);
    } else {
      wt = 0.0;
    }
  } else if (METHOD_REF_CALC == 2) {
    double nor = 8.0e6, nos = 3.0e6, nog = 4.0e4, ror = 1000.0, ros = 100.0, rog = 913.0, roi = 917.0, roo = 1.0;
    const double ki2 = 0.176, kr2 = 0.930;
    const double pip = std::pow(pi, 1.75);
    const double cf = 1.0e18 * 720;
    if (qr > 0.0) {
      zr = cf * std::pow(ro * qr, 1.75);
      zr = zr / (pip * std::pow(nor, 0.75) * std::pow(ror, 1.75));
    }
    if (qs > 0.0) {
      if (t <= (double)273.16f) {
        zs = cf * ki2 * std::pow(ros, 0.25) * std::pow(ro * qs, 1.75);
        zs = zs / (pip * kr2 * std::pow(nos, 0.75) * (roi * roi));
      } else {
        zs = cf * std::pow(ro * qs, 1.75);
        zs = zs / (pip * std::pow(nos, 0.75) * std::pow(roi, 1.75));
      }
    }
    if (qg > 0.0) {
      zg = std::pow(cf / (pip * std::pow(nog, 0.75) * std::pow(rog, 1.75)), (double)0.95f);
      zg = zg * std::pow(ro * qg, (double)1.6625f);
    }
    ref = zr + zs + zg;
    if (ref > 0.0) {
      nor = nor * (double)1e-3f;
      nos = nos * (double)1e-3f;
      nog = nog * (double)1e-3f;
      ror = ror * (double)1e-3f;
      ros = ros * (double)1e-3f;
      rog = rog * (double)1e-3f;
      roo = roo * (double)1e-3f;
      ro = ro * (double)1e-3f;
      const double a = 2115.0, b = 0.8, c = 152.93, d = 0.25, Cd = 0.6;
      const double rofactor = std::pow(roo / ro, 0.25);
      double wr, ws, wg, tmp_factor, lr, ls, lg;
      if (qr > 0.0) {
        tmp_factor = com_gamma(4.0 + b);
        lr = std::pow(pi * ror * nor / (ro * qr), 0.25);
        wr = a * tmp_factor / (6.0 * std::pow(lr, b));
        wr = 1.0e-2 * wr * rofactor;
      } else {
        wr = 0.0;
      }
      if (qs > 0.0) {
        tmp_factor = com_gamma(4.0 + d);
        ls = std::pow(pi * ros * nos / (ro * qs), 0.25);
        ws = c * tmp_factor / (6.0 * std::pow(ls, d));
        ws = 1.0e-2 * ws * rofactor;
      } else {
        ws = 0.0;
      }
      if (qg > 0.0) {
        tmp_factor = com_gamma(4.5);
        lg = std::pow(pi * rog * nog / (ro * qg), 0.25);
        wg = tmp_factor * std::pow((4.0 * gg * 100.0 * rog) / (3.0 * Cd * ro), 0.5);
        wg = 1.0e-2 * wg / (6.0 * std::pow(lg, 0.5));
      } else {
        wg = 0.0;
      }
      wt = (wr * zr + ws * zs + wg * zg) / (zr + zs + zg);
    } else {
      wt = 0.0;
    }
  } else {   /* METHOD_REF_CALC == 3 */
    const double MAXF = 0.5;
    double Fg = 0.0, Fs = 0.0, fwg = 0.0, fws = 0.0;
    if (qr > 0.0 && qg > 0.0) {
      Fg = MAXF * std::pow(std::fmin(qr / qg, qg / qr), 1.0 / 3.0);
      fwg = qr / (qr + qg);
    }
    if (qr > 0.0 && qs > 0.0) {
      Fs = MAXF * std::pow(std::fmin(qr / qs, qs / qr), 1.0 / 3.0);
      fws = qr / (qr + qs);
    }
    const double qrp = (1.0 - Fs - Fg) * qr;
    const double qsp = (1.0 - Fs) * qs;
    const double qgp = (1.0 - Fg) * qg;
    const double qms = Fs * (qr + qs);
    const double qmg = Fg * (qr + qg);
    if (qrp > 0.0) zr = 2.53e4 * std::pow(ro * qrp * 1.0e3, (double)1.84f);
    if (qsp > 0.0) zs = 3.48e3 * std::pow(ro * qsp * 1.0e3, (double)1.66f);
    if (qgp > 0.0) zg = 5.54e3 * std::pow(ro * qgp * 1.0e3, (double)1.70f);
    if (qms > 0.0) {
      zms = ((double)0.00491f + (double)5.75f * fws - (double)5.588f * (fws * fws)) * 1.0e5;
      zms = zms * std::pow(ro * qms * 1.0e3, (double)1.67f - (double)0.202f * fws + (double)0.398f * (fws * fws));
    }
    if (qmg > 0.0) {
      zmg = ((double)0.809f + (double)10.13f * fwg - (double)5.98f * (fwg * fwg)) * 1.0e5;
      zmg = zmg * std::pow(ro * qmg * 1.0e3, (double)1.48f + (double)0.0448f * fwg - (double)0.0313f * (fwg * fwg));
    }
    ref = zr + zg + zs + zms + zmg;
    if (ref > 0.0) {
      const double nor = 8.0e-2, nos = 3.0e-2, nog = 4.0e-4, ror = 1.0, ros = 0.1, rog = 0.917, roo = 0.001;
      ro = 1.0e-3 * ro;
      const double a = 2115.0, b = 0.8, c = 152.93, d = 0.25, Cd = 0.6;
      const double rofactor = std::pow(roo / ro, 0.5);
      double wr, ws, wg, tmp_factor, lr, ls, lg;
      if (qr > 0.0) {
        tmp_factor = com_gamma(4.0 + b);
        lr = std::pow(pi * ror * nor / (ro * qr), 0.25);
        wr = a * tmp_factor / (6.0 * std::pow(lr, b));
        wr = 1.0e-2 * wr * rofactor;
      } else {
        wr = 0.0;
      }
      if (qs > 0.0) {
        ls = std::pow(pi * ros * nos / (ro * qs), 0.25);
        tmp_factor = com_gamma(4.0 + d);
        ws = c * tmp_factor / (6.0 * std::pow(ls, d));
        ws = 1.0e-2 * ws * rofactor;
      } else {
        ws = 0.0;
      }
      if (qg > 0.0) {
        lg = std::pow(pi * rog * nog / (ro * qg), 0.25);
        tmp_factor = com_gamma(4.5);
        wg = tmp_factor * std::pow((4.0 * gg * 100.0 * rog) / (3.0 * Cd * ro), 0.5);
        wg = 1.0e-2 * wg / (6.0 * std::pow(lg, 0.5));
      } else {
        wg = 0.0;
      }
      wt = (wr * zr + ws * zs + ws * zms + wg * zg + wg * zmg) / (zr + zs + zg + zms + zmg);
    } else {
      wt = 0.0;
    }
  }
  vr = u * std::cos(elev * deg2rad) * std::sin(az * deg2rad);
  vr = vr + v * std::cos(elev * deg2rad) * std::cos(az * deg2rad);
  if (USE_TERMINAL_VELOCITY) vr = vr + (w - wt) * std::sin(elev * deg2rad);
  else vr = vr + w * std::sin(elev * deg2rad);
}

}  // namespace

extern "C" void oracle_obsope_radar(const letkf_b200_radar_config *r, int nobs, const int32_t *elm, const double *ril,
                                    const double *rjl, const double *lon, const double *lat, const double *lev,
                                    const double *rotc, int nmem, const double *const *v3dgh, int ld_out, double *yobs,
                                    int32_t *qc_out) {
  const int nlevh = r->nlevh, nlonh = r->nlonh, nlath = r->nlath, nlev = r->nlev, KHALO = r->KHALO;
  const size_t vsz = (size_t)nlevh * nlonh * nlath;
  const double MIN_RADAR_REF = std::pow(10.0, r->MIN_RADAR_REF_DBZ / 10.0);   /* common_obs_scale.f90:251 */
#pragma omp parallel for schedule(dynamic, 5) collapse(2)
  for (int m = 0; m < nmem; ++m)
    for (int nn = 0; nn < nobs; ++nn) {
      auto var = [&](int iv3dd) { return Var{v3dgh[m] + vsz * (size_t)(iv3dd - 1), nlevh, nlonh}; };
      const double ri = ril[nn], rj = rjl[nn], rlev = lev[nn];
      double yo = undef;
      int qc = iqc_good;
      double rkz = undef;
      /* obsope_tools.f90:477-485 */
      if (rlev > r->RADAR_ZMAX) {
        qc = iqc_radar_vhi;
      } else {
        /* phys2ijkz, common_obs_scale.f90:1138-1234 */
        const Var z_full = var(r->nv3dd);   /* iv3dd_hgt = 13 */
        if (ri < 1.0 || ri > (double)nlonh || rj < 1.0 || rj > (double)nlath) {
          qc = iqc_out_h;
        } else {
          const int i = (int)std::ceil(ri), j = (int)std::ceil(rj);
          int ks = 1 + KHALO;
          for (int jj = j - 1; jj <= j; ++jj)
            for (int ii = i - 1; ii <= i; ++ii) {
              int k;
              for (k = 1 + KHALO; k <= nlev + KHALO; ++k)
                if (z_full(k, ii, jj) > -300.0 && z_full(k, ii, jj) < 10000.0) break;
              if (k > ks) ks = k;
            }
          /* itpl_2d_column */
          const double ai = ri - (double)(i - 1), aj = rj - (double)(j - 1);
          auto zlev = [&](int k) {
            return z_full(k, i - 1, j - 1) * (1 - ai) * (1 - aj) + z_full(k, i, j - 1) * ai * (1 - aj) +
                   z_full(k, i - 1, j) * (1 - ai) * aj + z_full(k, i, j) * ai * aj;
          };
          if (rlev > zlev(nlev + KHALO)) {
            qc = iqc_out_vhi;
          } else if (rlev < zlev(ks)) {
            qc = iqc_out_vlo;
          } else {
            int k;
            for (k = ks + 1; k <= nlev + KHALO; ++k)
              if (zlev(k) > rlev) break;
            if (k > nlev + KHALO) k = nlev + KHALO;   /* rlev == zlev(top): the Fortran would index one past the end */
            const double ak = (rlev - zlev(k - 1)) / (zlev(k) - zlev(k - 1));
            rkz = (double)(k - 1) + ak;
          }
        }
      }
      if (qc == iqc_good) {
        /* Trans_XtoY_radar, common_obs_scale.f90:371-484; note the (rk, ri, rj) argument order of the calls */
        double ur = itpl_3d(var(1), rkz, ri, rj);
        double vr = itpl_3d(var(2), rkz, ri, rj);
        const double wr = itpl_3d(var(3), rkz, ri, rj);
        const double tr = itpl_3d(var(4), rkz, ri, rj);
        const double pr = itpl_3d(var(5), rkz, ri, rj);
        const double qvr = itpl_3d(var(6), rkz, ri, rj);
        const double qcr = itpl_3d(var(7), rkz, ri, rj);
        const double qrr = itpl_3d(var(8), rkz, ri, rj);
        const double qir = itpl_3d(var(9), rkz, ri, rj);
        const double qsr = itpl_3d(var(10), rkz, ri, rj);
        const double qgr = itpl_3d(var(11), rkz, ri, rj);
        const double utmp = ur, vtmp = vr;
        const double rotc1 = rotc ? rotc[nn] : 1.0, rotc2 = rotc ? rotc[(size_t)nobs + nn] : 0.0;
        ur = utmp * rotc1 - vtmp * rotc2;
        vr = utmp * rotc2 + vtmp * rotc1;
        const double dlon = lon[nn] - r->radar_lon, dlat = lat[nn] - r->radar_lat;
        if (dlon == 0.0 && dlat == 0.0) {
          qc = iqc_out_h;
        } else {
          double az = rad2deg * std::atan2(dlon * std::cos(r->radar_lat * deg2rad), dlat);
          if (az < 0) az = 360.0 + az;
          /* com_distll_1(lon, lat, radar_lon, radar_lat, dist), common/common.f90:401-424 */
          const double r180 = 1.0 / 180.0;
          const double lon1 = lon[nn] * pi * r180, lon2 = r->radar_lon * pi * r180;
          const double lat1 = lat[nn] * pi * r180, lat2 = r->radar_lat * pi * r180;
          double cosd = std::sin(lat1) * std::sin(lat2) + std::cos(lat1) * std::cos(lat2) * std::cos(lon2 - lon1);
          cosd = std::fmin(1.0, cosd);
          cosd = std::fmax(-1.0, cosd);
          const double dist = std::acos(cosd) * re;
          const double elev = rad2deg * std::atan2(rlev - r->radar_z, dist);
          double radar_ref, radar_rv;
          calc_ref_vr(r->METHOD_REF_CALC, r->USE_TERMINAL_VELOCITY, qvr, qcr, qrr, qir, qsr, qgr, ur, vr, wr, tr, pr, az, elev,
                      radar_ref, radar_rv);
          if (elm[nn] == 4001 || elm[nn] == 4004) {
            if (radar_ref < MIN_RADAR_REF) {
              qc = iqc_ref_low;
              yo = r->MIN_RADAR_REF_DBZ + r->LOW_REF_SHIFT;
            } else {
              yo = 10.0 * std::log10(radar_ref);
            }
          } else if (elm[nn] == 4002) {
            if (radar_ref < MIN_RADAR_REF) qc = iqc_ref_low;
            yo = radar_rv;
          } else {
            qc = iqc_otype;
          }
        }
        if (qc == iqc_ref_low) qc = iqc_good;   /* obsope_tools.f90:489 */
      }
      yobs[(size_t)nn * ld_out + m] = yo;
      qc_out[(size_t)nn * ld_out + m] = qc;
    }
}
