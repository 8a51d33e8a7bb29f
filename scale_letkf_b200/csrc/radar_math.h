// radar_math.h -- radar observation operator arithmetic, written ONCE for the CUDA kernel (radar.cuh).
// Follows scale/common/common_obs_scale.f90: itpl_3d :1339-1366, itpl_2d_column :1317-1337, phys2ijkz :1116-1237,
// Trans_XtoY_radar :342-493, calc_ref_vr :626-990; common/common.f90: com_distll_1 :401-424, com_gamma :861-912.
// (oracle/oracle_radar.cpp is an independent restatement of the same Fortran, not an include of this file.)
#pragma once
#include <math.h>

#ifndef LETKF_HD
#ifdef __CUDACC__
#define LETKF_HD __host__ __device__ __forceinline__
#else
#define LETKF_HD inline
#endif
#endif

namespace letkf_radar {

constexpr double kPi = 3.1415926535;          // common/common.f90:28 (the reference's 11-digit pi)
constexpr double kGG = 9.81, kRd = 287.05, kRe = 6371.3e3, kUndef = -9.99e33;
constexpr double kDeg2Rad = kPi / 180.0, kRad2Deg = 180.0 / kPi;
enum { IQC_GOOD = 0, IQC_REF_LOW = 11, IQC_RADAR_VHI = 19, IQC_OUT_VHI = 20, IQC_OUT_VLO = 21, IQC_OTYPE = 90, IQC_OUT_H = 98 };
enum { ID_REF = 4001, ID_RE0 = 4004, ID_VR = 4002 };

struct Grid {   // one member's v3dg(nlevh, nlonh, nlath, nv3dd), level fastest
  const double *v;
  int nlevh, nlonh, nlath;
  LETKF_HD double at(int k, int i, int j, int n) const {   // 1-based like the Fortran
    return v[(size_t)(k - 1) + (size_t)nlevh * ((size_t)(i - 1) + (size_t)nlonh * ((size_t)(j - 1) + (size_t)nlath * n))];
  }
};

// itpl_3d(var, ri, rj, rk): NOTE the callers pass (rk, ri, rj): the first coordinate runs along the first array dimension
LETKF_HD double itpl_3d(const Grid &g, int n, double r1, double r2, double r3) {
  const int i = (int)ceil(r1), j = (int)ceil(r2), k = (int)ceil(r3);
  const double ai = r1 - (double)(i - 1), aj = r2 - (double)(j - 1), ak = r3 - (double)(k - 1);
  return g.at(i - 1, j - 1, k - 1, n) * (1 - ai) * (1 - aj) * (1 - ak) + g.at(i, j - 1, k - 1, n) * ai * (1 - aj) * (1 - ak) +
         g.at(i - 1, j, k - 1, n) * (1 - ai) * aj * (1 - ak) + g.at(i, j, k - 1, n) * ai * aj * (1 - ak) +
         g.at(i - 1, j - 1, k, n) * (1 - ai) * (1 - aj) * ak + g.at(i, j - 1, k, n) * ai * (1 - aj) * ak +
         g.at(i - 1, j, k, n) * (1 - ai) * aj * ak + g.at(i, j, k, n) * ai * aj * ak;
}

// one level of itpl_2d_column
LETKF_HD double zcol(const Grid &g, int n, int k, int i, int j, double ai, double aj) {
  return g.at(k, i - 1, j - 1, n) * (1 - ai) * (1 - aj) + g.at(k, i, j - 1, n) * ai * (1 - aj) +
         g.at(k, i - 1, j, n) * (1 - ai) * aj + g.at(k, i, j, n) * ai * aj;
}

// phys2ijkz: height -> fractional level index; returns the QC flag
LETKF_HD int phys2ijkz(const Grid &g, int ihgt, int nlev, int khalo, double ri, double rj, double rlev, double &rk) {
  rk = kUndef;
  if (ri < 1.0 || ri > (double)g.nlonh || rj < 1.0 || rj > (double)g.nlath) return IQC_OUT_H;
  const int i = (int)ceil(ri), j = (int)ceil(rj);
  int ks = 1 + khalo;
  for (int jj = j - 1; jj <= j; ++jj)
    for (int ii = i - 1; ii <= i; ++ii) {
      int k = 1 + khalo;
      for (; k <= nlev + khalo; ++k) {
        const double z = g.at(k, ii, jj, ihgt);
        if (z > -300.0 && z < 10000.0) break;
      }
      if (k > ks) ks = k;
    }
  const double ai = ri - (double)(i - 1), aj = rj - (double)(j - 1);
  if (rlev > zcol(g, ihgt, nlev + khalo, i, j, ai, aj)) return IQC_OUT_VHI;
  if (rlev < zcol(g, ihgt, ks, i, j, ai, aj)) return IQC_OUT_VLO;
  int k = ks + 1;
  double zk = 0.0;
  for (; k <= nlev + khalo; ++k) {
    zk = zcol(g, ihgt, k, i, j, ai, aj);
    if (zk > rlev) break;   // assuming ascending order of zlev
  }
  if (k > nlev + khalo) {   // (the Fortran loop index runs one past the end: zlev(nlev+KHALO+1) would be read)
    k = nlev + khalo;
    zk = zcol(g, ihgt, k, i, j, ai, aj);
  }
  const double zkm = zcol(g, ihgt, k - 1, i, j, ai, aj);
  rk = (double)(k - 1) + (rlev - zkm) / (zk - zkm);
  return IQC_GOOD;
}

LETKF_HD double com_gamma(double x) {   // common/common.f90:861-912 (x > 1, non-integer: the only use here)
  const double G[26] = {1.0, 0.5772156649015329, -0.6558780715202538, -0.420026350340952e-1, 0.1665386113822915,
                        -.421977345555443e-1, -.96219715278770e-2, .72189432466630e-2, -.11651675918591e-2,
                        -.2152416741149e-3, .1280502823882e-3, -.201348547807e-4, -.12504934821e-5, .11330272320e-5,
                        -.2056338417e-6, .61160950e-8, .50020075e-8, -.11812746e-8, .1043427e-9, .77823e-11,
                        -.36968e-11, .51e-12, -.206e-13, -.54e-14, .14e-14, .1e-15};
  double z = fabs(x), r = 1.0;
  const int m = (int)z;
  for (int k = 1; k <= m; ++k) r = r * (z - k);
  z = z - m;
  double gr = G[25];
  for (int k = 24; k >= 0; --k) gr = gr * z + G[k];
  return 1.0 / (gr * z) * r;
}

// ---- conventional (prepbufr) observation operator: itpl_2d :1295-1315, phys2ijk :999-1110, prsadj :600-617, Trans_XtoY
// :264-337 of scale/common/common_obs_scale.f90 --------------------------------------------------------------------------------
enum { IQC_PS_TER = 10 };
enum { ID_U = 2819, ID_V = 2820, ID_T = 3073, ID_TV = 3074, ID_Q = 3330, ID_RH = 3331, ID_PS = 14593 };
constexpr double kRv = 461.50, kFvirt = kRv / kRd - 1.0;   // common/common.f90:31, 34

struct Grid2 {   // v2dgh(nlonh, nlath, nv2dd)
  const double *v;
  int nlonh, nlath;
  LETKF_HD double at(int i, int j, int n) const { return v[(size_t)(i - 1) + (size_t)nlonh * ((size_t)(j - 1) + (size_t)nlath * n)]; }
};

LETKF_HD double itpl_2d(const Grid2 &g, int n, double ri, double rj) {
  const int i = (int)ceil(ri), j = (int)ceil(rj);
  const double ai = ri - (double)(i - 1), aj = rj - (double)(j - 1);
  return g.at(i - 1, j - 1, n) * (1 - ai) * (1 - aj) + g.at(i, j - 1, n) * ai * (1 - aj) + g.at(i - 1, j, n) * (1 - ai) * aj +
         g.at(i, j, n) * ai * aj;
}

// one level of itpl_2d_column(LOG(p_full)) (:1055-1056)
LETKF_HD double lnpcol(const Grid &g, int n, int k, int i, int j, double ai, double aj) {
  return log(g.at(k, i - 1, j - 1, n)) * (1 - ai) * (1 - aj) + log(g.at(k, i, j - 1, n)) * ai * (1 - aj) +
         log(g.at(k, i - 1, j, n)) * (1 - ai) * aj + log(g.at(k, i, j, n)) * ai * aj;
}

// phys2ijk: pressure -> fractional level index (surface observations, elem > 9999, keep rlev); returns the QC flag
LETKF_HD int phys2ijk(const Grid &g, int ip, int elem, int nlev, int khalo, double ri, double rj, double rlev, double &rk) {
  rk = kUndef;
  if (ri < 1.0 || ri > (double)g.nlonh || rj < 1.0 || rj > (double)g.nlath) return IQC_OUT_H;
  if (elem > 9999) {
    rk = rlev;
    return IQC_GOOD;
  }
  const int i = (int)ceil(ri), j = (int)ceil(rj);
  int ks = 1 + khalo;   // lowest valid level
  for (int jj = j - 1; jj <= j; ++jj)
    for (int ii = i - 1; ii <= i; ++ii) {
      int k = 1 + khalo;
      for (; k <= nlev + khalo; ++k)
        if (g.at(k, ii, jj, ip) >= 0.0) break;
      if (k > ks) ks = k;
    }
  const double ai = ri - (double)(i - 1), aj = rj - (double)(j - 1);
  const double lr = log(rlev);
  if (lr < lnpcol(g, ip, nlev + khalo, i, j, ai, aj)) return IQC_OUT_VHI;
  if (lr > lnpcol(g, ip, ks, i, j, ai, aj)) return IQC_OUT_VLO;
  int k = ks + 1;
  double pk = 0.0;
  for (; k <= nlev + khalo; ++k) {
    pk = lnpcol(g, ip, k, i, j, ai, aj);
    if (pk < lr) break;   // assuming descending order of plev
  }
  if (k > nlev + khalo) {   // lr == plev(top): the Fortran loop index runs one past the end
    k = nlev + khalo;
    pk = lnpcol(g, ip, k, i, j, ai, aj);
  }
  const double pkm = lnpcol(g, ip, k - 1, i, j, ai, aj);
  rk = (double)(k - 1) + (lr - pkm) / (pk - pkm);
  return IQC_GOOD;
}

LETKF_HD double prsadj(double p, double dz, double t, double q) {
  const double gamma = 5.0e-3;
  if (dz != 0.0) {
    const double tv = t * (1.0 + 0.608 * q);
    p = p * pow((-gamma * dz + tv) / tv, kGG / (gamma * kRd));
  }
  return p;
}

// Trans_XtoY; rot1/rot2 = MPRJ_rotcoef(lon, lat) of the observation (SCALE-RM's map projection: carried as data)
LETKF_HD void trans_xtoy(int elm, double ri, double rj, double rk, double rot1, double rot2, const Grid &g3, const Grid2 &g2,
                         int stggrd, double ps_adjust_thres, double &yobs, int &qc) {
  yobs = kUndef;
  qc = IQC_GOOD;
  if (elm == ID_U || elm == ID_V) {
    double u, v;
    if (stggrd == 1) {
      u = itpl_3d(g3, 0, rk, ri - 0.5, rj);
      v = itpl_3d(g3, 1, rk, ri, rj - 0.5);
    } else {
      u = itpl_3d(g3, 0, rk, ri, rj);
      v = itpl_3d(g3, 1, rk, ri, rj);
    }
    yobs = (elm == ID_U) ? u * rot1 - v * rot2 : u * rot2 + v * rot1;
  } else if (elm == ID_T) {
    yobs = itpl_3d(g3, 3, rk, ri, rj);
  } else if (elm == ID_TV) {
    yobs = itpl_3d(g3, 3, rk, ri, rj);
    const double q = itpl_3d(g3, 5, rk, ri, rj);
    yobs = yobs * (1.0 + kFvirt * q);
  } else if (elm == ID_Q) {
    yobs = itpl_3d(g3, 5, rk, ri, rj);
  } else if (elm == ID_PS) {
    const double t = itpl_2d(g2, 5, ri, rj), q = itpl_2d(g2, 6, ri, rj), topo = itpl_2d(g2, 0, ri, rj);
    yobs = prsadj(itpl_2d(g2, 1, ri, rj), rk - topo, t, q);
    if (fabs(rk - topo) > ps_adjust_thres) qc = IQC_PS_TER;
  } else if (elm == ID_RH) {
    yobs = itpl_3d(g3, 11, rk, ri, rj);
  } else {
    qc = IQC_OTYPE;
  }
}

struct RadarCfg {
  int method, use_tv;
  double min_ref, min_ref_dbz, low_ref_shift;
};

// calc_ref_vr: reflectivity [mm^6/m^3] and radial velocity [m/s]
LETKF_HD void calc_ref_vr(const RadarCfg &c, double qv, double qc, double qr, double qci, double qs, double qg, double u, double v,
                          double w, double t, double p, double az, double elev, double &ref, double &vr) {
  (void)qv; (void)qc; (void)qci;
  double zr = 0.0, zs = 0.0, zg = 0.0, zms = 0.0, zmg = 0.0, wt = 0.0;
  ref = 0.0;
  double ro = p / (kRd * t);
  if (c.method == 1) {
    const double nor = 8.0e6, ror = 1000.0, pip = pow(kPi, 1.75), cf = 10.0e18 * 72, p0 = 1.0e5;
    const double qt = qr + qs + qg;
    if (qt > 0.0) {
      ref = cf * pow(ro * qt, 1.75);
      ref = ref / (pip * pow(nor, 0.75) * pow(ror, 1.75));
      const double a = pow(p0 / p, (double)0.4f);   // default-REAL exponent
      wt = 5.40 * a * pow(qt, 0.125);
    }
  } else if (c.method == 2) {
    double nor = 8.0e6, nos = 3.0e6, nog = 4.0e4, ror = 1000.0, ros = 100.0, rog = 913.0, roi = 917.0, roo = 1.0;
    const double ki2 = 0.176, kr2 = 0.930, pip = pow(kPi, 1.75), cf = 1.0e18 * 720;
    if (qr > 0.0) {
      zr = cf * pow(ro * qr, 1.75);
      zr = zr / (pip * pow(nor, 0.75) * pow(ror, 1.75));
    }
    if (qs > 0.0) {
      if (t <= (double)273.16f) {   // default-REAL literal
        zs = cf * ki2 * pow(ros, 0.25) * pow(ro * qs, 1.75);
        zs = zs / (pip * kr2 * pow(nos, 0.75) * (roi * roi));
      } else {
        zs = cf * pow(ro * qs, 1.75);
        zs = zs / (pip * pow(nos, 0.75) * pow(roi, 1.75));
      }
    }
    if (qg > 0.0) {
      zg = pow(cf / (pip * pow(nog, 0.75) * pow(rog, 1.75)), (double)0.95f);
      zg = zg * pow(ro * qg, (double)1.6625f);
    }
    ref = zr + zs + zg;
    if (ref > 0.0) {
      const double em3 = (double)1e-3f;   // default-REAL 1e-3 literals widened to double
      nor = nor * em3; nos = nos * em3; nog = nog * em3;
      ror = ror * em3; ros = ros * em3; rog = rog * em3; roo = roo * em3;
      ro = ro * em3;
      const double a = 2115.0, b = 0.8, cc = 152.93, d = 0.25, Cd = 0.6;
      const double rofactor = pow(roo / ro, 0.25);
      double wr = 0.0, ws = 0.0, wg = 0.0;
      if (qr > 0.0) {
        const double lr = pow(kPi * ror * nor / (ro * qr), 0.25);
        wr = a * com_gamma(4.0 + b) / (6.0 * pow(lr, b));
        wr = 1.0e-2 * wr * rofactor;
      }
      if (qs > 0.0) {
        const double ls = pow(kPi * ros * nos / (ro * qs), 0.25);
        ws = cc * com_gamma(4.0 + d) / (6.0 * pow(ls, d));
        ws = 1.0e-2 * ws * rofactor;
      }
      if (qg > 0.0) {
        const double lg = pow(kPi * rog * nog / (ro * qg), 0.25);
        wg = com_gamma(4.5) * pow((4.0 * kGG * 100.0 * rog) / (3.0 * Cd * ro), 0.5);
        wg = 1.0e-2 * wg / (6.0 * pow(lg, 0.5));
      }
      wt = (wr * zr + ws * zs + wg * zg) / (zr + zs + zg);
    }
  } else {
    const double MAXF = 0.5;
    double Fg = 0.0, Fs = 0.0, fwg = 0.0, fws = 0.0;
    if (qr > 0.0 && qg > 0.0) {
      Fg = MAXF * pow(fmin(qr / qg, qg / qr), 1.0 / 3.0);
      fwg = qr / (qr + qg);
    }
    if (qr > 0.0 && qs > 0.0) {
      Fs = MAXF * pow(fmin(qr / qs, qs / qr), 1.0 / 3.0);
      fws = qr / (qr + qs);
    }
    const double qrp = (1.0 - Fs - Fg) * qr, qsp = (1.0 - Fs) * qs, qgp = (1.0 - Fg) * qg;
    const double qms = Fs * (qr + qs), qmg = Fg * (qr + qg);
    if (qrp > 0.0) zr = 2.53e4 * pow(ro * qrp * 1.0e3, (double)1.84f);
    if (qsp > 0.0) zs = 3.48e3 * pow(ro * qsp * 1.0e3, (double)1.66f);
    if (qgp > 0.0) zg = 5.54e3 * pow(ro * qgp * 1.0e3, (double)1.70f);
    // default-REAL literals of the Fortran (0.00491, 5.75, 5.588, 1.67, 0.202, 0.398, ...) are single precision
    if (qms > 0.0) {
      zms = ((double)0.00491f + (double)5.75f * fws - (double)5.588f * (fws * fws)) * 1.0e5;
      zms = zms * pow(ro * qms * 1.0e3, (double)1.67f - (double)0.202f * fws + (double)0.398f * (fws * fws));
    }
    if (qmg > 0.0) {
      zmg = ((double)0.809f + (double)10.13f * fwg - (double)5.98f * (fwg * fwg)) * 1.0e5;
      zmg = zmg * pow(ro * qmg * 1.0e3, (double)1.48f + (double)0.0448f * fwg - (double)0.0313f * (fwg * fwg));
    }
    ref = zr + zg + zs + zms + zmg;
    if (ref > 0.0) {
      const double nor = 8.0e-2, nos = 3.0e-2, nog = 4.0e-4, ror = 1.0, ros = 0.1, rog = 0.917, roo = 0.001;
      ro = 1.0e-3 * ro;
      const double a = 2115.0, b = 0.8, cc = 152.93, d = 0.25, Cd = 0.6;
      const double rofactor = pow(roo / ro, 0.5);
      double wr = 0.0, ws = 0.0, wg = 0.0;
      if (qr > 0.0) {
        const double lr = pow(kPi * ror * nor / (ro * qr), 0.25);
        wr = a * com_gamma(4.0 + b) / (6.0 * pow(lr, b));
        wr = 1.0e-2 * wr * rofactor;
      }
      if (qs > 0.0) {
        const double ls = pow(kPi * ros * nos / (ro * qs), 0.25);
        ws = cc * com_gamma(4.0 + d) / (6.0 * pow(ls, d));
        ws = 1.0e-2 * ws * rofactor;
      }
      if (qg > 0.0) {
        const double lg = pow(kPi * rog * nog / (ro * qg), 0.25);
        wg = com_gamma(4.5) * pow((4.0 * kGG * 100.0 * rog) / (3.0 * Cd * ro), 0.5);
        wg = 1.0e-2 * wg / (6.0 * pow(lg, 0.5));
      }
      wt = (wr * zr + ws * zs + ws * zms + wg * zg + wg * zmg) / (zr + zs + zg + zms + zmg);
    }
  }
  vr = u * cos(elev * kDeg2Rad) * sin(az * kDeg2Rad);
  vr = vr + v * cos(elev * kDeg2Rad) * cos(az * kDeg2Rad);
  if (c.use_tv) vr = vr + (w - wt) * sin(elev * kDeg2Rad);
  else vr = vr + w * sin(elev * kDeg2Rad);
}

// Trans_XtoY_radar for one (observation, member); rk from phys2ijkz
LETKF_HD void trans_xtoy_radar(const RadarCfg &c, const Grid &g, int elm, double radar_lon, double radar_lat, double radar_z,
                               double ri, double rj, double rk, double lon, double lat, double lev, double rot1, double rot2,
                               double &yobs, int &qc) {
  yobs = kUndef;
  qc = IQC_GOOD;
  double ur = itpl_3d(g, 0, rk, ri, rj), vr = itpl_3d(g, 1, rk, ri, rj);
  const double wr = itpl_3d(g, 2, rk, ri, rj), tr = itpl_3d(g, 3, rk, ri, rj), pr = itpl_3d(g, 4, rk, ri, rj);
  const double qvr = itpl_3d(g, 5, rk, ri, rj), qcr = itpl_3d(g, 6, rk, ri, rj), qrr = itpl_3d(g, 7, rk, ri, rj);
  const double qir = itpl_3d(g, 8, rk, ri, rj), qsr = itpl_3d(g, 9, rk, ri, rj), qgr = itpl_3d(g, 10, rk, ri, rj);
  const double utmp = ur, vtmp = vr;
  ur = utmp * rot1 - vtmp * rot2;
  vr = utmp * rot2 + vtmp * rot1;
  const double dlon = lon - radar_lon, dlat = lat - radar_lat;
  if (dlon == 0.0 && dlat == 0.0) {
    qc = IQC_OUT_H;
    return;
  }
  double az = kRad2Deg * atan2(dlon * cos(radar_lat * kDeg2Rad), dlat);
  if (az < 0) az = 360.0 + az;
  double dist;
  {   // com_distll_1
    const double r180 = 1.0 / 180.0;
    const double lon1 = lon * kPi * r180, lon2 = radar_lon * kPi * r180, lat1 = lat * kPi * r180, lat2 = radar_lat * kPi * r180;
    double cosd = sin(lat1) * sin(lat2) + cos(lat1) * cos(lat2) * cos(lon2 - lon1);
    cosd = fmin(1.0, cosd);
    cosd = fmax(-1.0, cosd);
    dist = acos(cosd) * kRe;
  }
  const double elev = kRad2Deg * atan2(lev - radar_z, dist);
  double radar_ref, radar_rv;
  calc_ref_vr(c, qvr, qcr, qrr, qir, qsr, qgr, ur, vr, wr, tr, pr, az, elev, radar_ref, radar_rv);
  if (elm == ID_REF || elm == ID_RE0) {
    if (radar_ref < c.min_ref) {
      qc = IQC_REF_LOW;
      yobs = c.min_ref_dbz + c.low_ref_shift;
    } else {
      yobs = 10.0 * log10(radar_ref);
    }
  } else if (elm == ID_VR) {
    if (radar_ref < c.min_ref) qc = IQC_REF_LOW;
    yobs = radar_rv;
  } else {
    qc = IQC_OTYPE;
  }
}

}  // namespace letkf_radar
