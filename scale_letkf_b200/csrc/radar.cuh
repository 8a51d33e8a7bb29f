// radar.cuh -- radar observation operator for all members on the device: the obsfmt_radar
// branch of obsope_cal (scale/obs/obsope_tools.f90:476-494).  One thread per (observation, member), observations
// along the lanes: neighbouring observations of a volume scan touch neighbouring grid cells of the same member, so
// the 4-column height search and the 88 corner values of the tri-linear interpolations come out of L1/L2.
// HBM-bound: per (obs, member) 11 variables x 8 corners + the height column.
#pragma once
#include "common.cuh"
#include "radar_math.h"

namespace letkf {

struct RadarParams {
  letkf_radar::RadarCfg c;
  int nobs, nmem, nlevh, nlonh, nlath, nlev, khalo, nv3dd, ld_out;
  double zmax, radar_lon, radar_lat, radar_z;
  const int *elm;
  const double *ril, *rjl, *lon, *lat, *lev, *rotc;
  const double *const *v3dgh;   // [nmem] device pointers
  double *yobs;                 // [nobs][ld_out]
  int *qc;
};

__global__ void __launch_bounds__(128) obsope_radar_kernel(const RadarParams P) {
  using namespace letkf_radar;
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= P.nobs) return;
  Grid g{P.v3dgh[m], P.nlevh, P.nlonh, P.nlath};
  const double lev = P.lev[n];
  double y = kUndef;
  int qc;
  if (lev > P.zmax) {
    qc = IQC_RADAR_VHI;
  } else {
    double rk;
    qc = phys2ijkz(g, P.nv3dd - 1, P.nlev, P.khalo, P.ril[n], P.rjl[n], lev, rk);
    if (qc == IQC_GOOD) {
      const double r1 = P.rotc ? P.rotc[n] : 1.0, r2 = P.rotc ? P.rotc[(size_t)P.nobs + n] : 0.0;
      trans_xtoy_radar(P.c, g, P.elm[n], P.radar_lon, P.radar_lat, P.radar_z, P.ril[n], P.rjl[n], rk, P.lon[n], P.lat[n], lev,
                       r1, r2, y, qc);
      if (qc == IQC_REF_LOW) qc = IQC_GOOD;   // obsope_tools.f90:489
    }
  }
  P.yobs[(size_t)n * P.ld_out + m] = y;
  P.qc[(size_t)n * P.ld_out + m] = qc;
}

// ---- conventional (prepbufr) branch of obsope_cal (obsope_tools.f90:466-473) / monit_obs (common_obs_scale.f90:1530-1540):
// phys2ijk + Trans_XtoY for every (observation, member) ---------------------------------------------------------------------
struct ConvParams {
  int nobs, nmem, nlevh, nlonh, nlath, nlev, khalo, ld_out, stggrd;
  double ps_thres;
  const int *elm;
  const double *ril, *rjl, *lev, *rotc;
  const double *const *v3dgh, *const *v2dgh;   // [nmem] device pointers
  double *yobs;
  int *qc;
};

__global__ void __launch_bounds__(128) obsope_conv_kernel(const ConvParams P) {
  using namespace letkf_radar;
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= P.nobs) return;
  Grid g3{P.v3dgh[m], P.nlevh, P.nlonh, P.nlath};
  Grid2 g2{P.v2dgh[m], P.nlonh, P.nlath};
  double y = kUndef, rk;
  int qc = phys2ijk(g3, 4, P.elm[n], P.nlev, P.khalo, P.ril[n], P.rjl[n], P.lev[n], rk);
  if (qc == IQC_GOOD) {
    const double r1 = P.rotc ? P.rotc[n] : 1.0, r2 = P.rotc ? P.rotc[(size_t)P.nobs + n] : 0.0;
    trans_xtoy(P.elm[n], P.ril[n], P.rjl[n], rk, r1, r2, g3, g2, P.stggrd, P.ps_thres, y, qc);
  }
  P.yobs[(size_t)n * P.ld_out + m] = y;
  P.qc[(size_t)n * P.ld_out + m] = qc;
}

// monit_obs (common_obs_scale.f90:1516-1572): observations outside DEPARTURE_STAT_T_RANGE keep oqc = -1; accepted ones get
// ohx = dat - H(x), the others undef
__global__ void monit_ohx_kernel(int nobs, const double *__restrict__ dat, const double *__restrict__ dif, double t_range,
                                 const double *__restrict__ hx, const int *__restrict__ qc, double *__restrict__ ohx,
                                 int *__restrict__ oqc) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nobs) return;
  if (t_range > 0.0 && dif && !(fabs(dif[n]) <= t_range)) {
    oqc[n] = -1;
    ohx[n] = letkf_radar::kUndef;
    return;
  }
  const int q = qc[n];
  oqc[n] = q;
  ohx[n] = (q == letkf_radar::IQC_GOOD) ? __dsub_rn(dat[n], hx[n]) : letkf_radar::kUndef;
}

}  // namespace letkf
