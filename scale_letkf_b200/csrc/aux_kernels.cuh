// aux_kernels.cuh -- HBM-bound helper kernels: device-side bucket sort of observations
// (twin of set_letkf_obs' counting sort, scale/letkf/letkf_obs.f90:747-805, :922-976),
// ensmean_grd (scale/common/common_scale.f90:1513) and the pack/unpack halves of the
// member<->grid transposes (scale/common/common_mpi_scale.f90:1279-1476).
#pragma once
#include "search.cuh"

namespace letkf {

// ---- bucket sort ----------------------------------------------------------------------------
// key[n] = global bucket id of observation n: boff(ic) + (j-1)*ngrdext_i + (i-1), with (i, j)
// from ij_obsgrd (letkf_obs.f90:1187-1204) clamped to the subdomain (:758-761) and shifted
// by ngrdsch into the extended mesh (:936-940).
__global__ void bucket_key_kernel(const SearchTables *T, int nobs, const int *__restrict__ ic_of,
                                  const double *__restrict__ ri, const double *__restrict__ rj,
                                  int *__restrict__ key, int *__restrict__ count) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nobs) return;
  const CtypeDev &c = T->ct[ic_of[n]];
  int i = obsgrd_index(ri[n], T->IHALO, c.ngrd_i, T->nlon, 0);
  int j = obsgrd_index(rj[n], T->JHALO, c.ngrd_j, T->nlat, 0);
  i = min(max(i, 1), c.ngrd_i) + c.ngrdsch_i;
  j = min(max(j, 1), c.ngrd_j) + c.ngrdsch_j;
  const int b = c.boff + (j - 1) * c.ngrdext_i + (i - 1);
  key[n] = b;
  atomicAdd(&count[b], 1);
}

// ---- device-resident front half of set_letkf_obs (letkf_obs.f90:300-342, 752, 791): which (element, type) pairs occur
// among the accepted observations, their combined type, vertical coordinate and the compaction of qc == iqc_good ------
__device__ __forceinline__ int dev_uid_obs(int elm) {   // common_obs_scale.f90:171-211 (0 = unknown)
  switch (elm) {
    case 2819: return 1; case 2820: return 2; case 3073: return 3; case 3074: return 4; case 3330: return 5; case 3331: return 6;
    case 14593: return 7; case 19999: return 8; case 4001: return 9; case 4004: return 10; case 4002: return 11;
    case 4003: return 12; case 8800: return 13; case 99991: return 14; case 99992: return 15; case 99993: return 16;
    default: return 0;
  }
}
// use[(ielm_u-1) * NOBTYPE + ityp-1] |= 1 for every accepted observation; err[0] |= 1 on an unknown element / type
__global__ void obs_use_kernel(int nobs, int nobtype, const int *__restrict__ elm, const int *__restrict__ typ,
                               const int *__restrict__ qc, int *__restrict__ use, int *__restrict__ err) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nobs || (qc && qc[n] != 0)) return;
  const int u = dev_uid_obs(elm[n]), t = typ[n];
  if (u < 1 || t < 1 || t > nobtype) {
    atomicOr(err, 1);
    return;
  }
  use[(u - 1) * nobtype + t - 1] = 1;
}
// keep[n] = accepted; ic[n] = combined type; vc[n] = vertical coordinate (ln p, ln ps, or height for vmode 3)
__global__ void obs_prepare_kernel(int nobs, int nobtype, const int *__restrict__ elm, const int *__restrict__ typ,
                                   const int *__restrict__ qc, const double *__restrict__ lev, const double *__restrict__ dat,
                                   const int *__restrict__ ctype_elmtyp, const int *__restrict__ vmode, int *__restrict__ keep,
                                   int *__restrict__ ic, double *__restrict__ vc, int *__restrict__ err) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nobs) return;
  const bool k = !(qc && qc[n] != 0);
  keep[n] = k ? 1 : 0;
  if (!k) return;
  const int c = ctype_elmtyp[(dev_uid_obs(elm[n]) - 1) * nobtype + typ[n] - 1] - 1;
  ic[n] = c;
  const int vm = vmode[c];
  const double pr = (elm[n] == 14593) ? dat[n] : lev[n];
  if (vm == 1 && !(pr > 0.0)) atomicOr(err, 2);   // ln of a non-positive pressure
  vc[n] = (vm == 3) ? lev[n] : log(pr);
}
// stream compaction in arrival order (pos = exclusive scan of keep)
__global__ void obs_compact_kernel(int nobs, int nensobs, const int *__restrict__ keep, const int *__restrict__ pos,
                                   const int *__restrict__ ic, const double *__restrict__ vc, const double *__restrict__ ri,
                                   const double *__restrict__ rj, const double *__restrict__ err, const double *__restrict__ val,
                                   const double *__restrict__ ens, int *__restrict__ o_ic, double *__restrict__ o_vc,
                                   double *__restrict__ o_ri, double *__restrict__ o_rj, double *__restrict__ o_err,
                                   double *__restrict__ o_val, double *__restrict__ o_ens, int *__restrict__ kept_index) {
  const int n = blockIdx.x, lane = threadIdx.x;
  if (n >= nobs || !keep[n]) return;
  const int d = pos[n];
  if (lane == 0) {
    o_ic[d] = ic[n]; o_vc[d] = vc[n]; o_ri[d] = ri[n]; o_rj[d] = rj[n]; o_err[d] = err[n]; o_val[d] = val[n];
    kept_index[d] = n;
  }
  for (int m = lane; m < nensobs; m += blockDim.x) o_ens[(size_t)d * nensobs + m] = ens[(size_t)n * nensobs + m];
}

// single-CTA exclusive scan (setup path; nb up to a few million buckets)
__global__ void exclusive_scan_kernel(const int *__restrict__ count, int *__restrict__ start, int nb) {
  __shared__ int red[kMaxWarps];
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < nb ? count[i] : 0;
    int tot;
    const int ex = block_excl_scan_i(v, red, tot);
    const int carry = carry_s;
    if (i < nb) start[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[nb] = carry_s;
}

__global__ void bucket_scatter_kernel(int nobs, const int *__restrict__ key, const int *__restrict__ start,
                                      int *__restrict__ fill, int *__restrict__ tmp) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nobs) return;
  const int b = key[n];
  tmp[start[b] + atomicAdd(&fill[b], 1)] = n;
}

// restore arrival order inside each bucket (stable counting sort): rank = #entries of the
// same bucket with a smaller original index.
__global__ void bucket_rank_kernel(int nobs, const int *__restrict__ key, const int *__restrict__ start,
                                   const int *__restrict__ tmp, int *__restrict__ sorted_to_orig) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nobs) return;
  const int me = tmp[s];
  const int b = key[me];
  const int b0 = start[b], b1 = start[b + 1];
  int rank = 0;
  for (int e = b0; e < b1; ++e) rank += (tmp[e] < me);
  sorted_to_orig[b0 + rank] = me;
}

// Sorted observation table: rec[s], sval[s] and the row sens[s][0..ldens) =
//   [ ensval(1..k) | val (= y - mean H(x), "dep") | ensval(k+1) (= y - H(x_det), "depd", DET_RUN) | 0 .. ]
// so that one contiguous row copy brings everything the Gram + right-hand sides need.
__global__ void obs_gather_kernel(int nobs, int nensobs, int k, int ldens, const int *__restrict__ s2o,
                                  const double *__restrict__ ri, const double *__restrict__ rj,
                                  const double *__restrict__ vc, const double *__restrict__ err,
                                  const double *__restrict__ val, const double *__restrict__ ensval,
                                  ObsRec *__restrict__ rec, double *__restrict__ sval,
                                  double *__restrict__ sens) {
  const int s = blockIdx.x;
  if (s >= nobs) return;
  const int n = s2o[s];
  if (threadIdx.x == 0) {
    ObsRec r;
    r.ri = ri[n];
    r.rj = rj[n];
    r.vc = vc[n];
    r.err = err[n];
    rec[s] = r;
    sval[s] = val[n];
  }
  for (int m = threadIdx.x; m < ldens; m += blockDim.x) {
    double v = 0.0;
    if (m < k) v = ensval[(size_t)n * nensobs + m];
    else if (m == k) v = val[n];
    else if (m == k + 1 && nensobs > k) v = ensval[(size_t)n * nensobs + k];
    sens[(size_t)s * ldens + m] = v;
  }
}

__global__ void fill_kernel(double *__restrict__ p, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// work3d = max(work3d, INFL_MUL_MIN) over the whole field (letkf_tools.f90:264-267)
__global__ void clamp_min_kernel(double *__restrict__ p, size_t n, double lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = fmax(p[i], lo);
}

// ---- ensmean_grd ----------------------------------------------------------------------------
// slot mem+1 = (x_1 + x_2 + ... + x_mem) / mem, summed in member order like the reference.
__global__ void ensmean_kernel(int mem, int nens, size_t sl, int nvar, double *__restrict__ v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl * nvar) return;
  const size_t n = i / sl, p = i - n * sl;
  double *b = v + p + n * (size_t)nens * sl;
  double a = b[0];
  for (int m = 1; m < mem; ++m) a += b[(size_t)m * sl];
  b[(size_t)mem * sl] = a / (double)mem;
}

// ---- enssprd_grd (common_scale.f90:1557-1611) -------------------------------------------------
// sqrt(sum_m (x_m - mean)^2 / (mem - 1)), member order, no FMA contraction (bit-identical to the CPU).
__global__ void enssprd_kernel(int mem, int nens, size_t sl, int nvar, const double *__restrict__ v,
                               double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl * nvar) return;
  const size_t n = i / sl, p = i - n * sl;
  const double *b = v + p + n * (size_t)nens * sl;
  const double mean = b[(size_t)mem * sl];
  double d = __dsub_rn(b[0], mean);
  double a = __dmul_rn(d, d);
  for (int m = 1; m < mem; ++m) {
    d = __dsub_rn(b[(size_t)m * sl], mean);
    a = __dadd_rn(a, __dmul_rn(d, d));
  }
  out[i] = __dsqrt_rn(__ddiv_rn(a, (double)(mem - 1)));
}

// ---- additive inflation block of das_letkf (letkf_tools.f90:804-929) ------------------------------
// addinfl_weight(ij) of INFL_ADD_REF_ONLY (:816-840): exp(-d^2 / 2) of the nearest radar-reflectivity observation in units
// of its horizontal localisation scale, 0 beyond dist_zero_fac.  Brute force over all observations [b0, b1) of the
// combined type like the reference; observation positions are staged through shared memory 256 at a time.
__global__ void __launch_bounds__(256) addinfl_weight_kernel(int nij, const double *__restrict__ rig1, const double *__restrict__ rjg1,
                                                            const ObsRec *__restrict__ rec, int b0, int b1, double DX, double DY,
                                                            double hloc, double dzf2, double *__restrict__ w) {
  __shared__ double sri[256], srj[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  const double ri = i < nij ? rig1[i] : 0.0, rj = i < nij ? rjg1[i] : 0.0;
  double best = 1.0e33;
  for (int base = b0; base < b1; base += 256) {
    const int n = min(256, b1 - base);
    __syncthreads();
    if ((int)threadIdx.x < n) {
      sri[threadIdx.x] = rec[base + threadIdx.x].ri;
      srj[threadIdx.x] = rec[base + threadIdx.x].rj;
    }
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      const double rdx = __dmul_rn(__dsub_rn(ri, sri[j]), DX), rdy = __dmul_rn(__dsub_rn(rj, srj[j]), DY);
      const double d = __dadd_rn(__dmul_rn(rdx, rdx), __dmul_rn(rdy, rdy));
      if (d < best) best = d;
    }
  }
  if (i >= nij) return;
  best = __ddiv_rn(best, __dmul_rn(hloc, hloc));
  w[i] = (best <= dzf2) ? exp(-0.5 * best) : 0.0;
}
// anal(p, m, n) += (addi(p, ishuf(m), n) - mean_m addi(p, :, n)) * INFL_ADD * weight(i) [* gues mean(p, n) for the moisture
// variables with INFL_ADD_Q_RATIO] (:869-925): the ensemble mean of the additive ensemble is summed in member order like
// ensmean_grd, the products are taken left to right without FMA contraction -> bit-identical to the CPU.  One thread
// per (point, variable); HBM-bound: the additive members are read twice (mean, update), the analysis once each way.
__global__ void additive_inflation_kernel(int mem, int nens, int nij, size_t sl, int nvar, const double *__restrict__ addi,
                                          double *__restrict__ anal, const double *__restrict__ gues_mean, size_t gm_vstride,
                                          size_t gm_off, const double *__restrict__ w, const int *__restrict__ ishuf,
                                          double infl_add, int q_lo, int q_hi) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl * nvar) return;
  const size_t n = i / sl, p = i - n * sl;
  const double *b = addi + p + n * (size_t)nens * sl;
  double *a = anal + p + n * (size_t)nens * sl;
  double s = b[0];
  for (int m = 1; m < mem; ++m) s = __dadd_rn(s, b[(size_t)m * sl]);
  const double mean = __ddiv_rn(s, (double)mem);
  const double wi = w ? w[p % (size_t)nij] : 1.0;
  const bool q = gues_mean != nullptr && (int)n >= q_lo && (int)n <= q_hi;
  const double qf = q ? gues_mean[p + n * gm_vstride + gm_off] : 1.0;   // background mean: slot `mem` of gues3d, or staged planes
  for (int m = 0; m < mem; ++m) {
    const int ms = ishuf ? ishuf[m] - 1 : m;
    double t = __dmul_rn(__dmul_rn(__dsub_rn(b[(size_t)ms * sl], mean), infl_add), wi);
    if (q) t = __dmul_rn(t, qf);
    a[(size_t)m * sl] = __dadd_rn(a[(size_t)m * sl], t);
  }
}

// The same update with the additive members of the column held in registers (MEMBER <= KR): every value of the additive
// ensemble is read ONCE (the two-pass kernel above re-reads the members for the update and the second read misses L2 on
// large states: 1.3x the algorithmic DRAM traffic, ncu).  `inv` = inverse of the shuffle (destination member of source
// member ms; null = identity) so that the register index stays static.  Same operations in the same order: bit-identical.
template <int KR>
__global__ void __launch_bounds__(128) additive_inflation_reg_kernel(int mem, int nens, int nij, size_t sl, int nvar,
                                                                     const double *__restrict__ addi, double *__restrict__ anal,
                                                                     const double *__restrict__ gues_mean, size_t gm_vstride, size_t gm_off,
                                                                     const double *__restrict__ w, const int *__restrict__ inv,
                                                                     double infl_add, int q_lo, int q_hi) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sl * nvar) return;
  const size_t n = i / sl, p = i - n * sl;
  const double *b = addi + p + n * (size_t)nens * sl;
  double *a = anal + p + n * (size_t)nens * sl;
  double r[KR];
#pragma unroll
  for (int m = 0; m < KR; ++m) r[m] = (m < mem) ? b[(size_t)m * sl] : 0.0;
  double s = r[0];
#pragma unroll
  for (int m = 1; m < KR; ++m)
    if (m < mem) s = __dadd_rn(s, r[m]);
  const double mean = __ddiv_rn(s, (double)mem);
  const double wi = w ? w[p % (size_t)nij] : 1.0;
  const bool q = gues_mean != nullptr && (int)n >= q_lo && (int)n <= q_hi;
  const double qf = q ? gues_mean[p + n * gm_vstride + gm_off] : 1.0;
  // the analysis members in groups of eight: the eight loads are issued together (the compiler cannot move a load of
  // a[m'] above the store to a[m]: it does not know that the member planes are disjoint)
#pragma unroll
  for (int g = 0; g < KR; g += 8) {
    double av[8];
    size_t off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ms = g + j;
      off[j] = (size_t)((ms < mem) ? (inv ? inv[ms] : ms) : 0) * sl;
      av[j] = (ms < mem) ? a[off[j]] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ms = g + j;
      if (ms < mem) {
        double t = __dmul_rn(__dmul_rn(__dsub_rn(r[ms], mean), infl_add), wi);
        if (q) t = __dmul_rn(t, qf);
        a[off[j]] = __dadd_rn(av[j], t);
      }
    }
  }
}

// ---- transposes -----------------------------------------------------------------------------
struct TransposeDims {
  int nlon, nlat, nlev, nv3d, nv2d, np, nij1max, nlevall;
};
__device__ __forceinline__ int nij1_of(const TransposeDims &d, int rank) {
  const int r = (d.nlon * d.nlat) % d.np;
  return rank < r ? d.nij1max : d.nij1max - 1;
}
// grd_to_buf over all levels/variables of one member: bufs(nij1max, nlevall, np)
__global__ void grd_to_buf_kernel(TransposeDims d, const double *__restrict__ v3dg,
                                  const double *__restrict__ v2dg, double *__restrict__ bufs) {
  const size_t total = (size_t)d.nij1max * d.nlevall * d.np;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % d.nij1max);
    const size_t t = idx / d.nij1max;
    const int jl = (int)(t % d.nlevall), m = (int)(t / d.nlevall);
    double v = -9.99e33;   // undef (common/common.f90:38) in the padding row
    if (i < nij1_of(d, m)) {
      const int j = m + d.np * i;
      const int ilon = j % d.nlon, ilat = j / d.nlon;
      if (jl < d.nlev * d.nv3d) {
        const int n = jl / d.nlev, k = jl - n * d.nlev;
        v = v3dg[k + (size_t)d.nlev * (ilon + (size_t)d.nlon * (ilat + (size_t)d.nlat * n))];
      } else {
        const int n = jl - d.nlev * d.nv3d;
        v = v2dg[ilon + (size_t)d.nlon * (ilat + (size_t)d.nlat * n)];
      }
    }
    bufs[idx] = v;
  }
}
__global__ void buf_to_grd_kernel(TransposeDims d, const double *__restrict__ bufr,
                                  double *__restrict__ v3dg, double *__restrict__ v2dg) {
  const size_t total = (size_t)d.nij1max * d.nlevall * d.np;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % d.nij1max);
    const size_t t = idx / d.nij1max;
    const int jl = (int)(t % d.nlevall), m = (int)(t / d.nlevall);
    if (i >= nij1_of(d, m)) continue;
    const int j = m + d.np * i;
    const int ilon = j % d.nlon, ilat = j / d.nlon;
    if (jl < d.nlev * d.nv3d) {
      const int n = jl / d.nlev, k = jl - n * d.nlev;
      v3dg[k + (size_t)d.nlev * (ilon + (size_t)d.nlon * (ilat + (size_t)d.nlat * n))] = bufr[idx];
    } else {
      const int n = jl - d.nlev * d.nv3d;
      v2dg[ilon + (size_t)d.nlon * (ilat + (size_t)d.nlat * n)] = bufr[idx];
    }
  }
}
// bufr(nij1max, nlevall, mcount) -> v3d(nij1, nlev, nens, nv3d), v2d(nij1, nens, nv2d) slots mstart..mend
__global__ void buf_to_ens_kernel(TransposeDims d, int nij1, int nens, int mstart, int mcount,
                                  const double *__restrict__ bufr, double *__restrict__ v3d,
                                  double *__restrict__ v2d, int reverse) {
  const size_t total = (size_t)nij1 * d.nlevall * mcount;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % nij1);
    const size_t t = idx / nij1;
    const int jl = (int)(t % d.nlevall), mm = (int)(t / d.nlevall);
    const size_t bi = i + (size_t)d.nij1max * (jl + (size_t)d.nlevall * mm);
    const int m = mstart - 1 + mm;
    double *p;
    if (jl < d.nlev * d.nv3d) {
      const int n = jl / d.nlev, k = jl - n * d.nlev;
      p = v3d + i + (size_t)nij1 * (k + (size_t)d.nlev * (m + (size_t)nens * n));
    } else {
      const int n = jl - d.nlev * d.nv3d;
      p = v2d + i + (size_t)nij1 * (m + (size_t)nens * n);
    }
    if (reverse) const_cast<double *>(bufr)[bi] = *p; else *p = bufr[bi];
  }
}

// ---------------------------------------------------------------------------------------------
// Departure + QC half of set_letkf_obs (scale/letkf/letkf_obs.f90:355-560), one warp per observation.
// Members are read 32 at a time (coalesced) and summed by lane 0 in member order, so the mean has the
// reference's sequential rounding.
struct QcParams {
  double ge, ge_rain, ge_ref, ge_vr, ge_prh, ge_tcx, ge_tcy, ge_tcp, ref_thres;
  int use_ref, use_vr, min_mem, min_mem_obsref;
  int nobs, nensobs, member, det;
  const int *elm;
  const double *dat, *err;
  int *qc;
  double *ensval, *val;
};
__global__ void __launch_bounds__(256) obs_departure_qc_kernel(const QcParams P) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= P.nobs) return;
  if (P.qc[n] > 0) return;
  const int elm = P.elm[n], k = P.member;
  const double dat = P.dat[n];
  double *ev = P.ensval + (size_t)n * P.nensobs;
  if (elm == 4001 || elm == 4004) {   // id_radar_ref_obs, id_radar_ref_zero_obs (:370-414)
    if (!P.use_ref) { if (lane == 0) P.qc[n] = 90; return; }
    if (dat == -9.99e33) { if (lane == 0) P.qc[n] = 50; return; }
    int mem_ref = 0;
    for (int i0 = 0; i0 < k; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < k && ev[i] > P.ref_thres + 1.0e-6;
      mem_ref += __popc(__ballot_sync(LETKF_FULL_MASK, hit));
    }
    const int need = (dat > P.ref_thres + 1.0e-6) ? P.min_mem_obsref : P.min_mem;
    if (mem_ref < need) { if (lane == 0) P.qc[n] = 12; return; }
  }
  if (elm == 4002 && !P.use_vr) { if (lane == 0) P.qc[n] = 90; return; }   // id_radar_vr_obs (:416-421)
  // mean of H(x): val = ensval(1); val += ensval(i), i = 2..MEMBER; val /= MEMBER   (:474-478)
  double acc = 0.0;
  for (int i0 = 0; i0 < k; i0 += 32) {
    const int i = i0 + lane;
    const double v = i < k ? ev[i] : 0.0;
    const int cnt = min(32, k - i0);
    for (int j = 0; j < cnt; ++j) {
      const double vj = __shfl_sync(LETKF_FULL_MASK, v, j);
      acc = (i0 + j == 0) ? vj : __dadd_rn(acc, vj);
    }
  }
  const double mean = __ddiv_rn(acc, (double)k);
  for (int i = lane; i < k; i += 32) ev[i] = __dsub_rn(ev[i], mean);   // Hdx (:486-488)
  const double dep = __dsub_rn(dat, mean);                              // y - Hx (:489)
  if (lane == 0) {
    P.val[n] = dep;
    if (P.det) ev[k] = __dsub_rn(dat, ev[k]);                           // (:490-492)
    double ge;
    switch (elm) {   // gross error (:503-549)
      case 19999: ge = P.ge_rain; break;
      case 4001: case 4004: ge = P.ge_ref; break;
      case 4002: ge = P.ge_vr; break;
      case 4003: ge = P.ge_prh; break;
      case 99991: ge = P.ge_tcx; break;
      case 99992: ge = P.ge_tcy; break;
      case 99993: ge = P.ge_tcp; break;
      default: ge = P.ge; break;
    }
    if (fabs(dep) > __dmul_rn(ge, P.err[n])) P.qc[n] = 5;
  }
}

// ---------------------------------------------------------------------------------------------
// state_trans / state_trans_inv (scale/common/common_scale.f90:1181-1280), one thread per grid point of a
// member-major grid v3dg(nlev,nlon,nlat,nv3d): variable stride = npts, consecutive threads = consecutive
// points (coalesced); HBM-bound, (nv3d + 5) * 8 bytes per point.
struct StateTransParams {
  double Rdry, Rvap, CVdry, PRE00;
  double tracer_cv[16];
  int pos_q, pos_qhyd;
  int nv3d, iv3d_q;     // iv3d_q 1-based
  long long npts;
  double *v;
};
__global__ void __launch_bounds__(256) state_trans_kernel(const StateTransParams P, int inverse) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.npts) return;
  double *v = P.v + p;
  const long long st = P.npts;
  const int iq = P.iv3d_q - 1;
  if (inverse) {   // (:1243-1250) positive-definite clamps first
    for (int n = iq; n < P.nv3d; ++n) {
      const bool clamp = (n == iq) ? P.pos_q : (P.pos_qhyd && n <= iq + 5);
      if (clamp) v[n * st] = fmax(v[n * st], 0.0);
    }
  }
  double qdry = 1.0, cvtot = 0.0;
  for (int n = iq; n < P.nv3d; ++n) {   // (:1199-1203 / :1257-1261) no FMA contraction: same rounding as the CPU
    const double q = v[n * st];
    qdry = __dsub_rn(qdry, q);
    cvtot = __dadd_rn(cvtot, __dmul_rn(q, P.tracer_cv[n - iq]));
  }
  cvtot = __dadd_rn(__dmul_rn(P.CVdry, qdry), cvtot);
  const double rtot = __dadd_rn(__dmul_rn(P.Rdry, qdry), __dmul_rn(P.Rvap, v[iq * st]));
  if (!inverse) {
    const double cpovcv = __ddiv_rn(__dadd_rn(cvtot, rtot), cvtot);
    const double rho = v[0];
    const double pres = __dmul_rn(P.PRE00, pow(__ddiv_rn(__dmul_rn(v[4 * st], rtot), P.PRE00), cpovcv));
    const double temp = __ddiv_rn(pres, __dmul_rn(rho, rtot));
    v[0] = __ddiv_rn(v[1 * st], rho);
    v[1 * st] = __ddiv_rn(v[2 * st], rho);
    v[2 * st] = __ddiv_rn(v[3 * st], rho);
    v[3 * st] = temp;
    v[4 * st] = pres;
  } else {
    const double cvovcp = __ddiv_rn(cvtot, __dadd_rn(cvtot, rtot));
    const double pres = v[4 * st];
    const double rho = __ddiv_rn(pres, __dmul_rn(rtot, v[3 * st]));
    const double rhot = __dmul_rn(__ddiv_rn(P.PRE00, rtot), pow(__ddiv_rn(pres, P.PRE00), cvovcp));
    v[4 * st] = rhot;
    v[3 * st] = __dmul_rn(v[2 * st], rho);
    v[2 * st] = __dmul_rn(v[1 * st], rho);
    v[1 * st] = __dmul_rn(v[0], rho);
    v[0] = rho;
  }
}

// ---------------------------------------------------------------------------------------------
// Shared-memory tiled pack / unpack of one member-major grid with the state transform fused in
// (SURVEY.md section 8f rank 2): grd_to_buf + state_trans, buf_to_grd + state_trans_inv.
// A CTA owns a tile of 32 columns (of the cyclic deal of rank m) x 32 levels for ALL nv3d variables:
// reads run along the level index (contiguous in v3dg), writes along the column index (contiguous in the
// buffer), the transform works on the tile in shared memory.  grid (ceil(nij1max/32), ceil(nlev/32), np),
// block (32, 8), dynamic shared memory nv3d * 32 * 33 doubles.
__device__ __forceinline__ void state_trans_point(const StateTransParams &P, double *v, int st, int inverse) {
  // v[n * st]: variable n of one grid point (st = stride between variables); same arithmetic as state_trans_kernel
  const int iq = P.iv3d_q - 1;
  if (inverse) {
    for (int n = iq; n < P.nv3d; ++n) {
      const bool clamp = (n == iq) ? P.pos_q : (P.pos_qhyd && n <= iq + 5);
      if (clamp) v[n * st] = fmax(v[n * st], 0.0);
    }
  }
  double qdry = 1.0, cvtot = 0.0;
  for (int n = iq; n < P.nv3d; ++n) {
    const double q = v[n * st];
    qdry = __dsub_rn(qdry, q);
    cvtot = __dadd_rn(cvtot, __dmul_rn(q, P.tracer_cv[n - iq]));
  }
  cvtot = __dadd_rn(__dmul_rn(P.CVdry, qdry), cvtot);
  const double rtot = __dadd_rn(__dmul_rn(P.Rdry, qdry), __dmul_rn(P.Rvap, v[iq * st]));
  if (!inverse) {
    const double cpovcv = __ddiv_rn(__dadd_rn(cvtot, rtot), cvtot);
    const double rho = v[0];
    const double pres = __dmul_rn(P.PRE00, pow(__ddiv_rn(__dmul_rn(v[4 * st], rtot), P.PRE00), cpovcv));
    const double temp = __ddiv_rn(pres, __dmul_rn(rho, rtot));
    v[0] = __ddiv_rn(v[1 * st], rho);
    v[1 * st] = __ddiv_rn(v[2 * st], rho);
    v[2 * st] = __ddiv_rn(v[3 * st], rho);
    v[3 * st] = temp;
    v[4 * st] = pres;
  } else {
    const double cvovcp = __ddiv_rn(cvtot, __dadd_rn(cvtot, rtot));
    const double pres = v[4 * st];
    const double rho = __ddiv_rn(pres, __dmul_rn(rtot, v[3 * st]));
    const double rhot = __dmul_rn(__ddiv_rn(P.PRE00, rtot), pow(__ddiv_rn(pres, P.PRE00), cvovcp));
    v[4 * st] = rhot;
    v[3 * st] = __dmul_rn(v[2 * st], rho);
    v[2 * st] = __dmul_rn(v[1 * st], rho);
    v[1 * st] = __dmul_rn(v[0], rho);
    v[0] = rho;
  }
}

// dir = 0: v3dg -> (state_trans when trans) -> bufs;  dir = 1: bufr -> (state_trans_inv when trans) -> v3dg
__global__ void __launch_bounds__(256) grd_buf_tiled_kernel(TransposeDims d, StateTransParams T, int trans, int dir,
                                                            double *__restrict__ v3dg, double *__restrict__ buf) {
  extern __shared__ double tile[];   // [nv3d][32 columns][33]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32, m = blockIdx.z;
  const int nij1 = nij1_of(d, m);
  const size_t npts = (size_t)d.nlev * d.nlon * d.nlat;
  constexpr int TS = 32 * 33;
  auto gcol = [&](int i) -> size_t {   // first level of column i of rank m in a member-major field
    const int j = m + d.np * i;
    const int ilon = j % d.nlon, ilat = j / d.nlon;
    return (size_t)d.nlev * (ilon + (size_t)d.nlon * ilat);
  };
  auto bidx = [&](int i, int k, int n) -> size_t {
    return (size_t)i + (size_t)d.nij1max * ((size_t)k + (size_t)d.nlev * n + (size_t)d.nlevall * m);
  };
  if (dir == 0) {
    for (int n = 0; n < d.nv3d; ++n)
      for (int ii = ty; ii < 32; ii += 8) {
        const int i = i0 + ii, k = k0 + tx;
        tile[n * TS + ii * 33 + tx] = (i < nij1 && k < d.nlev) ? v3dg[gcol(i) + k + npts * n] : 0.0;
      }
  } else {
    for (int n = 0; n < d.nv3d; ++n)
      for (int kk = ty; kk < 32; kk += 8) {
        const int i = i0 + tx, k = k0 + kk;
        tile[n * TS + tx * 33 + kk] = (i < nij1 && k < d.nlev) ? buf[bidx(i, k, n)] : 0.0;
      }
  }
  __syncthreads();
  if (trans) {
    for (int e = ty * 32 + tx; e < 1024; e += 256) {
      const int ii = e >> 5, kk = e & 31;
      if (i0 + ii < nij1 && k0 + kk < d.nlev) state_trans_point(T, tile + ii * 33 + kk, TS, dir);
    }
    __syncthreads();
  }
  if (dir == 0) {
    for (int n = 0; n < d.nv3d; ++n)
      for (int kk = ty; kk < 32; kk += 8) {
        const int i = i0 + tx, k = k0 + kk;
        if (i < d.nij1max && k < d.nlev) buf[bidx(i, k, n)] = (i < nij1) ? tile[n * TS + tx * 33 + kk] : -9.99e33;   // undef padding row
      }
  } else {
    for (int n = 0; n < d.nv3d; ++n)
      for (int ii = ty; ii < 32; ii += 8) {
        const int i = i0 + ii, k = k0 + tx;
        if (i < nij1 && k < d.nlev) v3dg[gcol(i) + k + npts * n] = tile[n * TS + ii * 33 + tx];
      }
  }
}
// ---------------------------------------------------------------------------------------------
// One-pass member <-> grid transposes over peer memory (scatter/gather_grd_mpi_alltoall twins,
// common_mpi_scale.f90:1279-1396, with grd_to_buf / buf_to_grd :1428-1476 and the exchange folded in): every rank
// READS its own arrays and WRITES straight into the receiving rank's array -- its own memory, or a peer's over
// NVLink (CUDA IPC mapping) -- so pack, all-to-all and unpack are ONE pass over HBM and the transfer overlaps the
// tile loop.  state_trans / state_trans_inv (common_scale.f90:1181-1280) are applied on the tile in shared memory.
//   dir 0 (scatter): this rank holds the member that goes to slot `slot0` as v3dg(nlev,nlon,nlat,nv3d);
//                    blockIdx.z = receiving rank m: columns j = m + np i  ->  v3d_m(i, k, slot0, n)
//   dir 1 (gather):  this rank holds v3d(nij1,nlev,nens,nv3d); blockIdx.z = q: member slot0 + q lives on rank q:
//                    v3d(i, k, slot0 + q, n) -> v3dg_q at column j = myrank + np i
struct PeerPtrs {
  double *p3[16];   // scatter: v3d of rank m;  gather: v3dg of rank q
  double *p2[16];   // the 2-D twins (may be null)
};
__global__ void __launch_bounds__(256) grd_ens_p2p_kernel(TransposeDims d, StateTransParams T, int trans, int dir, int myrank,
                                                          int nens, int slot0, const double *__restrict__ src, PeerPtrs peers) {
  extern __shared__ double tile[];   // [nv3d][32 columns][33]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32, z = blockIdx.z;
  const int colrank = dir == 0 ? z : myrank;          // whose columns this CTA moves
  const int nij1 = nij1_of(d, colrank);
  if (i0 >= nij1) return;
  const size_t npts = (size_t)d.nlev * d.nlon * d.nlat;
  constexpr int TS = 32 * 33;
  auto gcol = [&](int i) -> size_t {   // first level of column i of rank `colrank` in a member-major field
    const int j = colrank + d.np * i;
    const int ilon = j % d.nlon, ilat = j / d.nlon;
    return (size_t)d.nlev * (ilon + (size_t)d.nlon * ilat);
  };
  auto eidx = [&](int i, int k, int slot, int n) -> size_t {   // v3d(nij1, nlev, nens, nv3d) of rank `colrank`
    return (size_t)i + (size_t)nij1 * ((size_t)k + (size_t)d.nlev * ((size_t)slot + (size_t)nens * n));
  };
  double *dst = peers.p3[z];
  // every load of the tile is issued before the first shared-memory store (4 x nv3d values per thread in flight):
  // the kernel is pure data movement and lives on bytes in flight
  constexpr int NVU = 16;
  double reg[NVU][4];
  if (dir == 0) {
    size_t gc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) gc[q] = (i0 + ty + 8 * q < nij1 && k0 + tx < d.nlev) ? gcol(i0 + ty + 8 * q) + k0 + tx : (size_t)-1;
#pragma unroll
    for (int n = 0; n < NVU; ++n)
#pragma unroll
      for (int q = 0; q < 4; ++q) reg[n][q] = (n < d.nv3d && gc[q] != (size_t)-1) ? src[gc[q] + npts * n] : 0.0;
#pragma unroll
    for (int n = 0; n < NVU; ++n)
      if (n < d.nv3d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tile[n * TS + (ty + 8 * q) * 33 + tx] = reg[n][q];
      }
  } else {
    const bool oki = i0 + tx < nij1;
#pragma unroll
    for (int n = 0; n < NVU; ++n)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        reg[n][q] = (n < d.nv3d && oki && k0 + ty + 8 * q < d.nlev) ? src[eidx(i0 + tx, k0 + ty + 8 * q, slot0 + z, n)] : 0.0;
#pragma unroll
    for (int n = 0; n < NVU; ++n)
      if (n < d.nv3d) {
#pragma unroll
        for (int q = 0; q < 4; ++q) tile[n * TS + tx * 33 + ty + 8 * q] = reg[n][q];
      }
  }
  __syncthreads();
  if (trans) {
    for (int e = ty * 32 + tx; e < 1024; e += 256) {
      const int ii = e >> 5, kk = e & 31;
      if (i0 + ii < nij1 && k0 + kk < d.nlev) state_trans_point(T, tile + ii * 33 + kk, TS, dir);
    }
    __syncthreads();
  }
  if (dir == 0) {
    for (int n = 0; n < d.nv3d; ++n)
      for (int kk = ty; kk < 32; kk += 8) {
        const int i = i0 + tx, k = k0 + kk;
        if (i < nij1 && k < d.nlev) dst[eidx(i, k, slot0, n)] = tile[n * TS + tx * 33 + kk];
      }
  } else {
    for (int n = 0; n < d.nv3d; ++n)
      for (int ii = ty; ii < 32; ii += 8) {
        const int i = i0 + ii, k = k0 + tx;
        if (i < nij1 && k < d.nlev) dst[gcol(i) + k + npts * n] = tile[n * TS + ii * 33 + tx];
      }
  }
}
// The same transposes with the tile loaded by cp.async (LDGSTS: no register staging, 8-byte elements straight to their
// transposed place in shared memory) and KT = 16 levels per tile: 48 KB of shared memory for 11 variables -> four CTAs
// per SM whose load / transform / store phases overlap (the register-staged 32-level tile above: 128 registers, 93 KB,
// two CTAs, 25 % warps active, latency-bound).  The level tiles of a column block are ADJACENT in launch order
// (blockIdx.x = level tile), so the 64-byte DRAM granules two level tiles share at a 480-byte column's cut are fetched
// once and hit L2 for the neighbour (ncu of the kernel above: 1.40x the algorithmic reads).
__device__ __forceinline__ void p2p_cp_async8(double *smem, const double *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
template <int KT>
__global__ void __launch_bounds__(256) grd_ens_p2p_async_kernel(TransposeDims d, StateTransParams T, int trans, int dir, int myrank,
                                                                int nens, int slot0, const double *__restrict__ src, PeerPtrs peers) {
  extern __shared__ double tile[];   // [nv3d][32 columns][KT + 1]
  constexpr int RS = KT + 1, TS = 32 * RS, NE = 32 * KT, PER = NE / 256;
  static_assert(NE % 256 == 0, "whole passes of the 256 threads over a tile plane");
  const int t = threadIdx.x;
  const int k0 = blockIdx.x * KT, i0 = blockIdx.y * 32, z = blockIdx.z;
  const int colrank = dir == 0 ? z : myrank;          // whose columns this CTA moves
  const int nij1 = nij1_of(d, colrank);
  if (i0 >= nij1) return;
  const size_t npts = (size_t)d.nlev * d.nlon * d.nlat;
  auto gcol = [&](int i) -> size_t {   // first level of column i of rank `colrank` in a member-major field
    const int j = colrank + d.np * i;
    const int ilon = j % d.nlon, ilat = j / d.nlon;
    return (size_t)d.nlev * (ilon + (size_t)d.nlon * ilat);
  };
  const size_t estride_n = (size_t)nij1 * d.nlev * nens;   // v3d(nij1, nlev, nens, nv3d) of rank `colrank`: variable stride
  auto eidx0 = [&](int i, int k, int slot) -> size_t { return (size_t)i + (size_t)nij1 * ((size_t)k + (size_t)d.nlev * slot); };
  double *dst = peers.p3[z];
  // ---- load: PER elements per thread and variable; member-major side: lanes along the levels, ensemble side: along i ----
  size_t goff[PER];
  int soff[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int e = t + 256 * q;
    int c, l;
    if (dir == 0) { l = e % KT; c = e / KT; } else { c = e & 31; l = e >> 5; }
    soff[q] = c * RS + l;
    const bool ok = i0 + c < nij1 && k0 + l < d.nlev;
    goff[q] = !ok ? (size_t)-1 : dir == 0 ? gcol(i0 + c) + k0 + l : eidx0(i0 + c, k0 + l, slot0 + z);
  }
  const size_t gstride = dir == 0 ? npts : estride_n;
  for (int n = 0; n < d.nv3d; ++n) {
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      if (goff[q] != (size_t)-1) p2p_cp_async8(&tile[n * TS + soff[q]], src + goff[q] + gstride * n);
      else tile[n * TS + soff[q]] = 0.0;
    }
  }
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (trans) {
    for (int e = t; e < NE; e += 256) {
      const int ii = e / KT, kk = e % KT;
      if (i0 + ii < nij1 && k0 + kk < d.nlev) state_trans_point(T, tile + ii * RS + kk, TS, dir);
    }
    __syncthreads();
  }
  // ---- store: ensemble side along i (256-byte rows), member-major side along the levels ----
  if (dir == 0) {
    const int tx = t & 31, ty = t >> 5;
    if (i0 + tx < nij1)
      for (int n = 0; n < d.nv3d; ++n)
        for (int kk = ty; kk < KT; kk += 8)
          if (k0 + kk < d.nlev) dst[eidx0(i0 + tx, k0 + kk, slot0) + estride_n * n] = tile[n * TS + tx * RS + kk];
  } else {
    size_t go[PER];
    int so[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int e = t + 256 * q, l = e % KT, c = e / KT;
      so[q] = c * RS + l;
      go[q] = (i0 + c < nij1 && k0 + l < d.nlev) ? gcol(i0 + c) + k0 + l : (size_t)-1;
    }
    for (int n = 0; n < d.nv3d; ++n)
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (go[q] != (size_t)-1) dst[go[q] + npts * n] = tile[n * TS + so[q]];
  }
}
// 2-D variables of the one-pass transposes: v2dg(nlon,nlat,nv2d) <-> v2d(nij1,nens,nv2d)
__global__ void grd_ens_p2p_2d_kernel(TransposeDims d, int dir, int myrank, int nens, int slot0, int npeers,
                                      const double *__restrict__ src, PeerPtrs peers) {
  const size_t total = (size_t)d.nij1max * d.nv2d * npeers;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % d.nij1max);
    const size_t t = idx / d.nij1max;
    const int n = (int)(t % d.nv2d), z = (int)(t / d.nv2d);
    const int colrank = dir == 0 ? z : myrank, nij1 = nij1_of(d, colrank);
    if (i >= nij1 || !peers.p2[z]) continue;
    const int j = colrank + d.np * i;
    const size_t g = (j % d.nlon) + (size_t)d.nlon * ((j / d.nlon) + (size_t)d.nlat * n);
    if (dir == 0) peers.p2[z][(size_t)i + (size_t)nij1 * ((size_t)slot0 + (size_t)nens * n)] = src[g];
    else peers.p2[z][g] = src[(size_t)i + (size_t)nij1 * ((size_t)(slot0 + z) + (size_t)nens * n)];
  }
}

// 2-D variables of the pack / unpack (v2dg(nlon,nlat,nv2d) <-> rows nlev*nv3d.. of the buffer)
__global__ void grd_buf_2d_kernel(TransposeDims d, int dir, double *__restrict__ v2dg, double *__restrict__ buf) {
  const size_t total = (size_t)d.nij1max * d.nv2d * d.np;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % d.nij1max);
    const size_t t = idx / d.nij1max;
    const int n = (int)(t % d.nv2d), m = (int)(t / d.nv2d);
    const size_t b = (size_t)i + (size_t)d.nij1max * ((size_t)d.nlev * d.nv3d + n + (size_t)d.nlevall * m);
    if (i < nij1_of(d, m)) {
      const int j = m + d.np * i;
      const size_t g = (j % d.nlon) + (size_t)d.nlon * ((j / d.nlon) + (size_t)d.nlat * n);
      if (dir == 0) buf[b] = v2dg[g]; else v2dg[g] = buf[b];
    } else if (dir == 0) {
      buf[b] = -9.99e33;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// monit_dep (scale/common/common_obs_scale.f90:1851-1895): per-element departure statistics.  Stage 1: every
// CTA reduces its grid-stride slice into 16 (count, sum, sum of squares) bins with a fixed tree; stage 2: one
// CTA adds the per-CTA partials in CTA order.  Deterministic.
__device__ __forceinline__ int monit_uid(int elm) {   // uid_obs(elm) - 1 with Tv -> T, RE0 -> REF; -1: unknown
  if (elm == 3074) elm = 3073;
  if (elm == 4004) elm = 4001;
  switch (elm) {
    case 2819: return 0; case 2820: return 1; case 3073: return 2; case 3074: return 3; case 3330: return 4;
    case 3331: return 5; case 14593: return 6; case 19999: return 7; case 4001: return 8; case 4004: return 9;
    case 4002: return 10; case 4003: return 11; case 8800: return 12; case 99991: return 13; case 99992: return 14;
    case 99993: return 15; default: return -1;
  }
}
__global__ void __launch_bounds__(256) monit_partial_kernel(int nobs, const int *__restrict__ elm, const double *__restrict__ dep,
                                                            const int *__restrict__ qc, double *__restrict__ part) {
  __shared__ double sm[3][16][8];   // [quantity][bin][warp]
  double cnt[16], sum[16], sq[16];
#pragma unroll
  for (int b = 0; b < 16; ++b) cnt[b] = sum[b] = sq[b] = 0.0;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nobs; n += gridDim.x * blockDim.x) {
    if (qc[n] != 0) continue;
    const int u = monit_uid(elm[n]);
    const double d = dep[n];
#pragma unroll
    for (int b = 0; b < 16; ++b)
      if (b == u) {
        cnt[b] += 1.0;
        sum[b] += d;
        sq[b] = fma(d, d, sq[b]);
      }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int b = 0; b < 16; ++b) {
    const double c = warp_sum(cnt[b]), s = warp_sum(sum[b]), q = warp_sum(sq[b]);
    if (lane == 0) {
      sm[0][b][w] = c;
      sm[1][b][w] = s;
      sm[2][b][w] = q;
    }
  }
  __syncthreads();
  if (threadIdx.x < 48) {
    const int qn = threadIdx.x / 16, b = threadIdx.x % 16;
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += sm[qn][b][i];
    part[(size_t)blockIdx.x * 48 + threadIdx.x] = a;
  }
}
__global__ void monit_final_kernel(int nblocks, const double *__restrict__ part, int *__restrict__ nobs_out,
                                   double *__restrict__ bias, double *__restrict__ rmse) {
  const int b = threadIdx.x;
  if (b >= 16) return;
  double c = 0.0, s = 0.0, q = 0.0;
  for (int i = 0; i < nblocks; ++i) {
    c += part[(size_t)i * 48 + b];
    s += part[(size_t)i * 48 + 16 + b];
    q += part[(size_t)i * 48 + 32 + b];
  }
  nobs_out[b] = (int)c;
  bias[b] = c > 0.0 ? s / c : -9.99e33;
  rmse[b] = c > 0.0 ? sqrt(q / c) : -9.99e33;
}

}  // namespace letkf
