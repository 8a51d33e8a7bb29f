// search.cuh -- device-side local-observation search (twin of obs_local,
// scale/letkf/letkf_tools.f90:1325-1759, obs_local_cal :1793-1906, obs_local_range :1765,
// ij_obsgrd_ext / obs_choose_ext scale/letkf/letkf_obs.f90:1209,1262).
//
// One CTA searches for one grid point.  Observations are bucket-sorted on the device in
// exactly the order of obsda_sort (ctype-major, bucket row j, bucket i, arrival order);
// `bstart` is the global exclusive prefix sum over all buckets, so that
//   ac_ext(i, j) of ctype ic == bstart[boff(ic) + (j-1)*ngrdext_i + i]     (i = 0..ngrdext_i).
//
// Selection semantics (decomposition independent): per master group, every
// observation passing obs_local_cal, then -- if the group has an obs-number limit N -- the N
// smallest by the criterion key.  The reference reaches the same set through its incremental
// square search + QUICKSELECT_arg; here the N-th key is found with a storage-free 8-bit radix
// select that re-evaluates the candidates, and ties are broken by scan order.
//
// All distance arithmetic uses explicit round-to-nearest intrinsics: no FMA contraction
// (SURVEY.md Appendix A.13), so thresholds are bit-identical to the CPU oracle.
#pragma once
#include "common.cuh"

namespace letkf {

constexpr int kMaxCtype = 64;
constexpr int kMaxGroup = 64;
constexpr int kMaxMerge = 4;
constexpr int kSegMax = 256;

struct __align__(16) ObsRec {
  double ri, rj;
  double vc;    // vertical coordinate: log(lev) | log(dat) for PS | height for PHARAD
  double err;
};

struct CtypeDev {
  double hori_loc, vert_loc, vconst;
  double grdspc_i, grdspc_j;
  int vmode;      // 0: no vertical localisation, 1: |vc - log p|, 2: |vconst - log p|, 3: |vc - z|
  int varlocal;   // 0-based uid_obs_varlocal
  int ngrd_i, ngrd_j, ngrdsch_i, ngrdsch_j, ngrdext_i, ngrdext_j;
  int boff;       // first bucket of this ctype in bstart
  int tot;
  int elm_u, typ;
};

struct GroupDev {
  int n;
  int ic[kMaxMerge];
  int limit;      // MAX_NOBS_PER_GRID of the master type (<= 0: no limit)
};

struct SearchTables {
  CtypeDev ct[kMaxCtype];
  GroupDev grp[kMaxGroup];
  int nctype, ngroup;
  int criterion;
  int IHALO, JHALO, nlon, nlat;
  double DX, DY, dzf, dzf2;
};

struct Point {
  double ri, rj, lp, rz;
};

struct LocalList {   // per-CTA scratch in global memory
  int *iob;
  double *rdiag, *rloc;
  int cap;
};

struct SearchSmem {
  int seg_start[kSegMax];
  int seg_cum[kSegMax + 1];
  int hist[256];
  int red[kMaxWarps];
  int misc[8];
};

__device__ __forceinline__ int obsgrd_index(double r, int halo, int ngrd, int nl, int nsch) {
  // ceiling((r - HALO - 0.5) * ngrd / nl) + ngrdsch   (letkf_obs.f90:1220-1224)
  const double x = __ddiv_rn(__dmul_rn(__dsub_rn(__dsub_rn(r, (double)halo), 0.5), (double)ngrd), (double)nl);
  return (int)ceil(x) + nsch;
}

struct Rect {
  int imin, imax, jmin, jmax;
};

// rectangle of buckets covering +-half_i / +-half_j grid units around the point
__device__ __forceinline__ Rect rect_of(const SearchTables &T, const CtypeDev &c, const Point &p,
                                        double half_i, double half_j) {
  Rect r;
  r.imin = obsgrd_index(__dsub_rn(p.ri, half_i), T.IHALO, c.ngrd_i, T.nlon, c.ngrdsch_i);
  r.jmin = obsgrd_index(__dsub_rn(p.rj, half_j), T.JHALO, c.ngrd_j, T.nlat, c.ngrdsch_j);
  r.imax = obsgrd_index(__dadd_rn(p.ri, half_i), T.IHALO, c.ngrd_i, T.nlon, c.ngrdsch_i);
  r.jmax = obsgrd_index(__dadd_rn(p.rj, half_j), T.JHALO, c.ngrd_j, T.nlat, c.ngrdsch_j);
  return r;
}
__device__ __forceinline__ Rect clamp_rect(Rect r, const CtypeDev &c) {
  r.imin = max(r.imin, 1);
  r.jmin = max(r.jmin, 1);
  r.imax = min(r.imax, c.ngrdext_i);
  r.jmax = min(r.jmax, c.ngrdext_j);
  return r;
}
// obs_local_range (letkf_tools.f90:1765-1788)
__device__ __forceinline__ Rect cutoff_rect(const SearchTables &T, const CtypeDev &c, const Point &p) {
  const double dzi = __ddiv_rn(__dmul_rn(c.hori_loc, T.dzf), T.DX);
  const double dzj = __ddiv_rn(__dmul_rn(c.hori_loc, T.dzf), T.DY);
  return clamp_rect(rect_of(T, c, p, dzi, dzj), c);
}

// obs_local_cal geometry (letkf_tools.f90:1852-1895): true if the observation survives the
// three cut-offs; ndist = normalised 3-D distance squared.
__device__ __forceinline__ bool obs_geom(const SearchTables &T, const CtypeDev &c, const Point &p,
                                         const ObsRec &o, double &ndist) {
  double nd_v;
  if (c.vmode == 0) {
    nd_v = 0.0;
  } else if (c.vmode == 1) {
    nd_v = __ddiv_rn(fabs(__dsub_rn(o.vc, p.lp)), c.vert_loc);
  } else if (c.vmode == 2) {
    nd_v = __ddiv_rn(fabs(__dsub_rn(c.vconst, p.lp)), c.vert_loc);
  } else {
    nd_v = __ddiv_rn(fabs(__dsub_rn(o.vc, p.rz)), c.vert_loc);
  }
  if (nd_v > T.dzf) return false;
  const double rdx = __dmul_rn(__dsub_rn(p.ri, o.ri), T.DX);
  const double rdy = __dmul_rn(__dsub_rn(p.rj, o.rj), T.DY);
  const double nd_h =
      __ddiv_rn(__dsqrt_rn(__dadd_rn(__dmul_rn(rdx, rdx), __dmul_rn(rdy, rdy))), c.hori_loc);
  if (nd_h > T.dzf) return false;
  ndist = __dadd_rn(__dmul_rn(nd_h, nd_h), __dmul_rn(nd_v, nd_v));
  if (ndist > T.dzf2) return false;
  return true;
}

__device__ __forceinline__ int block_excl_scan_i(int v, int *red, int &total) {
  // exclusive scan over threadIdx.x order; every thread calls
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(LETKF_FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();
  if (lane == 31) red[w] = x;
  __syncthreads();
  int base = 0, t = 0;
  for (int i = 0; i < nw; ++i) {
    const int c = red[i];
    if (i < w) base += c;
    t += c;
  }
  total = t;
  return base + x - v;
}

// Visit every observation of ctype `ic` inside bucket rectangle `r`.  `f(iob)` is invoked by
// ALL threads once per tile of blockDim.x candidates (iob = -1 on idle lanes) so that it may
// contain block-wide collectives.  Rows are processed in batches of kSegMax.
template <class F>
__device__ __forceinline__ void scan_rect(const SearchTables &T, const int *__restrict__ bstart,
                                          int ic, Rect r, SearchSmem &S, F &&f) {
  const CtypeDev &c = T.ct[ic];
  if (c.tot == 0 || r.imin > r.imax || r.jmin > r.jmax) return;
  for (int jb = r.jmin; jb <= r.jmax; jb += kSegMax) {
    const int nrows = min(kSegMax, r.jmax - jb + 1);
    int total = 0;
    for (int base = 0; base < nrows; base += blockDim.x) {   // nrows <= 256 = max blockDim
      const int row = base + threadIdx.x;
      int len = 0, st = 0;
      if (row < nrows) {
        const int rowoff = c.boff + (jb + row - 1) * c.ngrdext_i;
        st = bstart[rowoff + r.imin - 1];
        len = bstart[rowoff + r.imax] - st;
      }
      int tot_part;
      const int ex = block_excl_scan_i(len, S.red, tot_part);
      if (row < nrows) {
        S.seg_start[row] = st;
        S.seg_cum[row] = total + ex;
      }
      total += tot_part;
    }
    if (threadIdx.x == 0) S.seg_cum[nrows] = total;
    __syncthreads();
    for (int base = 0; base < total; base += blockDim.x) {
      const int v = base + threadIdx.x;
      int iob = -1;
      if (v < total) {
        int lo = 0, hi = nrows - 1;   // last segment with cum <= v
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (S.seg_cum[mid] <= v) lo = mid; else hi = mid - 1;
        }
        iob = S.seg_start[lo] + (v - S.seg_cum[lo]);
      }
      f(iob);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ unsigned long long key_bits(double x) {
  return (unsigned long long)__double_as_longlong(x);
}

// Twin of obs_local for one point and one variable-localisation group.  `vlfac[ic]` is
// var_local(nvar, uid_obs_varlocal(elm_ctype(ic))) for the group's representative variable.
// Fills L (ordered: group-major, merged-ctype-major, bucket scan order) and returns nobsl,
// or -1 if L.cap is too small.
__device__ __forceinline__ int search_point(const SearchTables &T, const ObsRec *__restrict__ rec,
                                            const int *__restrict__ bstart,
                                            const double *__restrict__ vlfac, const Point &p,
                                            LocalList &L, SearchSmem &S) {
  int nobsl = 0;
  bool overflow = false;
  const double tiny = 2.2250738585072014e-308;   // tiny(var_local)
  for (int g = 0; g < T.ngroup; ++g) {
    const GroupDev &G = T.grp[g];
    const int N = G.limit;
    if (N <= 0) {
      // ---- no obs-number limit (letkf_tools.f90:1438-1476) -------------------------------
      for (int icm = 0; icm < G.n; ++icm) {
        const int ic = G.ic[icm];
        const CtypeDev &c = T.ct[ic];
        const double vl = vlfac[ic];
        if (vl < tiny) continue;
        scan_rect(T, bstart, ic, cutoff_rect(T, c, p), S, [&](int iob) {
          double ndist = 0.0, err = 0.0;
          bool ok = false;
          if (iob >= 0) {
            const ObsRec o = rec[iob];
            ok = obs_geom(T, c, p, o, ndist);
            err = o.err;
          }
          int tot;
          const int rk = block_rank(ok, S.red, tot);
          if (ok) {
            const int pos = nobsl + rk;
            if (pos < L.cap) {
              const double rl = vl * exp(-0.5 * ndist);
              L.iob[pos] = iob;
              L.rloc[pos] = rl;
              L.rdiag[pos] = err * err / rl;
            }
          }
          nobsl += tot;
          if (nobsl > L.cap) overflow = true;
        });
      }
      continue;
    }
    // ---- obs-number limit N ----------------------------------------------------------------
    const int crit = T.criterion;
    // criterion 1: incremental square search (letkf_tools.f90:1502-1602); the rectangle only
    // bounds the candidate set, the selected set does not depend on the schedule of q.
    const CtypeDev &cm = T.ct[G.ic[0]];
    double search_incr0 = __ddiv_rn(__dmul_rn(cm.hori_loc, T.dzf), 8.0);
    search_incr0 = fmax(search_incr0, fmax(cm.grdspc_i, cm.grdspc_j));
    Rect rq[kMaxMerge];
    bool reach_cutoff = true;
    double dcf2 = T.dzf2;
    int count = 0;
    for (int q = (crit == 1 ? 1 : 1 << 20);; ++q) {
      reach_cutoff = true;
      for (int icm = 0; icm < G.n; ++icm) {
        const CtypeDev &c = T.ct[G.ic[icm]];
        const Rect rc = cutoff_rect(T, c, p);
        if (crit == 1 && q < (1 << 20)) {
          const double incr = (icm == 0) ? search_incr0 : search_incr0 / cm.hori_loc * c.hori_loc;
          const Rect r = rect_of(T, c, p, incr / T.DX * q, incr / T.DY * q);
          if (r.imin <= rc.imin && r.imax >= rc.imax && r.jmin <= rc.jmin && r.jmax >= rc.jmax) {
            rq[icm] = rc;
          } else {
            rq[icm] = clamp_rect(r, c);
            reach_cutoff = false;
          }
        } else {
          rq[icm] = rc;
        }
      }
      if (!reach_cutoff) {
        const double f = search_incr0 * q / cm.hori_loc;
        dcf2 = f * f;
      } else {
        dcf2 = T.dzf2;
      }
      // count valid candidates inside the current radius
      int cnt = 0;
      for (int icm = 0; icm < G.n; ++icm) {
        const int ic = G.ic[icm];
        const CtypeDev &c = T.ct[ic];
        if (vlfac[ic] < tiny) continue;
        scan_rect(T, bstart, ic, rq[icm], S, [&](int iob) {
          if (iob >= 0) {
            double ndist;
            if (obs_geom(T, c, p, rec[iob], ndist) && (reach_cutoff || !(ndist > dcf2))) ++cnt;
          }
        });
      }
      count = block_sum_i(cnt, S.red);
      if (reach_cutoff || count >= N) break;
    }
    if (count == 0) continue;
    // key of a candidate under the active criterion (ascending selection)
    auto cand_key = [&](const CtypeDev &c, double vl, int iob, unsigned long long &key, double &rl,
                        double &rd) -> bool {
      const ObsRec o = rec[iob];
      double ndist;
      if (!obs_geom(T, c, p, o, ndist)) return false;
      if (!reach_cutoff && ndist > dcf2) return false;
      rl = vl * exp(-0.5 * ndist);
      rd = o.err * o.err / rl;
      key = (crit == 1) ? key_bits(ndist) : (crit == 2) ? ~key_bits(rl) : key_bits(rd);
      return true;
    };
    unsigned long long tau = ~0ull;   // select key <= tau, plus `eq_budget` of key == tau_eq
    unsigned long long tau_eq = 0ull;
    int eq_budget = 0;
    bool exact = false;
    if (count > N) {
      unsigned long long prefix = 0ull;
      int remaining = N;
      for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) S.hist[i] = 0;
        __syncthreads();
        for (int icm = 0; icm < G.n; ++icm) {
          const int ic = G.ic[icm];
          const CtypeDev &c = T.ct[ic];
          const double vl = vlfac[ic];
          if (vl < tiny) continue;
          scan_rect(T, bstart, ic, rq[icm], S, [&](int iob) {
            if (iob >= 0) {
              unsigned long long key;
              double rl, rd;
              if (cand_key(c, vl, iob, key, rl, rd)) {
                const bool match = (pass == 7) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
                if (match) atomicAdd(&S.hist[(int)((key >> shift) & 255ull)], 1);
              }
            }
          });
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          int cum = 0, b = 0;
          for (; b < 256; ++b) {
            if (cum + S.hist[b] >= remaining) break;
            cum += S.hist[b];
          }
          S.misc[0] = b;
          S.misc[1] = cum;
          S.misc[2] = S.hist[b];
        }
        __syncthreads();
        const int b = S.misc[0];
        remaining -= S.misc[1];
        prefix |= ((unsigned long long)b) << shift;
        const int inbin = S.misc[2];
        __syncthreads();
        if (inbin == remaining) {   // the whole bin is selected: no tie to resolve
          tau = prefix | ((shift > 0) ? ((1ull << shift) - 1ull) : 0ull);
          exact = false;
          eq_budget = 0;
          remaining = 0;
          break;
        }
        if (pass == 0) {   // exact N-th key with ties
          exact = true;
          tau_eq = prefix;
          eq_budget = remaining;
        }
      }
    }
    // collect
    int eq_taken = 0;
    for (int icm = 0; icm < G.n; ++icm) {
      const int ic = G.ic[icm];
      const CtypeDev &c = T.ct[ic];
      const double vl = vlfac[ic];
      if (vl < tiny) continue;
      scan_rect(T, bstart, ic, rq[icm], S, [&](int iob) {
        unsigned long long key = 0ull;
        double rl = 0.0, rd = 0.0;
        bool ok = false;
        if (iob >= 0) ok = cand_key(c, vl, iob, key, rl, rd);
        bool take;
        if (exact) {
          const bool iseq = ok && key == tau_eq;
          int toteq;
          const int rkeq = block_rank(iseq, S.red, toteq);
          take = ok && (key < tau_eq || (iseq && eq_taken + rkeq < eq_budget));
          eq_taken += toteq;
        } else {
          take = ok && key <= tau;
        }
        int tot;
        const int rk = block_rank(take, S.red, tot);
        if (take) {
          const int pos = nobsl + rk;
          if (pos < L.cap) {
            L.iob[pos] = iob;
            L.rloc[pos] = rl;
            L.rdiag[pos] = rd;
          }
        }
        nobsl += tot;
        if (nobsl > L.cap) overflow = true;
      });
    }
  }
  __syncthreads();
  return overflow ? -1 : nobsl;
}

}  // namespace letkf
