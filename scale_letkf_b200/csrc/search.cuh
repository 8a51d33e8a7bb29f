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
  // conservative pre-filters (set_obs): a candidate failing one of them fails the exact test of
  // obs_local_cal too, so only survivors pay for the divisions and the square root
  double vmax;    // dist_zero_fac * vert_loc * (1 + 2^-40)
  double hmax2;   // (dist_zero_fac * hori_loc)^2 * (1 + 2^-40)
  double ih2, iv2;   // 1 / hori_loc^2, 1 / vert_loc^2 (0 when vmode == 0)
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
  int cand_cap_limit;   // candidate-buffer entries the obs-number-limited search may use (<= LocalList::ccap)
  int IHALO, JHALO, nlon, nlat;
  double DX, DY, dzf, dzf2;
};

struct Point {
  double ri, rj, lp, rz;
};

struct Rect {
  int imin, imax, jmin, jmax;
};

struct LocalList {   // per-CTA scratch in global memory
  int *iob;
  double *rdiag, *rloc;
  int cap;
  // candidate buffer of the obs-number-limited search: normalised distance^2 and packed
  // (slot << 28 | sorted obs index) of every candidate inside the current search radius
  double *cnd;
  unsigned *cpk;
  int ccap;
};

constexpr int kMaxScan = 8;    // ctypes scanned together (a run of unlimited groups / one merged group)

struct ScanList {
  int n;
  int ic[kMaxScan];
  Rect r[kMaxScan];
};

struct SearchSmem {
  int seg_start[kSegMax];
  int seg_cum[kSegMax + 1];
  int hist[256];
  int red[kMaxWarps];
  int misc[8];
  unsigned char seg_slot[kSegMax];
  int wcnt[kMaxWarps];   // survivors per warp chunk (fill_window)
  int wsel[kMaxWarps];   // selected per warp chunk / ties per warp chunk
  ScanList sl;           // the scan in progress (built by one thread per ctype, see build_scan_*)
  Rect tr[kMaxScan];     // per-entry rectangle / validity before compaction into `sl`
  int tv[kMaxScan];
};

__device__ __forceinline__ int obsgrd_index(double r, int halo, int ngrd, int nl, int nsch) {
  // ceiling((r - HALO - 0.5) * ngrd / nl) + ngrdsch   (letkf_obs.f90:1220-1224)
  const double x = __ddiv_rn(__dmul_rn(__dsub_rn(__dsub_rn(r, (double)halo), 0.5), (double)ngrd), (double)nl);
  return (int)ceil(x) + nsch;
}

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

// ---- multi-ctype segment scan --------------------------------------------------------------------
// The candidate set of a ScanList is the concatenation, slot-major then bucket-row-major, of the
// contiguous sorted-index ranges [bstart(imin-1, j), bstart(imax, j)) -- exactly the order in which
// obs_local visits them.  build_segments() lays rows [row0, row0 + kSegMax) of that list out in
// shared memory (start, exclusive prefix, slot); returns the number of candidates in the batch.
__device__ __forceinline__ int scan_rows_total(const ScanList &SL) {
  int t = 0;
  for (int m = 0; m < SL.n; ++m) t += SL.r[m].jmax - SL.r[m].jmin + 1;
  return t;
}

__device__ __forceinline__ int build_segments(const SearchTables &T, const int *__restrict__ bstart,
                                              const ScanList &SL, int row0, int nrows, SearchSmem &S) {
  int total = 0;
  for (int base = 0; base < nrows; base += blockDim.x) {
    const int row = base + threadIdx.x;
    int len = 0, st = 0, slot = 0;
    if (row < nrows) {
      int rr = row0 + row;
      for (; slot < SL.n - 1; ++slot) {
        const int nr = SL.r[slot].jmax - SL.r[slot].jmin + 1;
        if (rr < nr) break;
        rr -= nr;
      }
      const CtypeDev &c = T.ct[SL.ic[slot]];
      const Rect &r = SL.r[slot];
      const int rowoff = c.boff + (r.jmin + rr - 1) * c.ngrdext_i;
      st = bstart[rowoff + r.imin - 1];
      len = bstart[rowoff + r.imax] - st;
    }
    int tot_part;
    const int ex = block_excl_scan_i(len, S.red, tot_part);
    if (row < nrows) {
      S.seg_start[row] = st;
      S.seg_cum[row] = total + ex;
      S.seg_slot[row] = (unsigned char)slot;
    }
    total += tot_part;
  }
  if (threadIdx.x == 0) S.seg_cum[nrows] = total;
  __syncthreads();
  return total;
}

__device__ __forceinline__ int seg_find(const SearchSmem &S, int nrows, int v) {
  int lo = 0, hi = nrows - 1;   // last segment with cum <= v (skips empty segments)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (S.seg_cum[mid] <= v) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// f(iob, slot) for every candidate of the ScanList, any order (coalesced, no barriers inside)
template <class F>
__device__ __forceinline__ void visit_all(const SearchTables &T, const int *__restrict__ bstart,
                                          const ScanList &SL, SearchSmem &S, F &&f) {
  const int rows = scan_rows_total(SL);
  for (int row0 = 0; row0 < rows; row0 += kSegMax) {
    const int nrows = min(kSegMax, rows - row0);
    const int total = build_segments(T, bstart, SL, row0, nrows, S);
    for (int v = threadIdx.x; v < total; v += blockDim.x) {
      const int sg = seg_find(S, nrows, v);
      f(S.seg_start[sg] + (v - S.seg_cum[sg]), (int)S.seg_slot[sg]);
    }
    __syncthreads();
  }
}

// ---- warp-chunked ordered processing ---------------------------------------------------------------
// A window [v0, v1) of the candidate index space is cut into blockDim/32 contiguous chunks of C
// candidates (C a multiple of 32), chunk w owned by warp w.  fill_window() evaluates every candidate
// once (lanes stride the chunk: coalesced loads, no CTA barrier) and writes the survivors, in scan
// order, to the front of the chunk's slice [w C, w C + wcnt[w]) of the per-CTA candidate buffer.
// Concatenating the slices in warp order therefore IS the reference's visiting order; everything
// downstream (radix select, ordered emission) works slice by slice with one CTA barrier per stage.
template <class E>
__device__ __forceinline__ void fill_window(const SearchSmem &S, int nrows, int v0, int v1, int C,
                                            LocalList &L, int *wcnt, E &&eval) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int beg = v0 + w * C, end = min(v1, beg + C);
  int cnt = 0;
  int sg = (beg + lane < end) ? seg_find(S, nrows, beg + lane) : 0;
  for (int vb = beg; vb < end; vb += 32) {
    const int v = vb + lane;
    bool ok = false;
    double nd = 0.0;
    unsigned pk = 0u;
    if (v < end) {
      while (v >= S.seg_cum[sg + 1]) ++sg;   // v only grows: walk, do not search
      const int iob = S.seg_start[sg] + (v - S.seg_cum[sg]);
      const int slot = (int)S.seg_slot[sg];
      ok = eval(iob, slot, nd);
      pk = ((unsigned)slot << 28) | (unsigned)iob;
    }
    const unsigned m = __ballot_sync(LETKF_FULL_MASK, ok);
    if (ok) {
      const int pos = w * C + cnt + __popc(m & ((1u << lane) - 1u));
      L.cnd[pos] = nd;
      L.cpk[pos] = pk;
    }
    cnt += __popc(m);
  }
  if (lane == 0) wcnt[w] = cnt;
}
// chunk size for a window of n candidates; window capacity for a buffer of ccap entries
__device__ __forceinline__ int chunk_of(int n) {
  const int nw = blockDim.x >> 5;
  return (((n + nw - 1) / nw) + 31) & ~31;
}
__device__ __forceinline__ int window_cap(int ccap) {
  const int u = blockDim.x;   // 32 nw
  const int cap = min(ccap, 128 * u);   // a slice is walked in <= 128 rounds of 32 (per-lane decision bitmask)
  return max((cap / u) * u, 0);
}
// exclusive prefix of this warp's entry and the total of a per-warp count array
__device__ __forceinline__ int warp_prefix(const int *cnt, int &total) {
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int pre = 0, t = 0;
  for (int i = 0; i < nw; ++i) {
    const int c = cnt[i];
    if (i < w) pre += c;
    t += c;
  }
  total = t;
  return pre;
}

// obs_local_cal with the conservative pre-filters in front: identical accept/reject decisions
// and identical ndist as obs_geom + the (optional) search-radius test `ndist > lim2`.
__device__ __forceinline__ bool cand_eval(const SearchTables &T, const CtypeDev &c, const Point &p,
                                          const ObsRec *__restrict__ rec, int iob, double lim2,
                                          double &ndist, double &err) {
  const double2 ve = *reinterpret_cast<const double2 *>(&rec[iob].vc);
  double dv = 0.0;
  if (c.vmode == 1) dv = fabs(__dsub_rn(ve.x, p.lp));
  else if (c.vmode == 2) dv = fabs(__dsub_rn(c.vconst, p.lp));
  else if (c.vmode == 3) dv = fabs(__dsub_rn(ve.x, p.rz));
  if (dv > c.vmax) return false;
  const double v2 = dv * dv * c.iv2;
  if (v2 > lim2 * 1.0000000000009095) return false;   // vertical part alone outside the radius
  const double2 rr = *reinterpret_cast<const double2 *>(&rec[iob].ri);
  const double rdx = __dmul_rn(__dsub_rn(p.ri, rr.x), T.DX);
  const double rdy = __dmul_rn(__dsub_rn(p.rj, rr.y), T.DY);
  const double h2 = __dadd_rn(__dmul_rn(rdx, rdx), __dmul_rn(rdy, rdy));
  if (h2 > c.hmax2) return false;
  if (h2 * c.ih2 + v2 > lim2 * 1.0000000000009095) return false;   // 1 + 2^-40
  // exact path (letkf_tools.f90:1852-1895): same operations, same order, no contraction
  const double nd_v = (c.vmode == 0) ? 0.0 : __ddiv_rn(dv, c.vert_loc);
  if (nd_v > T.dzf) return false;
  const double nd_h = __ddiv_rn(__dsqrt_rn(h2), c.hori_loc);
  if (nd_h > T.dzf) return false;
  ndist = __dadd_rn(__dmul_rn(nd_h, nd_h), __dmul_rn(nd_v, nd_v));
  if (ndist > T.dzf2) return false;
  if (ndist > lim2) return false;
  err = ve.y;
  return true;
}

// ---- obs-number limit, candidates too many for the buffer: radix select by re-scanning ------------
// (the storage-free fallback; same selected set as the buffered path)
__device__ __forceinline__ void select_rescan(const SearchTables &T, const ObsRec *__restrict__ rec,
                                              const int *__restrict__ bstart, const double *__restrict__ vlfac,
                                              const Point &p, const GroupDev &G, const Rect *rq, bool reach_cutoff,
                                              double dcf2, int count, LocalList &L, SearchSmem &S, int &nobsl,
                                              bool &overflow) {
  const int N = G.limit, crit = T.criterion;
  const double tiny = 2.2250738585072014e-308;
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
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int wcap = window_cap(L.ccap);                                // unlimited scans (windowed)
  const int wcap_lim = window_cap(min(L.ccap, T.cand_cap_limit));   // one-window requirement of the limited search
  auto emit_local = [&](int pos, int iob, double vl, double ndist, double err) {
    if (pos < L.cap) {
      const double rl = vl * exp(-0.5 * ndist);
      L.iob[pos] = iob;
      L.rloc[pos] = rl;
      L.rdiag[pos] = err * err / rl;
    }
  };
  // append every buffered survivor (slice by slice = scan order) to the local list
  auto emit_slices = [&](const ScanList &SL, int C) {
    __syncthreads();   // slices and wcnt complete
    int tot;
    const int pre = warp_prefix(S.wcnt, tot);
    const int cw = S.wcnt[w];
    for (int i = lane; i < cw; i += 32) {
      const unsigned pk = L.cpk[w * C + i];
      const int iob = (int)(pk & 0x0fffffffu);
      emit_local(nobsl + pre + i, iob, vlfac[SL.ic[pk >> 28]], L.cnd[w * C + i], rec[iob].err);
    }
    nobsl += tot;
    if (nobsl > L.cap) overflow = true;
    __syncthreads();   // buffer and wcnt free for the next window
  };
  int g = 0;
  while (g < T.ngroup) {
    if (T.grp[g].limit <= 0) {
      // ---- run of groups without obs-number limit (letkf_tools.f90:1438-1476): one fused scan --
      // Entry e of the run = e-th (group, merged ctype) pair; thread e evaluates its cut-off rectangle
      // (double divisions -- not worth repeating in every thread), thread 0 compacts the valid ones.
      int ne = 0, my_ic = -1;
      while (g < T.ngroup && T.grp[g].limit <= 0 && ne + T.grp[g].n <= kMaxScan) {
        const GroupDev &G = T.grp[g];
        for (int icm = 0; icm < G.n; ++icm) {
          if (ne == (int)threadIdx.x) my_ic = G.ic[icm];
          ++ne;
        }
        ++g;
      }
      ScanList &SL = S.sl;
      __syncthreads();   // previous scan finished with S.sl / S.tr
      if (my_ic >= 0) {
        const CtypeDev &c = T.ct[my_ic];
        Rect r = cutoff_rect(T, c, p);
        const bool ok = !(vlfac[my_ic] < tiny) && c.tot != 0 && r.imin <= r.imax && r.jmin <= r.jmax;
        S.tr[threadIdx.x] = r;
        S.tv[threadIdx.x] = ok ? my_ic : -1;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int n = 0;
        for (int e = 0; e < ne; ++e)
          if (S.tv[e] >= 0) {
            SL.ic[n] = S.tv[e];
            SL.r[n] = S.tr[e];
            ++n;
          }
        SL.n = n;
      }
      __syncthreads();
      if (SL.n == 0) continue;
      const int rows = scan_rows_total(SL);
      for (int row0 = 0; row0 < rows; row0 += kSegMax) {
        const int nrows = min(kSegMax, rows - row0);
        const int total = build_segments(T, bstart, SL, row0, nrows, S);
        const int wstep = max(wcap, (int)blockDim.x);   // host guarantees ccap >= 512 >= blockDim
        for (int v0 = 0; v0 < total; v0 += wstep) {
          const int v1 = min(total, v0 + wstep), C = chunk_of(v1 - v0);
          fill_window(S, nrows, v0, v1, C, L, S.wcnt, [&](int iob, int slot, double &nd) {
            double err;
            return cand_eval(T, T.ct[SL.ic[slot]], p, rec, iob, T.dzf2, nd, err);
          });
          emit_slices(SL, C);
        }
      }
      continue;
    }
    // ---- obs-number limit N (letkf_tools.f90:1479-1729) --------------------------------------------
    const GroupDev &G = T.grp[g];
    ++g;
    const int N = G.limit;
    const int crit = T.criterion;
    // criterion 1: incremental square search (letkf_tools.f90:1502-1602); the rectangle only
    // bounds the candidate set, the selected set does not depend on the schedule of q.
    const CtypeDev &cm = T.ct[G.ic[0]];
    double search_incr0 = __ddiv_rn(__dmul_rn(cm.hori_loc, T.dzf), 8.0);
    search_incr0 = fmax(search_incr0, fmax(cm.grdspc_i, cm.grdspc_j));
    ScanList &SL = S.sl;
    bool reach_cutoff = true, buffered = false;
    double dcf2 = T.dzf2;
    int count = 0, C = 32;
    for (int q = (crit == 1 ? 1 : 1 << 20);; ++q) {
      // thread icm: search rectangle of merged ctype icm at step q (S.tr) and whether it already
      // covers the cut-off rectangle (S.tv); thread 0: the scan list of the valid ones
      __syncthreads();   // previous pass finished with S.sl / S.tr
      if ((int)threadIdx.x < G.n) {
        const int icm = threadIdx.x;
        const CtypeDev &c = T.ct[G.ic[icm]];
        const Rect rc = cutoff_rect(T, c, p);
        Rect rr = rc;
        int reach = 1;
        if (crit == 1 && q < (1 << 20)) {
          const double incr = (icm == 0) ? search_incr0 : search_incr0 / cm.hori_loc * c.hori_loc;
          const Rect r = rect_of(T, c, p, incr / T.DX * q, incr / T.DY * q);
          if (!(r.imin <= rc.imin && r.imax >= rc.imax && r.jmin <= rc.jmin && r.jmax >= rc.jmax)) {
            rr = clamp_rect(r, c);
            reach = 0;
          }
        }
        S.tr[icm] = rr;
        S.tv[icm] = reach;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int n = 0, reach = 1;
        for (int icm = 0; icm < G.n; ++icm) {
          reach &= S.tv[icm];
          const int ic = G.ic[icm];
          const Rect &r = S.tr[icm];
          if (vlfac[ic] < tiny || T.ct[ic].tot == 0 || r.imin > r.imax || r.jmin > r.jmax) continue;
          SL.ic[n] = ic;
          SL.r[n] = r;
          ++n;
        }
        SL.n = n;
        S.misc[4] = reach;
      }
      __syncthreads();
      reach_cutoff = S.misc[4] != 0;
      if (!reach_cutoff) {
        const double f = search_incr0 * q / cm.hori_loc;
        dcf2 = f * f;
      } else {
        dcf2 = T.dzf2;
      }
      // valid candidates inside the current radius: counted and -- when the rectangle fits one
      // window of the candidate buffer -- kept, so that the last pass of the loop is also the only
      // evaluation the selection below needs
      const int rows = scan_rows_total(SL);
      buffered = false;
      if (rows <= kSegMax) {
        const int total = (rows > 0) ? build_segments(T, bstart, SL, 0, rows, S) : 0;
        if (total <= wcap_lim) {
          C = chunk_of(total);
          fill_window(S, rows, 0, total, C, L, S.wcnt, [&](int iob, int slot, double &nd) {
            double err;
            return cand_eval(T, T.ct[SL.ic[slot]], p, rec, iob, dcf2, nd, err);
          });
          __syncthreads();
          warp_prefix(S.wcnt, count);
          buffered = true;
        }
      }
      if (!buffered) {
        int cnt = 0;
        visit_all(T, bstart, SL, S, [&](int iob, int slot) {
          double nd, err;
          if (cand_eval(T, T.ct[SL.ic[slot]], p, rec, iob, dcf2, nd, err)) ++cnt;
        });
        count = block_sum_i(cnt, S.red);
      }
      if (reach_cutoff || count >= N) break;
      __syncthreads();   // wcnt / buffer are rewritten by the next pass
    }
    if (count == 0) {
      __syncthreads();
      continue;
    }
    if (!buffered) {
      Rect rq[kMaxMerge];
      for (int icm = 0; icm < G.n; ++icm) rq[icm] = S.tr[icm];
      select_rescan(T, rec, bstart, vlfac, p, G, rq, reach_cutoff, dcf2, count, L, S, nobsl, overflow);
      continue;
    }
    if (count <= N) {   // everything inside the radius is selected
      emit_slices(SL, C);
      continue;
    }
    // ---- radix-select the N-th key among the buffered candidates (each warp reads its own slice) --
    const int cw = S.wcnt[w], sbase = w * C;
    auto buf_key = [&](int i) -> unsigned long long {
      const double nd = L.cnd[i];
      if (crit == 1) return key_bits(nd);
      const unsigned pk = L.cpk[i];
      const double rl = vlfac[SL.ic[pk >> 28]] * exp(-0.5 * nd);
      if (crit == 2) return ~key_bits(rl);
      const double err = rec[pk & 0x0fffffffu].err;
      return key_bits(err * err / rl);
    };
    unsigned long long tau = 0ull;
    int eq_budget = 0, eq_total = 0;
    {
      unsigned long long prefix = 0ull;
      int remaining = N;
      for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) S.hist[i] = 0;
        __syncthreads();
        for (int i = lane; i < cw; i += 32) {
          const unsigned long long key = buf_key(sbase + i);
          const bool match = (pass == 7) || ((key >> (shift + 8)) == (prefix >> (shift + 8)));
          if (match) atomicAdd(&S.hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (threadIdx.x < 32) {   // warp 0: bin holding the `remaining`-th key (8 bins per lane)
          int hv[8], sum = 0;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            hv[t] = S.hist[lane * 8 + t];
            sum += hv[t];
          }
          int inc = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(LETKF_FULL_MASK, inc, o);
            if (lane >= o) inc += y;
          }
          int cum = inc - sum;   // keys in the bins before this lane's
          if (cum < remaining && remaining <= inc) {   // exactly one lane
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              if (cum < remaining && remaining <= cum + hv[t]) {
                S.misc[0] = lane * 8 + t;
                S.misc[1] = cum;
                S.misc[2] = hv[t];
              }
              cum += hv[t];
            }
          }
        }
        __syncthreads();
        prefix |= ((unsigned long long)S.misc[0]) << shift;
        remaining -= S.misc[1];
        eq_total = S.misc[2];
        __syncthreads();
        if (eq_total == remaining && shift > 0) {   // the whole bin is selected: no finer digit needed
          prefix |= (1ull << shift) - 1ull;
          break;
        }
      }
      tau = prefix;          // N-th smallest key (or the top of the last fully selected bin)
      eq_budget = remaining; // how many of the eq_total keys == tau are selected (scan order)
    }
    // ties at the N-th key only partly selected: scan-order rank of every tied candidate
    int eq_pre = 0;
    const bool ties = eq_budget < eq_total;
    if (ties) {
      int ne = 0;
      for (int i0 = 0; i0 < cw; i0 += 32) {
        const int i = i0 + lane;
        ne += __popc(__ballot_sync(LETKF_FULL_MASK, i < cw && buf_key(sbase + i) == tau));
      }
      if (lane == 0) S.wsel[w] = ne;
      __syncthreads();
      int t;
      eq_pre = warp_prefix(S.wsel, t);
      __syncthreads();
    }
    // ordered emission: count the selected per slice, prefix over warps, then write
    unsigned long long keep_lo = 0ull, keep_hi = 0ull;   // this lane's decisions for up to 128 rounds
    int nsel = 0;
    {
      int eq_seen = eq_pre, r = 0;
      for (int i0 = 0; i0 < cw; i0 += 32, ++r) {
        const int i = i0 + lane;
        bool less = false, iseq = false;
        if (i < cw) {
          const unsigned long long key = buf_key(sbase + i);
          less = key < tau;
          iseq = key == tau;
        }
        const unsigned me = __ballot_sync(LETKF_FULL_MASK, iseq);
        const bool take = less || (iseq && (!ties || eq_seen + __popc(me & ((1u << lane) - 1u)) < eq_budget));
        eq_seen += __popc(me);
        const unsigned mt = __ballot_sync(LETKF_FULL_MASK, take);
        if (take) {
          if (r < 64) keep_lo |= 1ull << r; else keep_hi |= 1ull << (r - 64);
        }
        nsel += __popc(mt);
      }
    }
    if (lane == 0) S.wsel[w] = nsel;
    __syncthreads();
    {
      int tot;
      int pos = nobsl + warp_prefix(S.wsel, tot);
      int r = 0;
      for (int i0 = 0; i0 < cw; i0 += 32, ++r) {
        const bool take = (r < 64) ? ((keep_lo >> r) & 1ull) : ((keep_hi >> (r - 64)) & 1ull);
        const unsigned mt = __ballot_sync(LETKF_FULL_MASK, take);
        if (take) {
          const int i = sbase + i0 + lane;
          const unsigned pk = L.cpk[i];
          const int iob = (int)(pk & 0x0fffffffu);
          emit_local(pos + __popc(mt & ((1u << lane) - 1u)), iob, vlfac[SL.ic[pk >> 28]], L.cnd[i], rec[iob].err);
        }
        pos += __popc(mt);
      }
      nobsl += tot;
      if (nobsl > L.cap) overflow = true;
    }
    __syncthreads();
  }
  __syncthreads();
  return overflow ? -1 : nobsl;
}

}  // namespace letkf
