// das_ns_kernel.cuh -- fused per-grid-point LETKF analysis kernel on the FP64 tensor cores (twin of
// the main loop of das_letkf, scale/letkf/letkf_tools.f90:313-686), for MEMBER <= 102.
//
// Persistent CTAs (NB warps, one per 8-row block of the k x k matrices) pull (ij, ilev) points from
// a global counter:
//   relax_beta -> load members (all loads of the point in flight together), form perturbations ->
//   [per variable-localisation group] local observations (pooled list of presearch_kernel, PRE = true, or
//   in-kernel search) -> DMMA Gram [A | b | bd] = Yr^T [Y | dep | depd] from L2-resident obs rows
//   (cp.async, three staging buffers, one barrier per chunk; dep and depd ride in the padding columns)
//   -> interval-scaled coupled Newton-Schulz Z = sqrt(s) A^-1/2 with a third/fourth-order finishing step
//   (ns_solver.cuh) -> one skinny DMMA product Z [dX | b | bd] -> RTPP/RTPS relaxation ->
//   xa = xmean + dX T -> store.
// With t_c = A^-1/2 x_c:   dX W = sqrt(k-1) t_c,   x^T Pa y = t_x . t_y,   dX wbar = t_x . t_b,
// so neither W nor Pa is formed and nothing k x k ever goes to HBM.
#pragma once
#include "das_kernel.cuh"
#include "ns_solver.cuh"

namespace letkf {

// A/B aid: -DLETKF_NS_ZPATH forces the explicit-Z Newton-Schulz (the ill-conditioned path) on every point
#ifdef LETKF_NS_ZPATH
constexpr bool kForceZPath = true;
#else
constexpr bool kForceZPath = false;
#endif

template <int NB, bool PRE = false>
__host__ __device__ inline size_t das_ns_smem_bytes() {
  using C = NsCfg<NB>;
  size_t d = 3 * (size_t)C::PSZ;          // packed Y, Z, T (together: the three obs-chunk staging buffers of the Gram)
  d += (size_t)kMaxNV * C::LD;            // Xall
  if (kMaxNV * C::LD > C::PSZ) d += (size_t)kMaxNV * C::LD;   // Ts (else it aliases T: the Newton-Schulz scratch is free by then)
  d += 4 * (size_t)C::CR;                 // per-row weights of the four staged chunks
  d += 8 * kMaxNV + 56;                   // per-column scalars, reductions (3 NW + 1 <= 49 doubles)
  return d * sizeof(double) + (PRE ? 0 : sizeof(SearchSmem)) + 64;
}

// cp.async (LDGSTS) helpers of the tiled large-ensemble GEMM (tiled.cuh)
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
// the same with a 32-bit shared address (no generic -> shared conversion in the request path of the Gram)
__device__ __forceinline__ void cp_async16_s(unsigned s, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async4_s(unsigned s, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8_s(unsigned s, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_mbar_arrive_s(unsigned b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(b) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ---- mbarrier + bulk-copy (TMA) primitives of the Gram staging pipeline ---------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
  } while (!ok);
}
// one contiguous row, global -> shared, completion (bytes) signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
// the executing thread arrives on the mbarrier once all its earlier cp.async copies have landed (the pending count
// must already include this arrival: mbarrier.init with one count per thread)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long *b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// PRE = true: every local list comes from presearch_kernel's pool; the kernel contains no search code at
// all (fewer live registers, no SearchSmem).  A point whose list did not fit the pool is appended to
// P.redo_list untouched and analysed afterwards by the PRE = false instantiation (P.point_list mode).

template <int NB, bool PRE>
__global__ void __launch_bounds__(NsCfg<NB>::NT, NsCfg<NB>::MINB)
#ifdef LETKF_MAXNREG13
__maxnreg__(NB == 13 ? LETKF_MAXNREG13 : (NB == 9 ? 112 : (NB == 7 ? 72 : (NB == 5 ? 72 : 80))))
#endif
das_ns_kernel(const DasParams P) {
#ifdef LETKF_EXP_TRACE
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    g_trace_buf = P.trace;
    g_trace_n = 0;
  }
#endif
  using C = NsCfg<NB>;
  constexpr int KP = C::KP, LD = C::LD, H = C::H, CR = C::CR, PSZ = C::PSZ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k, nens = P.nens;
  const int tid = threadIdx.x, lane = tid & 31;
  const int wid = __shfl_sync(LETKF_FULL_MASK, tid >> 5, 0);   // warp of the CTA (warp-uniform: tile addressing on the uniform datapath)
  const int w = wid < NB ? wid : NB - 1;                       // its row block
  const TileOwn own = tile_own<NB>(wid);                       // the tiles of that row block it owns (ns_solver.cuh)
  const bool rowown = carries_vectors<NB>(wid);                // it also carries the vector rows of the block
  constexpr int NW = C::NW;
  double *Yp = reinterpret_cast<double *>(smem_raw);
  double *Zp = Yp + PSZ;
  double *Tp = Zp + PSZ;
  double *stage = Yp;                           // 4 x CR x LD doubles inside Yp..Tp
  double *Xall = Tp + PSZ;                      // [kMaxNV][LD]: perturbations of variable vv; rows 14/15: b, bd
  constexpr bool TS_ALIAS = kMaxNV * LD <= PSZ;   // [kMaxNV][LD]: Z Xall -- lives in T's storage when it fits
  double *Ts = TS_ALIAS ? Tp : Xall + (size_t)kMaxNV * LD;
  double *wv = Xall + (size_t)(TS_ALIAS ? 1 : 2) * kMaxNV * LD;      // [4][CR]
  double *colsc = wv + 4 * CR;                  // [8][kMaxNV]
  double *red = colsc + 8 * kMaxNV;
  SearchSmem &S = *reinterpret_cast<SearchSmem *>((reinterpret_cast<uintptr_t>(red + 56) + 15) & ~(uintptr_t)15);
  __shared__ long long s_work, s_next;
  __shared__ long long s_ploff[kMaxNV];
  __shared__ int s_pln[kMaxNV];
  __shared__ __align__(8) unsigned long long s_full[4];   // mbarriers of the four Gram staging buffers
  __shared__ int s_idx[C::NW][2][4];                      // sorted-obs indices of the rows a warp requests next (two chunks)
  const LaneFrag lf = lane_frag(lane);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], C::NT);   // every thread arrives once its copies of the chunk have landed
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned gchunk = 0;   // chunks staged so far by this CTA: buffer = gchunk % 4, mbarrier phase = (gchunk / 4) & 1
  if (P.stagger_ns > 0) {   // experiment: de-phase the CTAs that share an SM (they all start a launch in the same phase)
    const long long t_end = clock64() + 2ll * (long long)(blockIdx.x / P.stagger_div) * P.stagger_ns;   // ~2 clocks per ns
    while (clock64() < t_end) __nanosleep(1000);
  }

  LocalList L;
  L.cap = P.lcap;
  L.iob = P.l_iob + (size_t)blockIdx.x * P.lcap;
  L.rdiag = P.l_rdiag + (size_t)blockIdx.x * P.lcap;
  L.rloc = P.l_rloc + (size_t)blockIdx.x * P.lcap;
  L.ccap = P.ccap;
  L.cnd = P.l_cnd + (size_t)blockIdx.x * P.ccap;
  L.cpk = P.l_cpk + (size_t)blockIdx.x * P.ccap;

  const size_t sl = (size_t)P.nij1 * P.nlev;
  // Statistics and phase clocks live in shared memory and are kept by thread 0 only: as per-thread registers they
  // cost 30 registers that the software-pipelined tensor-core loops need (the kernel runs at 72 registers, 4 CTAs/SM).
  __shared__ unsigned long long s_stat[7];   // points, solved, fail, nobs, overflow, solver iterations, refined solves
  __shared__ long long s_ph[9];              // [0..7] phase clocks, [8] time stamp of the last phase change
  if (tid == 0) {
    for (int i = 0; i < 7; ++i) s_stat[i] = 0;
    for (int i = 0; i < 8; ++i) s_ph[i] = 0;
    s_ph[8] = clock64();
  }
  auto phase = [&](int i) {
    if (tid == 0) {
      const long long now = clock64();
      s_ph[i] += now - s_ph[8];
      s_ph[8] = now;
    }
  };
  auto stat = [&](int i, unsigned long long v) {
    if (tid == 0) s_stat[i] += v;
  };
  double *xm = colsc, *xdet = colsc + kMaxNV, *varg = colsc + 2 * kMaxNV, *vara = colsc + 3 * kMaxNV;
  double *ssum = colsc + 4 * kMaxNV, *sdsum = colsc + 5 * kMaxNV, *inflv = colsc + 6 * kMaxNV;
  double *parmv = colsc + 7 * kMaxNV;
  double *bvec = Xall + (size_t)(kMaxNV - 2) * LD, *bdvec = Xall + (size_t)(kMaxNV - 1) * LD;

  // The work counter is fetched one point ahead (thread 0): the atomic's round trip overlaps the previous point.
  unsigned long long nx_raw = 0;
  if (tid == 0) nx_raw = atomicAdd(&P.counters[0], 1ull);
  for (;;) {
    fence_proxy_async();   // generic-proxy writes to the staging buffers (matrices, Ts) before the next bulk copies
    __syncthreads();
    if (tid == 0) {
      long long v;
      const long long i = (long long)nx_raw;
      nx_raw = atomicAdd(&P.counters[0], 1ull);   // consumed at the top of the next iteration
      if (!PRE && P.point_list) {   // redo mode: explicit list of points, its length produced on the device
        v = (i < (long long)*P.point_count) ? P.point_list[i] : -2;
      } else {
        v = P.point_begin + i;
        if (v >= P.point_end) v = -2;
        if (PRE && v >= 0) {   // all lists of the point must be in the pool
          bool ok = true;
          for (int vg = 0; vg < P.nvgroup; ++vg) {   // (kept in shared memory: the group loop needs them again)
            const long long e = (v - P.pl_base) * P.nvgroup + vg;
            s_ploff[vg] = P.pl_off[e];
            s_pln[vg] = P.pl_n[e];
            ok = ok && s_ploff[vg] >= 0;
          }
          if (!ok) {
            P.redo_list[atomicAdd(P.redo_count, 1ull)] = v;
            v = -1;
          }
        }
      }
      s_work = v;
    }
    __syncthreads();
    const long long wp = s_work;
    LETKF_TRACE(1);
    if (wp == -2) break;      // no more work
    if (wp < 0) continue;     // handed to the redo pass
    phase(7);
    const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
    stat(0, 1);
    const int nvtot = P.nv3d + (il == 0 ? P.nv2d : 0);
    const size_t pbase = (size_t)ij + (size_t)il * P.nij1;

    // ---- relax_beta (letkf_tools.f90:1911-1948) -----------------------------------------------
    const double ri = P.rig1[ij], rj = P.rjg1[ij], rz = P.hgt1[pbase];
    double beta = 1.0;
    if (P.radar_only && rz > P.zcut) {
      beta = 0.0;
    } else if (P.BOUNDARY_BUFFER_WIDTH > 0.0) {
      const double dist_bdy =
          fmin(fmin(ri - P.IHALO, P.nlon + P.IHALO + 1 - ri) * P.DX,
               fmin(rj - P.JHALO, P.nlat + P.JHALO + 1 - rj) * P.DY) / P.BOUNDARY_BUFFER_WIDTH;
      if (dist_bdy < 1.0) beta = fmax(dist_bdy, 0.0);
    }

    // ---- Gram staging machinery of the point (used by every variable-localisation group) -------------------
    int p_use = 0, nchunks = 0;   // local observations / staging chunks of the current group
    double p3acc = 0.0;
    // Every warp requests ITS rows of a chunk (row = wid + NW * r, r < RPW <= 4): the warp copies one row per step with
    // 16-byte cp.async (LDGSTS; lane l moves bytes 16 l ..), lane r also the row's R^-1 weight.  The lists are padded
    // to a multiple of four entries (weight 0), so every staged row is a list entry: no special cases in here.
    constexpr int RPW = C::RPW;
    static_assert(RPW <= 32, "one row index per lane");
    static_assert(RPW <= 4, "s_idx holds four row indices per warp and chunk");
    int p4 = 0;              // list length rounded up to a multiple of four
    // Everything the request path needs is kept as 32-bit shared addresses / warp-uniform values: the path runs once
    // per chunk in every warp and, at four CTAs per SM, its instruction count weighs as much as the DMMAs of the chunk.
    const unsigned stage_s = smem_u32(stage), wv_s = smem_u32(wv), full_s = smem_u32(&s_full[0]);
    const unsigned idx_s = smem_u32(&s_idx[wid][0][0]);
    // The sorted-obs indices of the rows a warp will request travel global -> shared by cp.async as well (lane r: row
    // wid + NW r of chunk c, slot c & 1), one chunk ahead of the request that reads them: no register holds them across
    // the DMMA loop (it was spilled, which exposed the full load latency twice per chunk).
    auto fetch_idx = [&](int c) {
      const int o0 = c * CR, row = wid + NW * lane;
      if (lane < RPW && row < p4 - o0 && row < CR) cp_async4_s(idx_s + 16u * (unsigned)(c & 1) + 4u * (unsigned)lane, L.iob + o0 + row);
      cp_async_commit();
    };
    auto request = [&](int c) {   // every warp, all lanes
      const unsigned st = __shfl_sync(LETKF_FULL_MASK, (gchunk + (unsigned)c) & 3u, 0);   // (uniform for the compiler too)
      // this lane's 16 bytes of row 0 of the buffer / of observation 0 (rows are KP doubles: P.ldens == KP); made opaque
      // so that the compiler keeps them in registers for the three rows instead of recomputing them from scratch
      unsigned buf_l = stage_s + st * (unsigned)(CR * LD * 8) + 16u * (unsigned)lane;
      unsigned long long ens_l = reinterpret_cast<unsigned long long>(P.ensval) + 16ull * (unsigned)lane;
      asm volatile("" : "+r"(buf_l), "+l"(ens_l));
      const int o0 = c * CR, left = p4 - o0;   // rows of the list from this chunk on (a multiple of four)
      cp_async_wait<0>();   // this thread's index copy (and its row copies of the previous request) have landed
      __syncwarp();
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const int row = wid + NW * r;
        if (row < left && row < CR) {
          const int ib = s_idx[wid][c & 1][r];
          const char *src = reinterpret_cast<const char *>(ens_l + (unsigned long long)ib * (KP * 8));
          const unsigned d = buf_l + (unsigned)row * (unsigned)(LD * 8);
          if (KP / 2 >= 32 || lane < KP / 2) cp_async16_s(d, src);
          if (KP / 2 > 32 && lane < KP / 2 - 32) cp_async16_s(d + 512u, src + 512);
        }
      }
      {
        const int row = wid + NW * lane;   // lane r < RPW copies the R^-1 weight of the warp's r-th row (the list holds 1 / rdiag)
        if (lane < RPW && row < left && row < CR) cp_async8_s(wv_s + 8u * (st * (unsigned)CR + (unsigned)row), L.rdiag + o0 + row);
      }
      cp_async_mbar_arrive_s(full_s + 8u * st);
      __syncwarp();   // every lane has read the slot before fetch_idx overwrites it
    };
    auto gram_prologue = [&]() {   // chunks 0 and 1; the indices of chunk 2 are on their way
      fetch_idx(0);
      fetch_idx(1);
      if (nchunks > 0) request(0);
      if (nchunks > 1) request(1);
      fetch_idx(2);
    };
    // With pre-searched lists the first group's observation rows are requested right away: their flight overlaps the
    // member loads below.  (presearch_kernel stores an empty list for a group the solver skips: no stray request.)
    bool pro_done = false;
    if constexpr (PRE) {
      const long long pl_off = s_ploff[0];
      L.iob = P.pl_iob + pl_off;
      L.rdiag = P.pl_rdiag + pl_off;
      L.rloc = P.pl_rloc + pl_off;
      p_use = max(s_pln[0], 0);
      p4 = (p_use + 3) & ~3;
      nchunks = (p_use + CR - 1) / CR;
      if (p_use > 0) {
        gram_prologue();
        pro_done = true;
      }
    }

    // ---- load members, form perturbations (letkf_tools.f90:209-230), destroy gues ----------
    auto gaddr = [&](int vv, int m) -> size_t {   // m 0-based slot
      return (vv < P.nv3d) ? pbase + ((size_t)m + (size_t)vv * nens) * sl
                           : (size_t)ij + ((size_t)m + (size_t)(vv - P.nv3d) * nens) * P.nij1;
    };
    // all member values of the point are requested before anything is consumed: one DRAM round trip instead of
    // one per loop iteration (the write-back to gues3d would otherwise order the loads)
    constexpr int NLD = (kMaxNV * LD + C::NT - 1) / C::NT;
    double raw[NLD];
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
      const int idx = tid + it * C::NT, vv = idx / LD, m = idx - vv * LD;
      raw[it] = 0.0;
      if (idx < kMaxNV * LD && vv < nvtot && m < k) raw[it] = ((vv < P.nv3d) ? P.gues3d : P.gues2d)[gaddr(vv, m)];
    }
    if (tid < nvtot) {
      const double *src = (tid < P.nv3d) ? P.gues3d : P.gues2d;
      xm[tid] = src[gaddr(tid, k)];
      xdet[tid] = P.det ? src[gaddr(tid, k + 1)] : 0.0;
      double infl = P.INFL_MUL;
      if (P.infl_from_field && tid < P.nv3d) infl = P.infl3d[pbase + (size_t)tid * sl];
      if (P.INFL_MUL_MIN > 0.0) infl = fmax(infl, P.INFL_MUL_MIN);
      inflv[tid] = infl;
      parmv[tid] = P.RELAX_TO_INFLATED_PRIOR ? infl : 1.0;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
      const int idx = tid + it * C::NT, vv = idx / LD, m = idx - vv * LD;
      if (idx >= kMaxNV * LD) break;
      double pert = 0.0;
      if (vv < nvtot && m < k) {
        pert = raw[it] - xm[vv];
        ((vv < P.nv3d) ? P.gues3d : P.gues2d)[gaddr(vv, m)] = pert;
      }
      Xall[idx] = pert;
    }
    __syncthreads();
    phase(0);
    LETKF_TRACE(2);
    auto store_anal = [&](int vv, int m, double v) {
      double *dst = (vv < P.nv3d) ? P.anal3d : P.anal2d;
      dst[gaddr(vv, m)] = v;
    };

    if (beta == 0.0) {   // (letkf_tools.f90:333-359)
      for (int idx = tid; idx < nvtot * k; idx += blockDim.x) {
        const int vv = idx / k, m = idx - vv * k;
        store_anal(vv, m, xm[vv] + Xall[(size_t)vv * LD + m]);
      }
      if (P.det && tid < nvtot) store_anal(tid, k + 1, xdet[tid]);
      continue;
    }
    const double pmean = xm[P.iv3d_p - 1];
    Point pt;
    pt.ri = ri;
    pt.rj = rj;
    pt.rz = rz;
    pt.lp = P.logp ? P.logp[pbase] : log(pmean);
    bool solved_any = false;

    for (int vg = 0; vg < P.nvgroup; ++vg) {
      // bit vv: variable vv belongs to this group (masks prepared by the host) ...
      const unsigned gm = P.gmask[vg] & ((1u << nvtot) - 1u);
      const unsigned skip = (pmean < P.Q_UPDATE_TOP) ? (gm & P.qmask) : 0u;   // ... moisture above Q_UPDATE_TOP: left alone
      const unsigned colmask = gm & ~skip;                                     // ... and is analysed
      for (unsigned rest = skip; rest; rest &= rest - 1) {   // (letkf_tools.f90:371-385)
        const int vv = __ffs(rest) - 1;
        for (int m = tid; m < k; m += blockDim.x) store_anal(vv, m, xm[vv] + Xall[(size_t)vv * LD + m]);
        if (P.det && tid == 0) store_anal(vv, k + 1, xdet[vv]);
        if (P.infl3d && tid == 0 && vv < P.nv3d) P.infl3d[pbase + (size_t)vv * sl] = inflv[vv];
      }
      if (colmask == 0) continue;
      const int nc = __popc(colmask);
      const int vtrig = __ffs(colmask) - 1;
      const double infl = inflv[vtrig];   // parm_infl handed to letkf_core (work3d(ij,ilev,n))

      // ---- local observations: pre-searched list (presearch_kernel) or in-kernel search ----------
      // (L's pointers are simply re-aimed: no extra live registers in the Gram loop)
      int nobsl;
      if constexpr (PRE) {
        const long long pl_off = s_ploff[vg];
        nobsl = s_pln[vg];
        L.iob = P.pl_iob + pl_off;
        L.rdiag = P.pl_rdiag + pl_off;
        L.rloc = P.pl_rloc + pl_off;   // only dereferenced when INFL_MUL_ADAPTIVE (then the pool exists)
      } else {
        nobsl = search_point(*P.T, P.rec, P.bstart, P.vlfac + (size_t)vg * P.T->nctype, pt, L, S);
        // the Gram wants what presearch_kernel leaves in the pool: 1 / rdiag, padded to a multiple of four entries
        // with weight 0 (the padding rows repeat the first observation: finite data)
        const int n = max(nobsl, 0), n4 = (n + 3) & ~3;   // (lcap is a multiple of four)
        for (int i = tid; i < n4; i += C::NT) {
          if (i < n) {
            L.rdiag[i] = 1.0 / L.rdiag[i];
          } else {
            L.rdiag[i] = 0.0;
            L.iob[i] = L.iob[0];
          }
        }
        __syncthreads();
      }
      if (nobsl < 0) stat(4, 1);
      if (!pro_done) {
        p_use = nobsl < 0 ? 0 : nobsl;
        p4 = (p_use + 3) & ~3;
        nchunks = (p_use + CR - 1) / CR;
      }
      p3acc = 0.0;
      if (P.INFL_MUL_ADAPTIVE)   // sum of the localisation weights (common_letkf.f90:233)
        for (int i = tid; i < p_use; i += C::NT) p3acc += L.rloc[i];
      phase(1);
      LETKF_TRACE(3);
      if (P.nobsl_out && vg == 0 && tid == 0) P.nobsl_out[pbase] = p_use;
      stat(3, (unsigned long long)p_use);
      bool fail = false;
      bool mean_by_x = false;   // rows 14/15 of Ts hold w = (A/s)^-1 b (refined) instead of Z b: dX wbar = (x . w) / s
      double wscale = sqrt(infl);   // (W dx)_m = wscale * Ts[c][m];  p == 0: W = sqrt(infl) I
      double pscale = infl / (double)(k - 1);   // x^T Pa y = pscale * (t_x . t_y)

      if (p_use > 0) {
        solved_any = true;
        // ---- Gram [A | b | bd] = Yr^T [Y | dep | depd] (common_letkf.f90:111-128,182-195) on the
        // tensor cores: raw obs rows [y_1..y_k, dep, depd, 0..] stream in by cp.async, the R^-1 weight is applied
        // to the A operand.
        double acc[H + 1][2];
#pragma unroll
        for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
        // Staging pipeline: FOUR chunk buffers of CR obs rows (Y, Z and T are all free while the Gram accumulates in
        // registers).  Every warp requests its own rows of a chunk (16-byte cp.async; per-row TMA bulk copies cost
        // ~500 clocks of issue each at four CTAs per SM -- profiles/r02_trace_notes.md), and every thread arrives on
        // the chunk's `full` mbarrier when its copies have landed (cp.async.mbarrier.arrive).  A warp that has finished
        // chunk c requests its rows of chunk c + 2 into buffer (c + 2) % 4: that buffer held chunk c - 2, and passing the
        // wait on chunk c proves that every warp had finished chunk c - 2 (it arrived on chunk c's barrier only after
        // that).  So there is no release barrier, no counter and no CTA-wide barrier in the loop.  The sorted-obs index
        // of the row a lane requests next is loaded one chunk ahead, so a request pays the row latency only.
        if (!pro_done) gram_prologue();
        pro_done = false;
        for (int c = 0; c < nchunks; ++c) {
          const unsigned g = __shfl_sync(LETKF_FULL_MASK, gchunk + (unsigned)c, 0), st = g & 3u;
          LETKF_TRACE(10);
          mbar_wait(&s_full[st], (g >> 2) & 1u);
          LETKF_TRACE(11);
          const int nrows4 = min(CR, p4 - c * CR);
          if (nrows4 == CR && owns_row<NB>(own))
            gram_circ_full<NB, LD, C::NSTEP>(acc, stage + (size_t)st * CR * LD, wv + st * CR, w, lane);
          else gram_circ<NB, LD>(acc, stage + (size_t)st * CR * LD, wv + st * CR, nrows4, w, lane, own);
          LETKF_TRACE(12);
          if (c + 2 < nchunks) request(c + 2);
          LETKF_TRACE(13);
          fetch_idx(c + 3);
          LETKF_TRACE(14);
        }
        gchunk += (unsigned)nchunks;
        __syncthreads();   // the staging buffers (Y among them) are free again
        const double cdiag = (double)(k - 1) / infl;   // (common_letkf.f90:140-143)
        // ---- epilogue of the Gram in registers: ||G||_F, trace, b, bd (and the adaptive-inflation statistics) ----
        // lambda_max(A) = c0 + lambda_max(G) <= c0 + ||G||_F (G = Yr^T Y is positive semi-definite, so the Frobenius
        // norm is never above the trace and it is tight when the spectrum of G decays quickly -- the correlated-
        // observation case that costs Newton-Schulz iterations)
        double fs = 0.0, tr = 0.0;
        {
          const int r = lane >> 2, q = lane & 3, row = w * 8 + r;
#pragma unroll
          for (int d = 0; d <= H; ++d) {
            if (!has<NB>(own, d)) continue;
            const bool dgt = dact<NB>(own, d) == 0;   // the diagonal tile of the row block
            int jb = w + dact<NB>(own, d);
            if (jb >= NB) jb -= NB;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = jb * 8 + 2 * q + e;
              const double v = acc[d][e];
              if (row < k && col < k) {
                fs = fma(dgt ? v : 2.0 * v, v, fs);
                if (dgt && row == col) tr += v;
              }
              // dep / depd ride in columns k, k + 1; a block pair is stored once, as (w, jb) or as its mirror
              if (row < k) {
                if (col == k) bvec[row] = v;
                if (col == k + 1 && P.det) bdvec[row] = v;
              }
              if (!dgt && col < k) {
                if (row == k) bvec[col] = v;
                if (row == k + 1 && P.det) bdvec[col] = v;
              }
              if (dgt && row == k && col == k) red[3 * NW] = v;   // sum w dep^2
            }
          }
          fs = warp_sum(fs);
          tr = warp_sum(tr);
          if (P.INFL_MUL_ADAPTIVE) p3acc = warp_sum(p3acc);
          if (lane == 0) {
            red[wid] = fs;
            red[NW + wid] = tr;
            red[2 * NW + wid] = p3acc;
          }
        }
        __syncthreads();
        double gf2 = 0.0, trace = 0.0, parm3 = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          gf2 += red[i];
          trace += red[NW + i];
          parm3 += red[2 * NW + i];
        }
        const double s_norm = cdiag + sqrt(gf2) * (1.0 + 1.0e-12);   // (rounding guard)
        if (P.INFL_MUL_ADAPTIVE) {   // (common_letkf.f90:229-254)
          const double parm1 = red[3 * NW];
          const double parm2 = trace / (double)(k - 1);
          const double parm4 = (parm1 - parm3) / parm2 - infl;
          const double tq = (infl * parm2 + parm3) / parm2;
          const double sigma_o = 2.0 / parm3 * (tq * tq);
          const double gain = 0.04 * 0.04 / (sigma_o + 0.04 * 0.04);
          if (tid == 0) inflv[vtrig] = infl + gain * parm4;
        }
        // ---- M0 = (G + c0 I) / s on the leading k x k block, identity on the padding ------------------
        {
          const double is = 1.0 / s_norm;
          const int r = lane >> 2, q = lane & 3, row = w * 8 + r;
#pragma unroll
          for (int d = 0; d <= H; ++d) {
            int jb = w + dact<NB>(own, d);
            if (jb >= NB) jb -= NB;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = jb * 8 + 2 * q + e;
              const bool dg = (dact<NB>(own, d) == 0 && row == col);
              acc[d][e] = (row < k && col < k) ? (acc[d][e] + (dg ? cdiag : 0.0)) * is : (dg ? 1.0 : 0.0);
            }
          }
          store_circ<NB>(acc, Yp, w, lf, own);
        }
        // Ill-conditioned point (lambda_max(A) / c0 above ~1e4): the mean weight wbar = Pa b is refined below against
        // the ORIGINAL matrix, which is kept in global scratch for that purpose (rare: nothing is written otherwise).
        const bool refine = (cdiag / s_norm) < 1.0e-4;
        if (refine) store_circ<NB>(acc, P.m0_scratch + (size_t)blockIdx.x * PSZ, w, lf, own);
        if (tid == 0) {   // the next point of this CTA (its work counter was drawn at the top of this iteration)
          long long nv = -1;
          if (PRE || !P.point_list) {
            nv = P.point_begin + (long long)nx_raw;
            if (nv >= P.point_end) nv = -1;
          }
          s_next = nv;
        }
        phase(2);
        LETKF_TRACE(4);
        // ---- Ts = (A/s)^-1/2 [dX | b | bd] ------------------------------------------------------------
        // Normal path: the Newton-Schulz factors are applied to the 16 vectors directly (newton_schulz_apply), Z is
        // never formed.  Ill-conditioned points need Z itself for the refinement of the mean weight below.
        int its;
        if (!kForceZPath && !refine) {
          its = newton_schulz_apply<NB, LD>(acc, Yp, Zp, Xall, Ts, TS_ALIAS ? Zp : Tp, Yp, cdiag / s_norm, red,
                                            P.max_sweeps + 20);
        } else {
          stat(6, 1);
          its = newton_schulz_invsqrt<NB>(acc, Yp, Zp, Tp, cdiag / s_norm, red, P.max_sweeps + 20);
        }
        if (its < 0) fail = true;
        stat(5, (unsigned long long)(its < 0 ? -its : its));
        // mtx_eigen zeroes eigenvalues below lambda_max*sqrt(eps) (common_mtx.f90:69) and letkf_core
        // would then divide by zero; ||A||_F <= sqrt(k) lambda_max bounds the same condition.
        if (!(cdiag * (double)k >= s_norm * 1.4901161193847656e-08)) fail = true;
        phase(4);
        LETKF_TRACE(5);
        {   // L2 prefetch of the next point's members (and its list header): their DRAM + TLB latency overlaps the
            // rest of this point instead of opening the next one (s_next was published before the solve's barriers)
          const long long nv = s_next;
          if (nv >= 0) {
            const int nil = (int)(nv / P.nij1), nij = (int)(nv - (long long)nil * P.nij1);
            const size_t nbase = (size_t)nij + (size_t)nil * P.nij1;
            for (int idx = tid; idx < P.nv3d * (k + 1 + P.det); idx += C::NT) {
              const int vv = idx / (k + 1 + P.det), m = idx - vv * (k + 1 + P.det);
              const double *a = P.gues3d + nbase + ((size_t)m + (size_t)vv * nens) * sl;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
            if (PRE && tid == 0) {
              const long long e = (nv - P.pl_base) * P.nvgroup;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(P.pl_off + e));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(P.pl_n + e));
            }
          }
        }
        // ---- refine path: Ts = Z [dX | b | bd]  (k x 16 skinny product on the tensor cores) ----------
        if ((kForceZPath || refine) && rowown) {
          double a2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
          const int r = lane >> 2, q = lane & 3;
          const double *pb = Xall + (size_t)r * LD + q;
#pragma unroll
          for (int e = 0; e < NB; ++e) {
            int j = w + e;
            if (j >= NB) j -= NB;
            double a[2];
            symm_afrag<NB>(a, Zp, w, e, j * C::RS, lf);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int kk = j * 8 + h * 4;
              dmma884(a2[0][0], a2[0][1], a[h], pb[kk]);
              dmma884(a2[1][0], a2[1][1], a[h], pb[(size_t)8 * LD + kk]);
            }
          }
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) Ts[(size_t)(nt * 8 + 2 * q + e) * LD + w * 8 + r] = a2[nt][e];
        }
        __syncthreads();
        if (refine) {
          // One step of iterative refinement of w = (A/s)^-1 b (and of the deterministic-run twin):
          //   w0 = Z (Z b),  r = b - M0 w0,  w = w0 + Z (Z r),   then  dX wbar = (x . w) / s.
          // Z carries a normwise error ~ eps sqrt(cond); b = Yr^T d is huge in the directions where A^-1 is tiny, so the
          // product form (Z x).(Z b) loses ~10x against the eigen-decomposition of the reference above cond 1e6;
          // the residual against the original matrix restores it (tools/ns_model.py, tests/test_gpu_illcond.py).
          double *M0s = Yp;                                   // free since the last product of the solve
          double *vw = Tp + (TS_ALIAS ? kMaxNV * LD : 0);     // [4][LD] vector scratch behind Ts
          const double *m0g = P.m0_scratch + (size_t)blockIdx.x * PSZ;
          for (int i = tid; i < PSZ; i += C::NT) M0s[i] = m0g[i];
          const int nvec = P.det ? 2 : 1;
          auto matvec = [&](const double *Mat, const double *x, double *y) {   // y = Mat x, leading k x k block
            for (int i = tid; i < k; i += C::NT) {
              double sacc = 0.0;
              for (int j = 0; j < k; ++j) sacc = fma(Mat[caddr<NB>(i, j)], x[j], sacc);
              y[i] = sacc;
            }
          };
          __syncthreads();
          for (int iv = 0; iv < nvec; ++iv) {
            const double *bv = Xall + (size_t)(kMaxNV - 2 + iv) * LD;   // b (bd)
            double *tbv = Ts + (size_t)(kMaxNV - 2 + iv) * LD;           // Z b -> (refined) w
            double *w0 = vw, *rr = vw + LD, *t1 = vw + 2 * LD;
            matvec(Zp, tbv, w0);
            __syncthreads();
            matvec(M0s, w0, rr);
            __syncthreads();
            for (int i = tid; i < k; i += C::NT) rr[i] = bv[i] - rr[i];
            __syncthreads();
            matvec(Zp, rr, t1);
            __syncthreads();
            matvec(Zp, t1, rr);
            __syncthreads();
            for (int i = tid; i < k; i += C::NT) tbv[i] = w0[i] + rr[i];
            __syncthreads();
          }
        }
        mean_by_x = refine;
        wscale = sqrt((double)(k - 1) / s_norm);
        pscale = 1.0 / s_norm;
      } else {
        // nobsl == 0 (common_letkf.f90:89-107): W = sqrt(infl) I, wbar = 0, Pa = infl/(k-1) I
        for (int idx = tid; idx < kMaxNV * LD; idx += blockDim.x) {
          const int vv = idx / LD;
          Ts[idx] = (vv < kMaxNV - 2) ? Xall[idx] : 0.0;
        }
        __syncthreads();
      }
      if (fail) stat(2, 1);
      phase(5);
      LETKF_TRACE(6);
      // ---- one HALF-warp per column: var_g = x.x, var_a = x^T Pa x, s = x^T Pa b, sd = x^T Pa bd, RTPP/RTPS
      // relaxation, update (letkf_tools.f90:457-513), q-spread clamp and the store -- no CTA barrier, everything a
      // column needs after the skinny product is local to its 16 lanes (2 NB half-warps >= the 11 columns: one round)
      {
        const double *tb = Ts + (size_t)(kMaxNV - 2) * LD, *tbd = Ts + (size_t)(kMaxNV - 1) * LD;
        constexpr int MPL = (KP + 15) / 16;   // members per lane
        const int hl = lane & 15;
        auto hsum = [&](double v) {   // sum over the 16 lanes of the half-warp
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(LETKF_FULL_MASK, v, o);
          return v;
        };
        for (int c0 = 2 * wid; c0 < nc; c0 += 2 * NW) {
          const int c = c0 + (lane >> 4);
          const bool active = c < nc;
          const int vv = active ? (int)__fns(colmask, 0, c + 1) : vtrig;   // c-th analysed variable of the group
          const double *x = Xall + (size_t)vv * LD, *t = Ts + (size_t)vv * LD;
          double xv[MPL], tv[MPL], xa[MPL];
          double vg_ = 0.0, va_ = 0.0, s_ = 0.0, sdv_ = 0.0;
#pragma unroll
          for (int u = 0; u < MPL; ++u) {
            const int m = hl + 16 * u;
            xv[u] = tv[u] = 0.0;
            if (m < k) {
              xv[u] = x[m];
              tv[u] = t[m];
              vg_ = fma(xv[u], xv[u], vg_);
              va_ = fma(tv[u], tv[u], va_);
              s_ = fma(mean_by_x ? xv[u] : tv[u], tb[m], s_);
              sdv_ = fma(mean_by_x ? xv[u] : tv[u], tbd[m], sdv_);
            }
          }
          vg_ = hsum(vg_);
          va_ = hsum(va_);
          s_ = hsum(s_);
          if (P.det) sdv_ = hsum(sdv_);
          // RTPS factor of the column (letkf_tools.f90:1971-2002), once per variable instead of per member
          const double va = va_ * pscale, parm = parmv[vv], xmv = xm[vv], ss = s_ * pscale;
          double f = 1.0;
          if (P.RELAX_ALPHA == 0.0 && P.RELAX_ALPHA_SPREAD != 0.0 && vg_ > 0.0 && va > 0.0)
            f = P.RELAX_ALPHA_SPREAD * sqrt(vg_ * parm / (va * (double)(k - 1))) - P.RELAX_ALPHA_SPREAD + 1.0;
          const double rpp = (P.RELAX_ALPHA != 0.0) ? P.RELAX_ALPHA * sqrt(parm) : 0.0;
#pragma unroll
          for (int u = 0; u < MPL; ++u) {
            const double z = wscale * tv[u];   // (W dx)_m
            double wx;                         // (W_rlx dx)_m
            if (P.RELAX_ALPHA != 0.0) wx = (1.0 - P.RELAX_ALPHA) * z + rpp * xv[u];
            else wx = f * z;
            xa[u] = xmv + (wx + ss) * beta + (1.0 - beta) * xv[u];
          }
          if (P.Q_SPRD_MAX > 0.0) {   // (letkf_tools.f90:500-513); the shuffles are executed by every lane
            double part = 0.0;
#pragma unroll
            for (int u = 0; u < MPL; ++u)
              if (hl + 16 * u < k) part += xa[u];
            const double q_mean = hsum(part) / (double)k;
            part = 0.0;
#pragma unroll
            for (int u = 0; u < MPL; ++u)
              if (hl + 16 * u < k) part = fma(xa[u] - q_mean, xa[u] - q_mean, part);
            const double q_sprd = sqrt(hsum(part) / (double)(k - 1)) / q_mean;
            if (vv == P.iv3d_q - 1 && q_sprd > P.Q_SPRD_MAX) {
#pragma unroll
              for (int u = 0; u < MPL; ++u) xa[u] = q_mean + (xa[u] - q_mean) * P.Q_SPRD_MAX / q_sprd;
            }
          }
          if (active) {
#pragma unroll
            for (int u = 0; u < MPL; ++u)
              if (hl + 16 * u < k) store_anal(vv, hl + 16 * u, xa[u]);
            if (hl == 0) {
              if (P.det) store_anal(vv, k + 1, xdet[vv] + sdv_ * pscale * beta);   // (:489-497)
              if (P.rtps_out && vv < P.nv3d) P.rtps_out[pbase + (size_t)vv * sl] = f;
              if (P.infl3d && vv < P.nv3d) {
                const double v = (vv == vtrig || P.INFL_MUL_ADAPTIVE) ? inflv[P.INFL_MUL_ADAPTIVE ? P.vfirst[vv] : vv]
                                                                       : inflv[vv];
                P.infl3d[pbase + (size_t)vv * sl] = v;
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncthreads();
      phase(6);
      LETKF_TRACE(7);
    }
    if (solved_any) stat(1, 1);
  }
  if (tid == 0) {
    for (int i = 0; i < 7; ++i) atomicAdd(&P.counters[1 + i], s_stat[i]);
    for (int i = 0; i < 8; ++i) atomicAdd(&P.counters[8 + i], (unsigned long long)s_ph[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// presearch_kernel: the local-observation search of das_letkf for the points [point_begin, point_end),
// run AHEAD of the solver on its own stream so that its latency-bound work overlaps the solver's
// tensor-core work on the same SMs.  Lists go to a pool (atomic bump allocation); a point that does
// not fit keeps pl_off = -1 and is searched by the solver itself.  Same point / group activity rules
// as das_ns_kernel (relax_beta, Q_UPDATE_TOP mask), same Point inputs, hence bit-identical lists.
__global__ void __launch_bounds__(128) presearch_kernel(const DasParams P, int *pl_n, long long *pl_off) {
  __shared__ SearchSmem S;
  __shared__ long long s_work;
  __shared__ long long s_off;
  const int tid = threadIdx.x, k = P.k;
  LocalList L;
  L.cap = P.lcap;
  L.iob = P.l_iob + (size_t)blockIdx.x * P.lcap;
  L.rdiag = P.l_rdiag + (size_t)blockIdx.x * P.lcap;
  L.rloc = P.l_rloc + (size_t)blockIdx.x * P.lcap;
  L.ccap = P.ccap;
  L.cnd = P.l_cnd + (size_t)blockIdx.x * P.ccap;
  L.cpk = P.l_cpk + (size_t)blockIdx.x * P.ccap;
  const size_t sl = (size_t)P.nij1 * P.nlev;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = P.point_begin + (long long)atomicAdd(&P.counters[0], 1ull);
    __syncthreads();
    const long long wp = s_work;
    if (wp >= P.point_end) break;
    const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
    const int nvtot = P.nv3d + (il == 0 ? P.nv2d : 0);
    const size_t pbase = (size_t)ij + (size_t)il * P.nij1;
    const double ri = P.rig1[ij], rj = P.rjg1[ij], rz = P.hgt1[pbase];
    double beta = 1.0;
    if (P.radar_only && rz > P.zcut) {
      beta = 0.0;
    } else if (P.BOUNDARY_BUFFER_WIDTH > 0.0) {
      const double dist_bdy =
          fmin(fmin(ri - P.IHALO, P.nlon + P.IHALO + 1 - ri) * P.DX,
               fmin(rj - P.JHALO, P.nlat + P.JHALO + 1 - rj) * P.DY) / P.BOUNDARY_BUFFER_WIDTH;
      if (dist_bdy < 1.0) beta = fmax(dist_bdy, 0.0);
    }
    const double pmean = P.gues3d[pbase + ((size_t)k + (size_t)(P.iv3d_p - 1) * P.nens) * sl];   // mean slot
    Point pt;
    pt.ri = ri;
    pt.rj = rj;
    pt.rz = rz;
    pt.lp = P.logp ? P.logp[pbase] : log(pmean);
    for (int vg = 0; vg < P.nvgroup; ++vg) {
      const long long e = (wp - P.pl_base) * P.nvgroup + vg;
      bool any = false;
      if (beta != 0.0) {
        for (int vv = 0; vv < nvtot; ++vv) {
          if (P.vgroup[vv] != vg) continue;
          const bool masked = (vv < P.nv3d) && pmean < P.Q_UPDATE_TOP && (vv + 1) >= P.iv3d_q && (vv + 1) <= P.iv3d_qg;
          if (!masked) any = true;
        }
      }
      if (!any) {   // the solver never asks for this list
        if (tid == 0) {
          pl_n[e] = 0;
          pl_off[e] = 0;
        }
        continue;
      }
      const int n = search_point(*P.T, P.rec, P.bstart, P.vlfac + (size_t)vg * P.T->nctype, pt, L, S);
      if (tid == 0) {
        long long off = 0;
        if (n > 0) {
          const int n4 = (n + 3) & ~3;   // keep every list 32-byte aligned for the solver's staging loads
          off = (long long)atomicAdd(P.pl_cursor, (unsigned long long)n4);
          if (off + n4 > P.pl_cap) off = -1;
        }
        s_off = off;
        pl_n[e] = n;
        pl_off[e] = off;
      }
      __syncthreads();
      const long long off = s_off;
      if (n > 0 && off >= 0) {
        // what the solver's Gram stages: 1 / rdiag, the list padded to four entries with weight 0 (the padding
        // rows repeat the first observation: finite data)
        const int n4 = (n + 3) & ~3;
        for (int i = tid; i < n4; i += blockDim.x) {
          P.pl_iob[off + i] = L.iob[i < n ? i : 0];
          P.pl_rdiag[off + i] = i < n ? 1.0 / L.rdiag[i] : 0.0;
          if (P.pl_rloc) P.pl_rloc[off + i] = i < n ? L.rloc[i] : 0.0;
        }
      }
    }
  }
}

}  // namespace letkf
