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

template <int NB, bool PRE = false>
__host__ __device__ inline size_t das_ns_smem_bytes() {
  using C = NsCfg<NB>;
  size_t d = 3 * (size_t)C::PSZ;          // packed Y, Z, T (together: the three obs-chunk staging buffers of the Gram)
  d += (size_t)kMaxNV * C::LD;            // Xall
  if (kMaxNV * C::LD > C::PSZ) d += (size_t)kMaxNV * C::LD;   // Ts (else it aliases T: the Newton-Schulz scratch is free by then)
  d += 3 * (size_t)C::CR;                 // per-row weights of the three staged chunks
  d += 8 * kMaxNV + 40;                   // per-column scalars, reductions
  return d * sizeof(double) + (PRE ? 0 : sizeof(SearchSmem)) + 64;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// PRE = true: every local list comes from presearch_kernel's pool; the kernel contains no search code at
// all (fewer live registers, no SearchSmem).  A point whose list did not fit the pool is appended to
// P.redo_list untouched and analysed afterwards by the PRE = false instantiation (P.point_list mode).
template <int NB, bool PRE>
__global__ void __launch_bounds__(NsCfg<NB>::NT, NsCfg<NB>::MINB)
das_ns_kernel(const DasParams P) {
  using C = NsCfg<NB>;
  constexpr int KP = C::KP, LD = C::LD, H = C::H, CR = C::CR, PSZ = C::PSZ;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k, nens = P.nens;
  const int tid = threadIdx.x, lane = tid & 31;
  const int w = __shfl_sync(LETKF_FULL_MASK, tid >> 5, 0);   // warp-uniform: tile addressing on the uniform datapath
  double *Yp = reinterpret_cast<double *>(smem_raw);
  double *Zp = Yp + PSZ;
  double *Tp = Zp + PSZ;
  double *stage = Yp;                           // 3 x CR x LD doubles inside Yp..Tp
  double *Xall = Tp + PSZ;                      // [kMaxNV][LD]: perturbations of variable vv; rows 14/15: b, bd
  constexpr bool TS_ALIAS = kMaxNV * LD <= PSZ;   // [kMaxNV][LD]: Z Xall -- lives in T's storage when it fits
  double *Ts = TS_ALIAS ? Tp : Xall + (size_t)kMaxNV * LD;
  double *wv = Xall + (size_t)(TS_ALIAS ? 1 : 2) * kMaxNV * LD;      // [2][CR]
  double *colsc = wv + 3 * CR;                  // [8][kMaxNV]
  double *red = colsc + 8 * kMaxNV;
  SearchSmem &S = *reinterpret_cast<SearchSmem *>((reinterpret_cast<uintptr_t>(red + 40) + 15) & ~(uintptr_t)15);
  __shared__ long long s_work;
  __shared__ long long s_ploff[kMaxNV];
  __shared__ int s_pln[kMaxNV];
  const LaneFrag lf = lane_frag(lane);

  LocalList L;
  L.cap = P.lcap;
  L.iob = P.l_iob + (size_t)blockIdx.x * P.lcap;
  L.rdiag = P.l_rdiag + (size_t)blockIdx.x * P.lcap;
  L.rloc = P.l_rloc + (size_t)blockIdx.x * P.lcap;
  L.ccap = P.ccap;
  L.cnd = P.l_cnd + (size_t)blockIdx.x * P.ccap;
  L.cpk = P.l_cpk + (size_t)blockIdx.x * P.ccap;

  const size_t sl = (size_t)P.nij1 * P.nlev;
  unsigned long long c_points = 0, c_solved = 0, c_fail = 0, c_nobs = 0, c_over = 0, c_iters = 0;
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tph = clock64();
  auto phase = [&](int i) {
    const long long now = clock64();
    ph[i] += now - tph;
    tph = now;
  };
  double *xm = colsc, *xdet = colsc + kMaxNV, *varg = colsc + 2 * kMaxNV, *vara = colsc + 3 * kMaxNV;
  double *ssum = colsc + 4 * kMaxNV, *sdsum = colsc + 5 * kMaxNV, *inflv = colsc + 6 * kMaxNV;
  double *parmv = colsc + 7 * kMaxNV;
  double *bvec = Xall + (size_t)(kMaxNV - 2) * LD, *bdvec = Xall + (size_t)(kMaxNV - 1) * LD;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      long long v;
      if (!PRE && P.point_list) {   // redo mode: explicit list of points, its length produced on the device
        const long long i = (long long)atomicAdd(&P.counters[0], 1ull);
        v = (i < (long long)*P.point_count) ? P.point_list[i] : -2;
      } else {
        v = P.point_begin + (long long)atomicAdd(&P.counters[0], 1ull);
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
    if (wp == -2) break;      // no more work
    if (wp < 0) continue;     // handed to the redo pass
    phase(7);
    const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
    ++c_points;
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
      int cols[kMaxNV];
      int nc = 0;
      for (int vv = 0; vv < nvtot; ++vv) {
        if (P.vgroup[vv] != vg) continue;
        const bool masked = (vv < P.nv3d) && pmean < P.Q_UPDATE_TOP && (vv + 1) >= P.iv3d_q &&
                            (vv + 1) <= P.iv3d_qg;
        if (masked) {   // (letkf_tools.f90:371-385)
          for (int m = tid; m < k; m += blockDim.x) store_anal(vv, m, xm[vv] + Xall[(size_t)vv * LD + m]);
          if (P.det && tid == 0) store_anal(vv, k + 1, xdet[vv]);
          if (P.infl3d && tid == 0 && vv < P.nv3d) P.infl3d[pbase + (size_t)vv * sl] = inflv[vv];
        } else {
          cols[nc++] = vv;
        }
      }
      if (nc == 0) continue;
      const int vtrig = cols[0];
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
      }
      if (nobsl < 0) ++c_over;
      const int p_use = nobsl < 0 ? 0 : nobsl;
      phase(1);
      if (P.nobsl_out && vg == 0 && tid == 0) P.nobsl_out[pbase] = p_use;
      c_nobs += (unsigned long long)p_use;
      bool fail = false;
      double wscale = sqrt(infl);   // (W dx)_m = wscale * Ts[c][m];  p == 0: W = sqrt(infl) I
      double pscale = infl / (double)(k - 1);   // x^T Pa y = pscale * (t_x . t_y)

      if (p_use > 0) {
        solved_any = true;
        // ---- Gram [A | b | bd] = Yr^T [Y | dep | depd] (common_letkf.f90:111-128,182-195) on the
        // tensor cores: raw obs rows [y_1..y_k, dep, depd, 0..] stream in with cp.async, the
        // R^-1 weight is applied to the A operand.
        double acc[H + 1][2];
#pragma unroll
        for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
        double p3acc = 0.0;
        const int nchunks = (p_use + CR - 1) / CR;
        static_assert(CR <= 4 * NB, "a warp stages at most four rows of a chunk");
        // One warp per obs row (a row is KP/2 16-byte pieces), rows w, w + NB, w + 2 NB, w + 3 NB of a chunk.
        // Chunks are requested strictly in order, and the sorted-obs indices of the NEXT chunk are fetched
        // while the current one is requested, so that a request only pays the row latency, not index + row.
        // The same holds for the R^-1 weights: the thread that owns row `tid` of a chunk loads rdiag one chunk
        // ahead, otherwise its warp would sit out an L2 round trip inside every request while the other
        // warps wait for it at the chunk barrier.
        int nxt[4];
        double rd_nxt = 0.0, rl_nxt = 0.0;
        auto fetch_idx = [&](int c) {
          const int o0 = c * CR, nrows = min(CR, p_use - o0);
#pragma unroll
          for (int u = 0; u < 4; ++u) nxt[u] = (w + u * NB < nrows) ? L.iob[o0 + w + u * NB] : -1;
          if (tid < nrows) {
            rd_nxt = L.rdiag[o0 + tid];
            if (P.INFL_MUL_ADAPTIVE) rl_nxt = L.rloc[o0 + tid];
          }
        };
        fetch_idx(0);
        auto issue = [&](int c) {
          double *dst = stage + (size_t)(c % 3) * CR * LD;
          double *wdst = wv + (c % 3) * CR;
          const int o0 = c * CR;
          const int nrows = min(CR, p_use - o0), nrows4 = (nrows + 3) & ~3;
          const double rd_cur = rd_nxt, rl_cur = rl_nxt;
          {
            int iobs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) iobs[u] = nxt[u];
            fetch_idx(c + 1);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (iobs[u] < 0) continue;
              const double *src = P.ensval + (size_t)iobs[u] * P.ldens;
              double *drow = dst + (size_t)(w + u * NB) * LD;
              for (int pc = lane; pc < KP / 2; pc += 32) cp_async16(drow + 2 * pc, src + 2 * pc);
            }
          }
          for (int idx = tid; idx < (nrows4 - nrows) * KP; idx += blockDim.x)
            dst[(size_t)(nrows + idx / KP) * LD + idx % KP] = 0.0;
          if (tid < nrows4) {
            double wt = 0.0;
            if (tid < nrows) {
              wt = 1.0 / rd_cur;
              if (P.INFL_MUL_ADAPTIVE) p3acc += rl_cur;
            }
            wdst[tid] = wt;
          }
          cp_async_commit();
        };
        // three staging buffers (Y, Z and T are all free while the Gram accumulates in registers): chunk c+2 is
        // requested while chunk c is consumed, and ONE barrier per chunk both publishes chunk c and retires
        // the buffer of chunk c-1 that the next request overwrites
        issue(0);
        if (nchunks > 1) issue(1);
        for (int c = 0; c < nchunks; ++c) {
          if (c + 1 < nchunks) cp_async_wait<1>();
          else cp_async_wait<0>();
          __syncthreads();
          if (c + 2 < nchunks) issue(c + 2);
          const int nrows = min(CR, p_use - c * CR);
          gram_circ<NB, LD>(acc, stage + (size_t)(c % 3) * CR * LD, wv + (c % 3) * CR, (nrows + 3) & ~3, w, lane);
        }
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
            int jb = w + d;
            if (jb >= NB) jb -= NB;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = jb * 8 + 2 * q + e;
              const double v = acc[d][e];
              if (row < k && col < k) {
                fs = fma(d == 0 ? v : 2.0 * v, v, fs);
                if (d == 0 && row == col) tr += v;
              }
              // dep / depd ride in columns k, k + 1; a block pair is stored once, as (w, jb) or as its mirror
              if (row < k) {
                if (col == k) bvec[row] = v;
                if (col == k + 1 && P.det) bdvec[row] = v;
              }
              if (d != 0 && col < k) {
                if (row == k) bvec[col] = v;
                if (row == k + 1 && P.det) bdvec[col] = v;
              }
              if (d == 0 && row == k && col == k) red[3 * NB] = v;   // sum w dep^2
            }
          }
          fs = warp_sum(fs);
          tr = warp_sum(tr);
          p3acc = warp_sum(p3acc);
          if (lane == 0) {
            red[w] = fs;
            red[NB + w] = tr;
            red[2 * NB + w] = p3acc;
          }
        }
        __syncthreads();
        double gf2 = 0.0, trace = 0.0, parm3 = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          gf2 += red[i];
          trace += red[NB + i];
          parm3 += red[2 * NB + i];
        }
        const double s_norm = cdiag + sqrt(gf2) * (1.0 + 1.0e-12);   // (rounding guard)
        if (P.INFL_MUL_ADAPTIVE) {   // (common_letkf.f90:229-254)
          const double parm1 = red[3 * NB];
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
            int jb = w + d;
            if (jb >= NB) jb -= NB;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = jb * 8 + 2 * q + e;
              const bool dg = (d == 0 && row == col);
              acc[d][e] = (row < k && col < k) ? (acc[d][e] + (dg ? cdiag : 0.0)) * is : (dg ? 1.0 : 0.0);
            }
          }
          store_circ<NB>(acc, Yp, w, lf);
        }
        phase(2);
        // ---- Z = (A/s)^-1/2 --------------------------------------------------------------------------
        const int its = newton_schulz_invsqrt<NB>(acc, Yp, Zp, Tp, cdiag / s_norm, red, P.max_sweeps + 20);
        if (its < 0) fail = true;
        c_iters += (unsigned long long)(its < 0 ? -its : its);
        // mtx_eigen zeroes eigenvalues below lambda_max*sqrt(eps) (common_mtx.f90:69) and letkf_core
        // would then divide by zero; ||A||_F <= sqrt(k) lambda_max bounds the same condition.
        if (!(cdiag * (double)k >= s_norm * 1.4901161193847656e-08)) fail = true;
        phase(4);
        // ---- Ts = Z [dX | b | bd]  (k x 16 skinny product on the tensor cores) --------------------
        {
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
      // ---- per-column scalars: var_g = x.x, var_a = x^T Pa x, s = x^T Pa b, sd = x^T Pa bd ----------
      {
        const int nw = blockDim.x >> 5;
        const double *tb = Ts + (size_t)(kMaxNV - 2) * LD, *tbd = Ts + (size_t)(kMaxNV - 1) * LD;
        for (int c = w; c < nc; c += nw) {
          const int vv = cols[c];
          const double *x = Xall + (size_t)vv * LD, *t = Ts + (size_t)vv * LD;
          double vg_ = 0.0, va_ = 0.0, s_ = 0.0, sdv_ = 0.0;
          for (int m = lane; m < k; m += 32) {
            const double xv = x[m], tv = t[m];
            vg_ = fma(xv, xv, vg_);
            va_ = fma(tv, tv, va_);
            s_ = fma(tv, tb[m], s_);
            sdv_ = fma(tv, tbd[m], sdv_);
          }
          vg_ = warp_sum(vg_);
          va_ = warp_sum(va_);
          s_ = warp_sum(s_);
          sdv_ = warp_sum(sdv_);
          if (lane == 0) {
            // RTPS factor of the column (letkf_tools.f90:1971-2002), once per variable instead of per member
            const double va = va_ * pscale;
            double f = 1.0;
            if (P.RELAX_ALPHA == 0.0 && P.RELAX_ALPHA_SPREAD != 0.0 && vg_ > 0.0 && va > 0.0)
              f = P.RELAX_ALPHA_SPREAD * sqrt(vg_ * parmv[vv] / (va * (double)(k - 1))) - P.RELAX_ALPHA_SPREAD + 1.0;
            varg[c] = vg_;
            vara[c] = f;
            ssum[c] = s_ * pscale;
            sdsum[c] = P.det ? sdv_ * pscale : 0.0;
          }
        }
      }
      __syncthreads();
      if (fail) ++c_fail;
      phase(5);

      // ---- relaxation + update (letkf_tools.f90:457-513); result staged in Ts -------------------
      for (int idx = tid; idx < nc * k; idx += blockDim.x) {
        const int c = idx / k, m = idx - c * k;
        const int vv = cols[c];
        const double x = Xall[(size_t)vv * LD + m];
        const double z = wscale * Ts[(size_t)vv * LD + m];   // (W dx)_m
        const double parm = parmv[vv];
        double wx;   // (W_rlx dx)_m
        if (P.RELAX_ALPHA != 0.0) {
          wx = (1.0 - P.RELAX_ALPHA) * z + P.RELAX_ALPHA * sqrt(parm) * x;
        } else if (P.RELAX_ALPHA_SPREAD != 0.0) {
          wx = vara[c] * z;
        } else {
          wx = z;
        }
        Ts[(size_t)vv * LD + m] = xm[vv] + (wx + ssum[c]) * beta + (1.0 - beta) * x;
      }
      __syncthreads();
      if (P.Q_SPRD_MAX > 0.0) {   // (letkf_tools.f90:500-513)
        for (int c = 0; c < nc; ++c) {
          if (cols[c] != P.iv3d_q - 1) continue;
          double *tq = Ts + (size_t)cols[c] * LD;
          double part = 0.0;
          for (int m = tid; m < k; m += blockDim.x) part += tq[m];
          const double q_mean = block_sum(part, red) / (double)k;
          part = 0.0;
          for (int m = tid; m < k; m += blockDim.x) {
            const double d = tq[m] - q_mean;
            part = fma(d, d, part);
          }
          const double q_sprd = sqrt(block_sum(part, red) / (double)(k - 1)) / q_mean;
          if (q_sprd > P.Q_SPRD_MAX) {
            for (int m = tid; m < k; m += blockDim.x) {
              const double d = tq[m] - q_mean;
              tq[m] = q_mean + d * P.Q_SPRD_MAX / q_sprd;
            }
          }
          __syncthreads();
        }
      }
      for (int idx = tid; idx < nc * k; idx += blockDim.x) {
        const int c = idx / k, m = idx - c * k;
        store_anal(cols[c], m, Ts[(size_t)cols[c] * LD + m]);
      }
      if (tid < nc) {
        const int vv = cols[tid];
        if (P.det) store_anal(vv, k + 1, xdet[vv] + sdsum[tid] * beta);   // (:489-497)
        if (P.rtps_out && vv < P.nv3d) {
          P.rtps_out[pbase + (size_t)vv * sl] = vara[tid];
        }
        if (P.infl3d && vv < P.nv3d) {
          const double v = (vv == vtrig || P.INFL_MUL_ADAPTIVE) ? inflv[P.INFL_MUL_ADAPTIVE ? P.vfirst[vv] : vv]
                                                                 : inflv[vv];
          P.infl3d[pbase + (size_t)vv * sl] = v;
        }
      }
      __syncthreads();
      phase(6);
    }
    if (solved_any) ++c_solved;
  }
  if (tid == 0) {
    atomicAdd(&P.counters[1], c_points);
    atomicAdd(&P.counters[2], c_solved);
    atomicAdd(&P.counters[3], c_fail);
    atomicAdd(&P.counters[4], c_nobs);
    atomicAdd(&P.counters[5], c_over);
    atomicAdd(&P.counters[6], c_iters);
    for (int i = 0; i < 8; ++i) atomicAdd(&P.counters[8 + i], (unsigned long long)ph[i]);
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
        for (int i = tid; i < n; i += blockDim.x) {
          P.pl_iob[off + i] = L.iob[i];
          P.pl_rdiag[off + i] = L.rdiag[i];
          if (P.pl_rloc) P.pl_rloc[off + i] = L.rloc[i];
        }
      }
    }
  }
}

}  // namespace letkf
