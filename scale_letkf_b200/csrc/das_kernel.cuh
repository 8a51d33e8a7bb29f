// das_kernel.cuh -- fused per-grid-point LETKF analysis kernel (twin of the main loop of
// das_letkf, scale/letkf/letkf_tools.f90:313-686) and the batched letkf_core twin.
//
// Persistent CTAs (one resident set per SM) pull (ij, ilev) points from a global counter:
//   relax_beta -> load members, form perturbations -> [per variable-localisation group]
//   local-obs search -> SYRK Gram from L2-resident obs rows -> Cholesky -> one-sided Jacobi
//   -> G^T x / G u products -> RTPP/RTPS relaxation -> xa = xmean + dX * T  -> store.
// The k x k matrices W and Pa are never formed: with A = sum_j g_j g_j^T,
//   dX W    = G (D2 (G^T dx)),     D2 = sqrt(k-1) / lambda^1.5
//   x^T Pa y = sum_j (G^T x)_j (G^T y)_j / lambda_j^2
// so the update costs O(k^2) per variable and nothing k x k ever goes to HBM.
#pragma once
#include "search.cuh"
#include "solver.cuh"

namespace letkf {

constexpr int kMaxNV = 16;

struct DasParams {
  // sizes
  int k, nens, nij1, nlev, nv3d, nv2d, det;
  int ld, ldk, npairs, ncols;
  // state (reference layout, Fortran order)
  double *gues3d, *anal3d, *gues2d, *anal2d;
  double *infl3d, *rtps_out;
  int *nobsl_out;
  const double *logp;
  const double *rig1, *rjg1, *hgt1;
  // observations
  const SearchTables *T;
  const ObsRec *rec;
  const int *bstart;
  const double *ensval;   // [nobstotal][ldens], member fastest
  const double *val;      // [nobstotal]
  int ldens;
  // variable-localisation groups (letkf_tools.f90:130-163)
  int nvgroup;
  int vgroup[kMaxNV];     // group of variable vv (0-based)
  int vfirst[kMaxNV];     // var_local_n2n - 1: first variable of vv's group
  unsigned gmask[kMaxNV]; // bit vv: variable vv belongs to group vg (das_ns_kernel: no per-point loop over the variables)
  unsigned qmask;         // bit vv: 3-D moisture variable iv3d_q .. iv3d_qg (left alone above Q_UPDATE_TOP)
  const double *vlfac;    // [nvgroup][nctype]
  // namelist scalars
  double INFL_MUL, INFL_MUL_MIN, RELAX_ALPHA, RELAX_ALPHA_SPREAD, Q_UPDATE_TOP, Q_SPRD_MAX;
  int RELAX_TO_INFLATED_PRIOR, INFL_MUL_ADAPTIVE, infl_from_field;
  int iv3d_p, iv3d_q, iv3d_qg;
  // relax_beta (letkf_tools.f90:1911-1948)
  int radar_only;
  double zcut, BOUNDARY_BUFFER_WIDTH, DX, DY;
  int IHALO, JHALO, nlon, nlat;
  // per-CTA scratch and counters
  int *l_iob;
  double *l_rdiag, *l_rloc;
  int lcap;
  double *l_cnd;          // [grid][ccap] candidate buffers of the obs-number-limited search
  unsigned *l_cpk;
  int ccap;
  unsigned long long *counters;   // [0] work, [1] npoints, [2] nsolved, [3] nfail, [4] nobsl_sum, [5] overflow, [6] Jacobi sweeps, [8..15] phase clocks
  long long point_begin, point_end;   // (ij, ilev) points [begin, end) of this launch, ilev-major
  int max_sweeps;
  int stagger_ns, stagger_div;   // experiments (LETKF_B200_STAGGER_US): CTA b sleeps (b / stagger_div) * stagger_ns at kernel start
  // Local lists produced ahead of time by presearch_kernel (das_ns_kernel.cuh) for the points
  // [pl_base, ...): entry (wp - pl_base) * nvgroup + vg holds the count (pl_n, -1 = list overflow) and
  // the offset into the pools (pl_off, -1 = not pre-searched: the solver searches by itself).
  const int *pl_n;
  const long long *pl_off;
  int *pl_iob;
  double *pl_rdiag, *pl_rloc;
  long long pl_base, pl_cap;
  unsigned long long *pl_cursor;
  // points the PRE solver could not take (list not in the pool), and the redo pass over them
  long long *redo_list;
  unsigned long long *redo_count;
  const long long *point_list;
  const unsigned long long *point_count;
  double *m0_scratch;   // [grid][PSZ]: M0 = A / s of ill-conditioned points, kept for the mean-weight refinement
  long long *trace;   // LETKF_EXP_TRACE builds only: (tag, clock64) pairs of CTA 0 / thread 0
};

__host__ __device__ inline size_t das_smem_bytes(int k, int nthreads) {
  const int ld = ld_of(k), ldk = ldk_of(k), ncols = 2 * ((k + 1) / 2);
  size_t d = 0;
  d += (size_t)ld * ncols;                 // G
  size_t ys = (size_t)kChunk * ldk;        // Ys chunk (aliased by Ts, Zs after the Gram)
  const size_t tz = 2 * (size_t)ncols * kMaxNV;
  if (ys < tz) ys = tz;
  d += ys;
  d += (size_t)round_up(k, 2) * kMaxNV;    // Xs
  d += ncols;                              // lam
  d += 3 * kChunk;                         // sw, sd, sdd
  d += 8 * kMaxNV;                         // per-column scalars
  d += kMaxWarps + 8;                      // reductions, pivot
  (void)nthreads;
  return d * sizeof(double) + sizeof(SearchSmem) + 64;
}

// Stage `nrows` local observations [o0, o0+nrows) of the CTA's list into Ys (scaled by
// sqrt(1/rdiag)), and sd/sdd = sqrt(w) * dep / depd.
__device__ __forceinline__ void stage_chunk(const DasParams &P, const LocalList &L, int o0, int nrows,
                                            double *Ys, double *sw, double *sd, double *sdd) {
  const int k = P.k, ldk = P.ldk;
  if (threadIdx.x < nrows) {
    const int o = o0 + threadIdx.x;
    const int iob = L.iob[o];
    const double w = sqrt(1.0 / L.rdiag[o]);
    sw[threadIdx.x] = w;
    sd[threadIdx.x] = w * P.val[iob];
    sdd[threadIdx.x] = P.det ? w * P.ensval[(size_t)iob * P.ldens + k + 1] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < nrows * ldk; idx += blockDim.x) {
    const int o = idx / ldk, a = idx - o * ldk;
    double v = 0.0;
    if (a < k) v = P.ensval[(size_t)L.iob[o0 + o] * P.ldens + a] * sw[o];
    Ys[idx] = v;
  }
  __syncthreads();
}

template <int KC>
__global__ void __launch_bounds__(SizeClass<KC>::NT)
das_kernel(const DasParams P) {
  using SC = SizeClass<KC>;
  constexpr int R = SC::R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k, ld = P.ld, ldk = P.ldk, ncols = P.ncols, nens = P.nens;
  const int tid = threadIdx.x;
  double *G = reinterpret_cast<double *>(smem_raw);
  size_t ys_sz = (size_t)kChunk * ldk;
  if (ys_sz < 2 * (size_t)ncols * kMaxNV) ys_sz = 2 * (size_t)ncols * kMaxNV;
  double *Ys = G + (size_t)ld * ncols;
  double *Ts = Ys;                              // alias: valid after the Gram phase
  double *Zs = Ys + (size_t)ncols * kMaxNV;
  double *Xs = Ys + ys_sz;
  double *lam = Xs + (size_t)round_up(k, 2) * kMaxNV;
  double *sw = lam + ncols;
  double *sd = sw + kChunk;
  double *sdd = sd + kChunk;
  double *colsc = sdd + kChunk;                 // [8][kMaxNV]: xm, xdet, var_g, var_a, s, sdt, infl, parm
  double *red = colsc + 8 * kMaxNV;
  double *piv = red + kMaxWarps;
  SearchSmem &S = *reinterpret_cast<SearchSmem *>(
      (reinterpret_cast<uintptr_t>(piv + 8) + 15) & ~(uintptr_t)15);
  __shared__ long long s_work;
  __shared__ int s_flag;

  LocalList L;
  L.cap = P.lcap;
  L.iob = P.l_iob + (size_t)blockIdx.x * P.lcap;
  L.rdiag = P.l_rdiag + (size_t)blockIdx.x * P.lcap;
  L.rloc = P.l_rloc + (size_t)blockIdx.x * P.lcap;
  L.ccap = P.ccap;
  L.cnd = P.l_cnd + (size_t)blockIdx.x * P.ccap;
  L.cpk = P.l_cpk + (size_t)blockIdx.x * P.ccap;

  const size_t sl = (size_t)P.nij1 * P.nlev;
  const int ntiles = ((k + 3) / 4) * ((k + 3) / 4 + 1) / 2;
  unsigned long long c_points = 0, c_solved = 0, c_fail = 0, c_nobs = 0, c_over = 0, c_sweeps = 0;
  // SM-clock cycles of this CTA per phase (thread 0): load, search, gram, cholesky, jacobi, apply, store
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

  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = P.point_begin + (long long)atomicAdd(&P.counters[0], 1ull);
    __syncthreads();
    const long long wp = s_work;
    if (wp >= P.point_end) break;
    phase(7);
    const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
    ++c_points;
    const int nvtot = P.nv3d + (il == 0 ? P.nv2d : 0);
    const size_t pbase = (size_t)ij + (size_t)il * P.nij1;

    // ---- relax_beta ------------------------------------------------------------------------
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
    for (int idx = tid; idx < nvtot * k; idx += blockDim.x) {
      const int vv = idx / k, m = idx - vv * k;
      double *src = (vv < P.nv3d) ? P.gues3d : P.gues2d;
      const size_t ad = gaddr(vv, m);
      const double pert = src[ad] - xm[vv];
      src[ad] = pert;
      Xs[(size_t)m * kMaxNV + vv] = pert;
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
        store_anal(vv, m, xm[vv] + Xs[(size_t)m * kMaxNV + vv]);
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
      // active variables of this group, in ascending order; the first one triggers the solve
      int cols[kMaxNV];
      int nc = 0;
      for (int vv = 0; vv < nvtot; ++vv) {
        if (P.vgroup[vv] != vg) continue;
        const bool masked = (vv < P.nv3d) && pmean < P.Q_UPDATE_TOP && (vv + 1) >= P.iv3d_q &&
                            (vv + 1) <= P.iv3d_qg;
        if (masked) {   // (letkf_tools.f90:371-385)
          for (int m = tid; m < k; m += blockDim.x) store_anal(vv, m, xm[vv] + Xs[(size_t)m * kMaxNV + vv]);
          if (P.det && tid == 0) store_anal(vv, k + 1, xdet[vv]);
          if (P.infl3d && tid == 0 && vv < P.nv3d) P.infl3d[pbase + (size_t)vv * sl] = inflv[vv];
        } else {
          cols[nc++] = vv;
        }
      }
      if (nc == 0) continue;
      const int vtrig = cols[0];
      double infl = inflv[vtrig];   // parm_infl handed to letkf_core (work3d(ij,ilev,n))

      // ---- local observations ---------------------------------------------------------------
      const int nobsl = search_point(*P.T, P.rec, P.bstart, P.vlfac + (size_t)vg * P.T->nctype, pt, L, S);
      if (nobsl < 0) {
        ++c_over;
      }
      const int p_use = nobsl < 0 ? 0 : nobsl;
      phase(1);
      if (P.nobsl_out && vg == 0 && tid == 0) P.nobsl_out[pbase] = p_use;
      c_nobs += (unsigned long long)p_use;
      const int cb = nc, cbd = nc + 1;             // columns of b = Yr^T dep, bd = Yr^T depd
      const int ncx = nc + 1 + (P.det ? 1 : 0);
      const int nvb = (ncx + 3) / 4;
      bool fail = false;

      if (p_use > 0) {
        solved_any = true;
        // ---- Gram A = Yr^T Y (common_letkf.f90:111-128) and b = Yr^T dep ----------------------
        double acc[R][16];
        gram_zero<R>(acc);
        double bacc = 0.0, bdacc = 0.0, tracc = 0.0, p1acc = 0.0, p3acc = 0.0;
        for (int o0 = 0; o0 < p_use; o0 += kChunk) {
          const int nrows = min(kChunk, p_use - o0);
          stage_chunk(P, L, o0, nrows, Ys, sw, sd, sdd);
          gram_accumulate<R>(acc, Ys, nrows, ldk, ntiles);
          if (tid < k) {
            for (int o = 0; o < nrows; ++o) {
              const double y = Ys[o * ldk + tid];
              bacc = fma(y, sd[o], bacc);
              bdacc = fma(y, sdd[o], bdacc);
              tracc = fma(y, y, tracc);
            }
          }
          if (P.INFL_MUL_ADAPTIVE && tid < nrows) {
            p1acc += sd[tid] * sd[tid];
            p3acc += L.rloc[o0 + tid];
          }
          __syncthreads();
        }
        gram_store<R>(acc, G, ld, k, ntiles, false);
        if (tid < k) {
          // column slots of the right-hand sides in Xs (masked variables keep their slots)
          Xs[(size_t)tid * kMaxNV + kMaxNV - 2] = bacc;
          Xs[(size_t)tid * kMaxNV + kMaxNV - 1] = bdacc;
        }
        __syncthreads();
        const double cdiag = (double)(k - 1) / infl;   // (common_letkf.f90:140-143)
        if (tid < k) G[(size_t)tid * ld + tid] += cdiag;
        if (P.INFL_MUL_ADAPTIVE) {   // (common_letkf.f90:229-254)
          const double parm1 = block_sum(p1acc, red);
          const double parm2 = block_sum(tid < k ? tracc : 0.0, red) / (double)(k - 1);
          const double parm3 = block_sum(p3acc, red);
          const double parm4 = (parm1 - parm3) / parm2 - infl;
          const double tq = (infl * parm2 + parm3) / parm2;
          const double sigma_o = 2.0 / parm3 * (tq * tq);
          const double gain = 0.04 * 0.04 / (sigma_o + 0.04 * 0.04);
          if (tid == 0) inflv[vtrig] = infl + gain * parm4;
        }
        __syncthreads();
        phase(2);
        // ---- A = L L^T, one-sided Jacobi on L -> G = U S ------------------------------------------
        const bool ok = cholesky_lower(G, k, ld, ncols, piv);
        phase(3);
        bool conv = false;
        c_sweeps += (unsigned long long)jacobi_onesided<SC::RJ>(G, k, ld, P.npairs, red, P.max_sweeps, cdiag, &conv);
        phase(4);
        column_norms(G, k, ld, ncols, lam);
        if (!ok || !conv) fail = true;
        double lmax = 0.0, lmin = 1.0e300;
        for (int j = 0; j < ncols; ++j)
          if (lam[j] > 0.0) {
            lmax = fmax(lmax, lam[j]);
            lmin = fmin(lmin, lam[j]);
          }
        // mtx_eigen zeroes eigenvalues below lambda_max*sqrt(eps) (common_mtx.f90:69) and
        // letkf_core would then divide by zero: report instead.
        if (!(lmin >= lmax * 1.4901161193847656e-08)) fail = true;
        // ---- gather the group's columns contiguously: Xc[a][c] (aliases Zs region) -----------
        // Xs holds all variables; build the compact right-hand-side block in Zs.
        for (int idx = tid; idx < k * 4 * nvb; idx += blockDim.x) {
          const int a = idx / (4 * nvb), c = idx - a * 4 * nvb;
          double v = 0.0;
          if (c < nc) v = Xs[(size_t)a * kMaxNV + cols[c]];
          else if (c == cb) v = Xs[(size_t)a * kMaxNV + kMaxNV - 2];
          else if (c == cbd && P.det) v = Xs[(size_t)a * kMaxNV + kMaxNV - 1];
          Zs[(size_t)a * kMaxNV + c] = v;
        }
        __syncthreads();
        gemm_gt_x(G, k, ld, ncols, Zs, Ts, nvb, kMaxNV);   // Ts[j][c] = (G^T X)_jc
        __syncthreads();
        // ---- per-column scalars: var_g, var_a = x^T Pa x, s = x^T Pa b, sd = x^T Pa bd --------
        // shifted form: x^T Pa y = x.y / c0 + sum_j (1/lambda_j - 1/c0)/lambda_j (G^T x)_j (G^T y)_j
        {
          const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
          const double ic0 = 1.0 / cdiag;
          for (int c = warp; c < nc; c += nw) {
            double vg_ = 0.0, va_ = 0.0, s_ = 0.0, sdv_ = 0.0, xb_ = 0.0, xbd_ = 0.0;
            for (int a = lane; a < k; a += 32) {
              const double x = Xs[(size_t)a * kMaxNV + cols[c]];
              vg_ = fma(x, x, vg_);
              xb_ = fma(x, Xs[(size_t)a * kMaxNV + kMaxNV - 2], xb_);
              xbd_ = fma(x, Xs[(size_t)a * kMaxNV + kMaxNV - 1], xbd_);
            }
            for (int j = lane; j < ncols; j += 32) {
              const double l = lam[j];
              if (l > 0.0) {
                const double w = (cdiag - l) / (l * cdiag) / l;
                const double t = Ts[(size_t)j * kMaxNV + c];
                va_ = fma(t * t, w, va_);
                s_ = fma(t * Ts[(size_t)j * kMaxNV + cb], w, s_);
                if (P.det) sdv_ = fma(t * Ts[(size_t)j * kMaxNV + cbd], w, sdv_);
              }
            }
            vg_ = warp_sum(vg_);
            va_ = warp_sum(va_);
            s_ = warp_sum(s_);
            sdv_ = warp_sum(sdv_);
            xb_ = warp_sum(xb_);
            xbd_ = warp_sum(xbd_);
            if (lane == 0) {
              varg[c] = vg_;
              vara[c] = fma(vg_, ic0, va_);
              ssum[c] = fma(xb_, ic0, s_);
              sdsum[c] = P.det ? fma(xbd_, ic0, sdv_) : 0.0;
            }
          }
        }
        __syncthreads();
        // U = D2 T for the variable columns, then Z = sqrt(rho) dx + G U = W dx with
        // D2_j = (sqrt((k-1)/lambda_j) - sqrt((k-1)/c0)) / lambda_j
        const double sk1 = sqrt((double)(k - 1));
        const double fw0 = sk1 / sqrt(cdiag);
        for (int idx = tid; idx < ncols * nc; idx += blockDim.x) {
          const int j = idx / nc, c = idx - j * nc;
          const double l = lam[j];
          const double d2 = (l > 0.0) ? (sk1 / sqrt(l) - fw0) / l : 0.0;
          Ts[(size_t)j * kMaxNV + c] *= d2;
        }
        __syncthreads();
        gemm_g_u(G, k, ld, ncols, Ts, Zs, (nc + 3) / 4, kMaxNV);   // Zs[m][c] = ((W - sqrt(rho) I) dx)_m
        __syncthreads();
        for (int idx = tid; idx < k * nc; idx += blockDim.x) {
          const int a = idx / nc, c = idx - a * nc;
          Zs[(size_t)a * kMaxNV + c] = fma(fw0, Xs[(size_t)a * kMaxNV + cols[c]], Zs[(size_t)a * kMaxNV + c]);
        }
        __syncthreads();
      } else {
        // nobsl == 0 (common_letkf.f90:89-107): W = sqrt(infl) I, wbar = 0, Pa = infl/(k-1) I
        const double sq = sqrt(infl);
        const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
        for (int c = warp; c < nc; c += nw) {
          double vg_ = 0.0;
          for (int a = lane; a < k; a += 32) {
            const double x = Xs[(size_t)a * kMaxNV + cols[c]];
            vg_ = fma(x, x, vg_);
          }
          vg_ = warp_sum(vg_);
          if (lane == 0) {
            varg[c] = vg_;
            vara[c] = vg_ * (infl / (double)(k - 1));
            ssum[c] = 0.0;
            sdsum[c] = 0.0;
          }
        }
        for (int idx = tid; idx < k * nc; idx += blockDim.x) {
          const int a = idx / nc, c = idx - a * nc;
          Zs[(size_t)a * kMaxNV + c] = sq * Xs[(size_t)a * kMaxNV + cols[c]];
        }
        __syncthreads();
      }
      if (fail) ++c_fail;
      phase(5);

      // ---- relaxation + update (letkf_tools.f90:457-513) --------------------------------------
      for (int idx = tid; idx < nc * k; idx += blockDim.x) {
        const int c = idx / k, m = idx - c * k;
        const int vv = cols[c];
        const double x = Xs[(size_t)m * kMaxNV + vv];
        const double z = Zs[(size_t)m * kMaxNV + c];
        const double parm = parmv[vv];
        double wx;   // (W_rlx dx)_m
        if (P.RELAX_ALPHA != 0.0) {
          wx = (1.0 - P.RELAX_ALPHA) * z + P.RELAX_ALPHA * sqrt(parm) * x;
        } else if (P.RELAX_ALPHA_SPREAD != 0.0) {
          double f = 1.0;
          if (varg[c] > 0.0 && vara[c] > 0.0)
            f = P.RELAX_ALPHA_SPREAD * sqrt(varg[c] * parm / (vara[c] * (double)(k - 1))) -
                P.RELAX_ALPHA_SPREAD + 1.0;
          wx = f * z;
        } else {
          wx = z;
        }
        const double xa = xm[vv] + (wx + ssum[c]) * beta + (1.0 - beta) * x;
        Ts[(size_t)m * kMaxNV + c] = xa;   // staged for the q-spread clamp
      }
      __syncthreads();
      if (P.Q_SPRD_MAX > 0.0) {   // (letkf_tools.f90:500-513)
        for (int c = 0; c < nc; ++c) {
          if (cols[c] != P.iv3d_q - 1) continue;
          double part = 0.0;
          for (int m = tid; m < k; m += blockDim.x) part += Ts[(size_t)m * kMaxNV + c];
          const double q_mean = block_sum(part, red) / (double)k;
          part = 0.0;
          for (int m = tid; m < k; m += blockDim.x) {
            const double d = Ts[(size_t)m * kMaxNV + c] - q_mean;
            part = fma(d, d, part);
          }
          const double q_sprd = sqrt(block_sum(part, red) / (double)(k - 1)) / q_mean;
          if (q_sprd > P.Q_SPRD_MAX) {
            for (int m = tid; m < k; m += blockDim.x) {
              const double d = Ts[(size_t)m * kMaxNV + c] - q_mean;
              Ts[(size_t)m * kMaxNV + c] = q_mean + d * P.Q_SPRD_MAX / q_sprd;
            }
          }
          __syncthreads();
        }
      }
      for (int idx = tid; idx < nc * k; idx += blockDim.x) {
        const int c = idx / k, m = idx - c * k;
        store_anal(cols[c], m, Ts[(size_t)m * kMaxNV + c]);
      }
      if (tid < nc) {
        const int vv = cols[tid];
        if (P.det) store_anal(vv, k + 1, xdet[vv] + sdsum[tid] * beta);   // (:489-497)
        if (P.rtps_out && vv < P.nv3d) {
          double f = 1.0;
          if (P.RELAX_ALPHA == 0.0 && P.RELAX_ALPHA_SPREAD != 0.0 && varg[tid] > 0.0 && vara[tid] > 0.0)
            f = P.RELAX_ALPHA_SPREAD * sqrt(varg[tid] * parmv[vv] / (vara[tid] * (double)(k - 1))) -
                P.RELAX_ALPHA_SPREAD + 1.0;
          P.rtps_out[pbase + (size_t)vv * sl] = f;
        }
        if (P.infl3d && vv < P.nv3d) {
          // the trigger variable carries the (possibly adapted) value; the others copy it when
          // adaptive (letkf_tools.f90:396-398), else keep their own
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
    atomicAdd(&P.counters[6], c_sweeps);
    for (int i = 0; i < 8; ++i) atomicAdd(&P.counters[8 + i], (unsigned long long)ph[i]);
  }
  (void)s_flag;
}

// ---------------------------------------------------------------------------------------------
// obs_local twin for a batch of points (parity tests of the selection).
struct SearchParams {
  const SearchTables *T;
  const ObsRec *rec;
  const int *bstart;
  const double *vlfac;    // [nctype] for the requested nvar
  const double *ri, *rj, *lp, *rz;
  int npts, max_out;
  int *nobsl, *idx;
  double *rdiag, *rloc;
  int *l_iob;
  double *l_rdiag, *l_rloc;
  int lcap;
  double *l_cnd;
  unsigned *l_cpk;
  int ccap;
  unsigned long long *counters;
};

__global__ void __launch_bounds__(128) search_kernel(const SearchParams P) {
  __shared__ SearchSmem S;
  __shared__ int s_work;
  LocalList L;
  L.cap = P.lcap;
  L.iob = P.l_iob + (size_t)blockIdx.x * P.lcap;
  L.rdiag = P.l_rdiag + (size_t)blockIdx.x * P.lcap;
  L.rloc = P.l_rloc + (size_t)blockIdx.x * P.lcap;
  L.ccap = P.ccap;
  L.cnd = P.l_cnd + (size_t)blockIdx.x * P.ccap;
  L.cpk = P.l_cpk + (size_t)blockIdx.x * P.ccap;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_work = (int)atomicAdd(&P.counters[0], 1ull);
    __syncthreads();
    const int w = s_work;
    if (w >= P.npts) break;
    Point pt;
    pt.ri = P.ri[w];
    pt.rj = P.rj[w];
    pt.lp = P.lp[w];
    pt.rz = P.rz[w];
    const int n = search_point(*P.T, P.rec, P.bstart, P.vlfac, pt, L, S);
    if (threadIdx.x == 0) {
      P.nobsl[w] = n;
      if (n < 0) atomicAdd(&P.counters[5], 1ull);
    }
    if (P.idx && n > 0 && n <= P.max_out) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        P.idx[(size_t)w * P.max_out + i] = L.iob[i];
        if (P.rdiag) P.rdiag[(size_t)w * P.max_out + i] = L.rdiag[i];
        if (P.rloc) P.rloc[(size_t)w * P.max_out + i] = L.rloc[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NOBS_OUT fields of das_letkf (letkf_tools.f90:281-284, 440-447, 767-778) at every analysis point, for the
// variable-localisation group of one model variable: obs_local with its optional outputs nobsl_t / cutd_t
// (:1380-1390, 1427-1431, 1473-1475, 1653-1660, 1718-1729) reduced to the eleven fields the reference writes:
//   out(:,:,0..4) = sum over elements of nobsl_t(:, type) for report types 1, 3, 4, 8, 22
//   out(:,:,5..7) = nobsl_t(REF | RE0 | VR, PHARAD),  out(:,:,8..10) = cutd_t(REF | RE0 | VR, PHARAD)
// The search is search_point (the list das_letkf analyses with); the epilogue classifies the selected observations by
// combined type (ranges of the sorted order) and, for obs-number-limited groups that filled their budget, takes the worst
// criterion key of the selected set (criterion 1: normalised distance recomputed with obs_geom's exact arithmetic).
// Reference quirks kept: nobsl_t of an UNLIMITED merged group is cumulative over its members (nobsl_prev is set once,
// :1444); merged non-master types report cutd_t = 0.
struct NobsOutParams {
  const SearchTables *T;
  const ObsRec *rec;
  const int *bstart;
  const double *vlfac;
  const double *rig1, *rjg1, *hgt1, *logp, *pmean;   // logp (nij1,nlev) or NULL -> log(pmean)
  int nij1, nlev, radar_only, IHALO, JHALO, nlon, nlat;
  double zcut, BOUNDARY_BUFFER_WIDTH, DX, DY;
  double *out;   // (nij1, nlev, 11)
  int *l_iob;
  double *l_rdiag, *l_rloc;
  int lcap;
  double *l_cnd;
  unsigned *l_cpk;
  int ccap;
  unsigned long long *counters;
};

__global__ void __launch_bounds__(128) nobs_out_kernel(const NobsOutParams P) {
  __shared__ SearchSmem S;
  __shared__ long long s_work;
  __shared__ int s_cnt[kMaxCtype], s_cstart[kMaxCtype];
  __shared__ unsigned long long s_key[kMaxGroup];
  __shared__ unsigned char s_grp[kMaxCtype];
  const SearchTables &T = *P.T;
  const int tid = threadIdx.x;
  LocalList L;
  L.cap = P.lcap;
  L.iob = P.l_iob + (size_t)blockIdx.x * P.lcap;
  L.rdiag = P.l_rdiag + (size_t)blockIdx.x * P.lcap;
  L.rloc = P.l_rloc + (size_t)blockIdx.x * P.lcap;
  L.ccap = P.ccap;
  L.cnd = P.l_cnd + (size_t)blockIdx.x * P.ccap;
  L.cpk = P.l_cpk + (size_t)blockIdx.x * P.ccap;
  const long long npts = (long long)P.nij1 * P.nlev;
  const size_t sl = (size_t)npts;
  if (tid < T.nctype) s_cstart[tid] = P.bstart[T.ct[tid].boff];   // first sorted index of the combined type
  if (tid < T.ngroup)
    for (int m = 0; m < T.grp[tid].n; ++m) s_grp[T.grp[tid].ic[m]] = (unsigned char)tid;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = (long long)atomicAdd(&P.counters[0], 1ull);
    if (tid < kMaxCtype) s_cnt[tid] = 0;
    if (tid < kMaxGroup) s_key[tid] = (T.criterion == 2) ? ~0ull : 0ull;
    __syncthreads();
    const long long wp = s_work;
    if (wp >= npts) break;
    const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
    const size_t pbase = (size_t)ij + (size_t)il * P.nij1;
    const double ri = P.rig1[ij], rj = P.rjg1[ij], rz = P.hgt1[pbase];
    double beta = 1.0;   // relax_beta (letkf_tools.f90:1911-1948), as in the das kernels
    if (P.radar_only && rz > P.zcut) {
      beta = 0.0;
    } else if (P.BOUNDARY_BUFFER_WIDTH > 0.0) {
      const double dist_bdy =
          fmin(fmin(ri - P.IHALO, P.nlon + P.IHALO + 1 - ri) * P.DX,
               fmin(rj - P.JHALO, P.nlat + P.JHALO + 1 - rj) * P.DY) / P.BOUNDARY_BUFFER_WIDTH;
      if (dist_bdy < 1.0) beta = fmax(dist_bdy, 0.0);
    }
    if (beta == 0.0) {   // no obs_local call: work3dn keeps its zeros (:283, :323-352)
      if (tid < 11) P.out[pbase + sl * tid] = 0.0;
      continue;
    }
    Point pt;
    pt.ri = ri;
    pt.rj = rj;
    pt.rz = rz;
    pt.lp = P.logp ? P.logp[pbase] : log(P.pmean[pbase]);
    const int n = search_point(T, P.rec, P.bstart, P.vlfac, pt, L, S);
    __syncthreads();
    if (n < 0) {   // list capacity exceeded (cannot happen with maxl from set_obs): flagged, the host returns an error
      if (tid == 0) atomicAdd(&P.counters[5], 1ull);
      if (tid < 11) P.out[pbase + sl * tid] = -1.0;
      continue;
    }
    for (int i = tid; i < n; i += blockDim.x) {
      const int iob = L.iob[i];
      int ic = T.nctype - 1;
      while (ic > 0 && s_cstart[ic] > iob) --ic;
      atomicAdd(&s_cnt[ic], 1);
      const int g = s_grp[ic];
      if (T.grp[g].limit > 0) {
        if (T.criterion == 1) {
          double nd = 0.0;
          obs_geom(T, T.ct[ic], pt, P.rec[iob], nd);
          atomicMax(&s_key[g], (unsigned long long)__double_as_longlong(nd));     // non-negative doubles order like integers
        } else if (T.criterion == 2) {
          atomicMin(&s_key[g], (unsigned long long)__double_as_longlong(L.rloc[i]));
        } else {
          atomicMax(&s_key[g], (unsigned long long)__double_as_longlong(L.rdiag[i]));
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int rep[5] = {0, 0, 0, 0, 0}, nt[3] = {0, 0, 0};
      double cut[3] = {0.0, 0.0, 0.0};
      auto put = [&](const CtypeDev &c, int cnt, double cutd) {
        const int slot = c.typ == 1 ? 0 : c.typ == 3 ? 1 : c.typ == 4 ? 2 : c.typ == 8 ? 3 : c.typ == 22 ? 4 : -1;
        if (slot >= 0) rep[slot] += cnt;
        if (c.typ == 22 && c.elm_u >= 9 && c.elm_u <= 11) {
          nt[c.elm_u - 9] = cnt;
          cut[c.elm_u - 9] = cutd;
        }
      };
      for (int g = 0; g < T.ngroup; ++g) {
        const GroupDev &G = T.grp[g];
        const CtypeDev &cm = T.ct[G.ic[0]];
        const double cut0 = (T.criterion == 1) ? __dmul_rn(cm.hori_loc, T.dzf) : 0.0;   // (:1385-1389)
        int tot = 0;
        for (int m = 0; m < G.n; ++m) tot += s_cnt[G.ic[m]];
        if (G.limit <= 0) {
          int cum = 0;
          for (int m = 0; m < G.n; ++m) {
            cum += s_cnt[G.ic[m]];
            put(T.ct[G.ic[m]], cum, m == 0 ? cut0 : 0.0);
          }
        } else {
          double cutd = cut0;
          if (tot == G.limit) {
            const double key = __longlong_as_double((long long)s_key[g]);
            cutd = (T.criterion == 1) ? __dmul_rn(cm.hori_loc, __dsqrt_rn(key)) : key;
          }
          put(cm, tot, cutd);
          for (int m = 1; m < G.n; ++m) put(T.ct[G.ic[m]], 0, 0.0);
        }
      }
      for (int f = 0; f < 5; ++f) P.out[pbase + sl * f] = (double)rep[f];
      for (int f = 0; f < 3; ++f) {
        P.out[pbase + sl * (5 + f)] = (double)nt[f];
        P.out[pbase + sl * (8 + f)] = cut[f];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Batched letkf_core twin (common/common_letkf.f90:52-257): explicit trans / transm / pao.
struct CoreParams {
  int ne, nobs, npts, ld, ldk, npairs, ncols;
  const int *nobsl;
  const double *hdxb, *rdiag, *rloc, *dep, *depd;
  double *parm_infl, *trans, *transm, *pao, *transmd;
  int rdiag_wloc, infl_update;
  unsigned long long *counters;
  int max_sweeps;
};

__host__ __device__ inline size_t core_smem_bytes(int k) {
  const int ld = ld_of(k), ldk = ldk_of(k), ncols = 2 * ((k + 1) / 2);
  size_t d = (size_t)ld * ncols + (size_t)kChunk * ldk + ncols + 4 * (size_t)round_up(k, 2) + 3 * kChunk +
             kMaxWarps + 8;
  return d * sizeof(double) + 64;
}

template <int KC>
__global__ void __launch_bounds__(SizeClass<KC>::NT)
core_kernel(const CoreParams P) {
  using SC = SizeClass<KC>;
  constexpr int R = SC::R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.ne, ld = P.ld, ldk = P.ldk, ncols = P.ncols, tid = threadIdx.x;
  double *G = reinterpret_cast<double *>(smem_raw);
  double *Ys = G + (size_t)ld * ncols;
  double *lam = Ys + (size_t)kChunk * ldk;
  double *bv = lam + ncols;                 // b = Yr^T dep
  double *bdv = bv + round_up(k, 2);
  double *tb = bdv + round_up(k, 2);        // G^T b
  double *tbd = tb + round_up(k, 2);
  double *sw = tbd + round_up(k, 2);
  double *sd = sw + kChunk;
  double *sdd = sd + kChunk;
  double *red = sdd + kChunk;
  double *piv = red + kMaxWarps;
  const int ntiles = ((k + 3) / 4) * ((k + 3) / 4 + 1) / 2;
  const size_t k2 = (size_t)k * k;

  for (int pt = blockIdx.x; pt < P.npts; pt += gridDim.x) {
    __syncthreads();
    const int p = P.nobsl[pt];
    double *trans = P.trans + pt * k2;
    double *pao = P.pao ? P.pao + pt * k2 : nullptr;
    double *transm = P.transm ? P.transm + (size_t)pt * k : nullptr;
    double *transmd = (P.transmd && P.depd) ? P.transmd + (size_t)pt * k : nullptr;
    const double infl = P.parm_infl[pt];
    if (p == 0) {   // (:89-107)
      const double sq = sqrt(infl), pd = infl / (double)(k - 1);
      for (size_t i = tid; i < k2; i += blockDim.x) {
        const bool dg = (i / k) == (i % k);
        trans[i] = dg ? sq : 0.0;
        if (pao) pao[i] = dg ? pd : 0.0;
      }
      for (int i = tid; i < k; i += blockDim.x) {
        if (transm) transm[i] = 0.0;
        if (P.transmd) P.transmd[(size_t)pt * k + i] = 0.0;
      }
      continue;
    }
    const double *hd = P.hdxb + (size_t)pt * P.nobs * k;
    const double *rd = P.rdiag + (size_t)pt * P.nobs;
    const double *rl = P.rloc + (size_t)pt * P.nobs;
    const double *dp = P.dep + (size_t)pt * P.nobs;
    const double *dpd = P.depd ? P.depd + (size_t)pt * P.nobs : nullptr;
    double acc[R][16];
    gram_zero<R>(acc);
    double bacc = 0.0, bdacc = 0.0, tracc = 0.0, p1acc = 0.0, p3acc = 0.0;
    for (int o0 = 0; o0 < p; o0 += kChunk) {
      const int nrows = min(kChunk, p - o0);
      if (tid < nrows) {
        const int o = o0 + tid;
        const double winv = P.rdiag_wloc ? 1.0 / rd[o] : rl[o] / rd[o];   // (:111-123)
        const double w = sqrt(winv);
        sw[tid] = w;
        sd[tid] = w * dp[o];
        sdd[tid] = dpd ? w * dpd[o] : 0.0;
        if (P.infl_update) {
          p1acc += dp[o] * dp[o] * winv;
          p3acc += rl[o];
        }
      }
      __syncthreads();
      // hdxb is (nobs, ne) column-major: consecutive threads read consecutive observations
      for (int idx = tid; idx < nrows * ldk; idx += blockDim.x) {
        const int a = idx / nrows, o = idx - a * nrows;
        double v = 0.0;
        if (a < k) v = hd[(size_t)a * P.nobs + o0 + o] * sw[o];
        if (a < ldk) Ys[o * ldk + a] = v;
      }
      __syncthreads();
      gram_accumulate<R>(acc, Ys, nrows, ldk, ntiles);
      if (tid < k) {
        for (int o = 0; o < nrows; ++o) {
          const double y = Ys[o * ldk + tid];
          bacc = fma(y, sd[o], bacc);
          bdacc = fma(y, sdd[o], bdacc);
          tracc = fma(y, y, tracc);
        }
      }
      __syncthreads();
    }
    gram_store<R>(acc, G, ld, k, ntiles, false);
    if (tid < k) {
      bv[tid] = bacc;
      bdv[tid] = bdacc;
    }
    __syncthreads();
    if (tid < k) G[(size_t)tid * ld + tid] += (double)(k - 1) / infl;   // (:140-143)
    if (P.infl_update) {   // (:229-254)
      const double parm1 = block_sum(p1acc, red);
      const double parm2 = block_sum(tid < k ? tracc : 0.0, red) / (double)(k - 1);
      const double parm3 = block_sum(p3acc, red);
      const double parm4 = (parm1 - parm3) / parm2 - infl;
      const double tq = (infl * parm2 + parm3) / parm2;
      const double sigma_o = 2.0 / parm3 * (tq * tq);
      const double gain = 0.04 * 0.04 / (sigma_o + 0.04 * 0.04);
      if (tid == 0) P.parm_infl[pt] = infl + gain * parm4;
    }
    __syncthreads();
    const double c0 = (double)(k - 1) / infl;
    const bool ok = cholesky_lower(G, k, ld, ncols, piv);
    bool conv = false;
    jacobi_onesided<SC::RJ>(G, k, ld, P.npairs, red, P.max_sweeps, c0, &conv);
    column_norms(G, k, ld, ncols, lam);
    double lmax = 0.0, lmin = 1.0e300;
    for (int j = 0; j < ncols; ++j)
      if (lam[j] > 0.0) {
        lmax = fmax(lmax, lam[j]);
        lmin = fmin(lmin, lam[j]);
      }
    if ((!ok || !conv || !(lmin >= lmax * 1.4901161193847656e-08)) && tid == 0)
      atomicAdd(&P.counters[3], 1ull);
    // transm = Pa b = b/c0 + G diag((1/lambda - 1/c0)/lambda) G^T b   (:169-195 restructured: Pa (Yr^T d))
    for (int j = tid; j < ncols; j += blockDim.x) {
      const double *g = G + (size_t)j * ld;
      double s1 = 0.0, s2 = 0.0;
      for (int a = 0; a < k; ++a) {
        s1 = fma(g[a], bv[a], s1);
        s2 = fma(g[a], bdv[a], s2);
      }
      const double l = lam[j];
      const double w = l > 0.0 ? (c0 - l) / (l * c0) / l : 0.0;
      tb[j] = s1 * w;
      tbd[j] = s2 * w;
    }
    __syncthreads();
    double wm = 0.0, wmd = 0.0;
    if (tid < k) {
      for (int j = 0; j < ncols; ++j) {
        const double g = G[(size_t)j * ld + tid];
        wm = fma(g, tb[j], wm);
        wmd = fma(g, tbd[j], wmd);
      }
      wm += bv[tid] / c0;
      wmd += bdv[tid] / c0;
    }
    __syncthreads();
    if (tid < k) {
      if (transm) transm[tid] = wm;
      if (transmd) transmd[tid] = wmd;
      bv[tid] = wm;   // reused below when transm is absent
    }
    __syncthreads();
    // pao = I/c0 - G D1 G^T, trans = sqrt(rho) I - G D2 G^T through the chunked SYRK (rows =
    // scaled columns of G); D1 = (1/c0 - 1/lambda)/lambda >= 0, D2 = (f(c0) - f(lambda))/lambda >= 0,
    // f = sqrt((k-1)/lambda)
    const double fw0 = sqrt((double)(k - 1) / c0);
    for (int which = 0; which < 2; ++which) {
      double *out = which == 0 ? pao : trans;
      if (!out) continue;
      const double dg = which == 0 ? 1.0 / c0 : fw0;
      gram_zero<R>(acc);
      for (int j0 = 0; j0 < ncols; j0 += kChunk) {
        const int nrows = min(kChunk, ncols - j0);
        for (int idx = tid; idx < nrows * ldk; idx += blockDim.x) {
          const int o = idx / ldk, a = idx - o * ldk;
          const double l = lam[j0 + o];
          double sc = 0.0;
          if (l > 0.0) {
            const double d = which == 0 ? (l - c0) / (l * c0) / l : (fw0 - sqrt((double)(k - 1) / l)) / l;
            sc = d > 0.0 ? sqrt(d) : 0.0;
          }
          Ys[idx] = (a < k) ? G[(size_t)(j0 + o) * ld + a] * sc : 0.0;
        }
        __syncthreads();
        gram_accumulate<R>(acc, Ys, nrows, ldk, ntiles);
        __syncthreads();
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int t = tid + r * blockDim.x;
        if (t >= ntiles) continue;
        int ta, tbb;
        tile_coords(t, ta, tbb);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[r][i * 4 + j] = ((4 * ta + i == 4 * tbb + j) ? dg : 0.0) - acc[r][i * 4 + j];
      }
      if (which == 1 && !transm) {   // add the mean weight to every column (:218-226)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int t = tid + r * blockDim.x;
          if (t >= ntiles) continue;
          int ta, tbb;
          tile_coords(t, ta, tbb);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int row = 4 * ta + i, col = 4 * tbb + j;
              if (row < k && col < k && row >= col) {
                const double v = acc[r][i * 4 + j];
                out[(size_t)col * k + row] = v + bv[row];
                if (row != col) out[(size_t)row * k + col] = v + bv[col];
              }
            }
        }
      } else {
        gram_store<R>(acc, out, k, k, ntiles, true);
      }
      __syncthreads();
    }
  }
}

}  // namespace letkf
