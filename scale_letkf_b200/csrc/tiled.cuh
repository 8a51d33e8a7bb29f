// tiled.cuh -- the large-ensemble ("tiled") analysis path: MEMBER > 102, up to thousands of members
// (BASELINE config C4: k = 1000).  Same mathematics as das_ns_kernel.cuh -- twin of the main loop of
// das_letkf (scale/letkf/letkf_tools.f90:313-686) around letkf_core (common/common_letkf.f90:52-257)
// -- but the k x k matrices no longer fit in one SM's shared memory, so a batch of G grid points is
// processed by a sequence of whole-GPU kernels over matrices that live in HBM / L2:
//
//   tl_load      relax_beta, ensemble mean / perturbations (letkf_tools.f90:209-230, :1911-1948)
//   tl_group     per variable-localisation group: unmasked columns, inflation, (k-1)/rho
//   search       local observations of every point of the batch (search.cuh)
//   tl_gather    E_g = [Y^T ; dep^T ; depd^T] R^-1/2  (k+2 rows, obs contiguous), or in the dual form
//                Yt_g = R^-1/2 Y (obs rows, members contiguous)
//   tl_gemm      batched C = A B^T on the FP64 tensor cores (DMMA m8n8k4, cp.async multi-stage tiles);
//                symmetric results compute only the lower triangle of tiles and mirror it
//   tl_rowsum / tl_scale / tl_step / tl_poly + tl_gemm
//                interval-scaled coupled Newton-Schulz Z = (A/s)^-1/2 exactly as ns_solver.cuh, one
//                GEMM launch per product for the whole batch, convergence tracked per point on the device
//   tl_gemm      Ts = Z [dX | b | bd]   (skinny)
//   tl_update    RTPP/RTPS relaxation, xa = xmean + dX T, q-spread clamp, stores
//
// Low-rank ("dual") form, used when every point of the batch has fewer local observations than
// members (p < k, the C4 shape: k = 1000, p <= 200).  With Yt = R^-1/2 Y (p x k), c0 = (k-1)/rho,
// S = Yt Yt^T, B = c0 I + S, C = B^1/2:
//     A^-1/2 = c0^-1/2 I - c0^-1/2 Yt^T C^-1 (c0^1/2 I + C)^-1 Yt,          A^-1/2 Yt^T = Yt^T B^-1/2
// (f(A) = f(c0) I + Yt^T phi(S) Yt with phi(x) = (f(c0 + x) - f(c0)) / x, smooth at x = 0, so a
// rank-deficient S is harmless).  Both factors are inverse square roots of p x p SPD matrices --
// B and D = (c0^1/2 I + C)^2 = c0 I + 2 c0^1/2 C + B -- and come from the same Newton-Schulz kernels
// on p x p instead of k x k matrices: O(p^2 k + p^3) instead of O(p k^2 + k^3) work per point.
#pragma once
#include "das_ns_kernel.cuh"

namespace letkf {

// ------------------------------------------------------------------------------------------------
// Batched NT GEMM on DMMA:  C_g = A_g B_g^T,  A [M][K] (lda), B [N][K] (ldb), C [M][N] (ldc), all
// row-major fp64, dimensions multiples of 8, K even.
struct GemmJob {
  const double *A, *A_alt;   // A_alt: alternative A base chosen per item by `sel` (Z ping-pong buffers)
  const double *B;
  double *C;
  long long sA, sB, sC;      // batch strides in doubles
  int lda, ldb, ldc;
  const double *B_alt;       // alternative B base chosen per item by `selB`
};
struct GemmParams {
  GemmJob job[2];
  int njobs;
  int M, N, K;               // extents (upper bounds when mdims / kdims are given)
  const int *mdims;          // optional per-item M (= N when sym)
  const int *kdims;          // optional per-item K
  const int *state;          // optional per-item solver state (0 active, 1 last quadratic step in flight, 2 done,
                             // 3 / 4 third- / fourth-order finishing step in flight)
  unsigned mask[2];          // job j processes the items whose state bit is set in mask[j]
  const int *sel;            // optional per-item selector of job[].A_alt
  const int *selB;           // optional per-item selector of job[].B_alt
  int sym;                   // 1: C symmetric (M == N): lower-triangular tiles only, mirrored on store
  unsigned long long *res;   // optional per-item max |delta_ij - C_ij| of job 0 (bits of a double >= 0)
};

constexpr int kGemmKC = 16;            // k-extent of one pipeline stage
constexpr int kGemmLDS = kGemmKC + 4;  // smem row stride: 20 doubles -> conflict-free fragment loads

template <int BM, int BN, int STAGES>
__host__ __device__ constexpr size_t gemm_smem_bytes() {
  return (size_t)STAGES * (BM + BN) * kGemmLDS * sizeof(double);
}

// CTA tile BM x BN, one warp per 32 x 32 sub-tile (4 x 4 DMMA accumulators).
template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__((BM / 32) * (BN / 32) * 32) tl_gemm_kernel(const GemmParams P) {
  constexpr int KC = kGemmKC, LDS = kGemmLDS;
  constexpr int NWN = BN / 32, NT = (BM / 32) * (BN / 32) * 32;
  constexpr int ROWS = BM + BN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  const int item = blockIdx.y / P.njobs, jb = blockIdx.y - item * P.njobs;
  if (P.state) {
    const int st = P.state[item];
    if (!((P.mask[jb] >> st) & 1u)) return;
  }
  const int M = P.mdims ? P.mdims[item] : P.M;
  const int N = P.sym ? M : P.N;
  const int K = P.kdims ? P.kdims[item] : P.K;
  int bi, bj;
  if (P.sym) {
    const int t = blockIdx.x;
    bi = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while (bi * (bi + 1) / 2 > t) --bi;
    bj = t - bi * (bi + 1) / 2;
  } else {
    const int ntn = (P.N + BN - 1) / BN;
    bi = blockIdx.x / ntn;
    bj = blockIdx.x - bi * ntn;
  }
  const int m0 = bi * BM, n0 = bj * BN;
  if (m0 >= M || n0 >= N) return;
  const GemmJob &J = P.job[jb];
  const double *A = ((P.sel && P.sel[item]) ? J.A_alt : J.A) + (size_t)item * J.sA;
  const double *B = ((P.selB && P.selB[item]) ? J.B_alt : J.B) + (size_t)item * J.sB;
  double *C = J.C + (size_t)item * J.sC;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = wid / NWN, wn = wid - wm * NWN;
  const int r = lane >> 2, q = lane & 3;

  // stage loader: (BM + BN) rows x KC doubles = ROWS * 8 pieces of 16 bytes
  auto load_stage = [&](int s, int k0) {
    double *dst = sm + (size_t)s * ROWS * LDS;
    for (int pc = tid; pc < ROWS * (KC / 2); pc += NT) {
      const int row = pc >> 3, c2 = (pc & 7) * 2;
      double *d = dst + (size_t)row * LDS + c2;
      const bool isA = row < BM;
      const int gr = isA ? m0 + row : n0 + row - BM;
      const int lim = isA ? M : N;
      if (gr < lim && k0 + c2 < K) {
        const double *src = isA ? A + (size_t)gr * J.lda : B + (size_t)gr * J.ldb;
        cp_async16(d, src + k0 + c2);
      } else {
        d[0] = 0.0;
        d[1] = 0.0;
      }
    }
    cp_async_commit();
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // 8-row blocks of this warp inside the matrix (warp-uniform)
  const int wrow0 = m0 + wm * 32, wcol0 = n0 + wn * 32;
  const int mi_n = min(4, max(0, (M - wrow0 + 7) >> 3)), ni_n = min(4, max(0, (N - wcol0 + 7) >> 3));

  const int nk = (K + KC - 1) / KC;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s * KC);
    else cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {   // prefetch stage kt + STAGES - 1 into the buffer consumed at iteration kt - 1
      const int kn = kt + STAGES - 1;
      if (kn < nk) load_stage(kn % STAGES, kn * KC);
      else cp_async_commit();
    }
    if (mi_n > 0 && ni_n > 0) {
      const double *As = sm + (size_t)(kt % STAGES) * ROWS * LDS + (size_t)(wm * 32 + r) * LDS + q;
      const double *Bs = sm + (size_t)(kt % STAGES) * ROWS * LDS + (size_t)(BM + wn * 32 + r) * LDS + q;
#pragma unroll
      for (int kk = 0; kk < KC; kk += 4) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[(size_t)i * 8 * LDS + kk];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[(size_t)j * 8 * LDS + kk];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: residual, store, mirror
  double resv = 0.0;
  const bool want_res = P.res && jb == 0;
  const bool mirror = P.sym && bi != bj;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i >= mi_n) continue;
    const int row = wrow0 + i * 8 + r;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= ni_n) continue;
      const int col = wcol0 + j * 8 + 2 * q;
      const double c0 = acc[i][j][0], c1 = acc[i][j][1];
      if (want_res) {
        resv = fmax(resv, fabs((row == col ? 1.0 : 0.0) - c0));
        resv = fmax(resv, fabs((row == col + 1 ? 1.0 : 0.0) - c1));
      }
      *reinterpret_cast<double2 *>(C + (size_t)row * J.ldc + col) = make_double2(c0, c1);
      if (mirror) {
        C[(size_t)col * J.ldc + row] = c0;
        C[(size_t)(col + 1) * J.ldc + row] = c1;
      }
    }
  }
  if (want_res) {
    resv = warp_max(resv);
    if (lane == 0 && resv > 0.0) atomicMax(&P.res[item], (unsigned long long)__double_as_longlong(resv));
  }
}

// ------------------------------------------------------------------------------------------------
// Per-batch state of the tiled path (device pointers; [G] = one entry per point of the batch).
struct TiledParams {
  int G;                 // points in this batch
  long long wp0;         // first (ij, ilev) point of the batch, ilev-major
  int n8;                // padded ensemble dimension: round_up(k + 2, 8)
  int kK;                // dual form: round_up(k, 16), member extent of Yt rows
  int maxl;              // capacity of the per-point local lists
  int pK;                // padded obs extent of E / Yt of this batch (multiple of 16)
  int vg;                // variable-localisation group being analysed
  int dual;              // 1: low-rank form (matrices are p x p)
  int max_iter;
  // per point
  double *X;             // [G][kMaxNV][n8]  perturbations of variable vv; rows 14 / 15: b, bd
  double *Ts;            // [G][n8][kMaxNV]  Z [dX | b | bd]  (transposed: member-major)
  double *colsc;         // [G][8][kMaxNV]   xm, xdet, var_g, var_a, s, sd, infl, parm
  double *beta;          // [G]
  double *ri, *rj, *lp, *rz;   // [G] search inputs
  int *skip;             // [G] 1: beta == 0 (point already stored)
  int *ncols;            // [G] unmasked columns of the group
  int *cols;             // [G][kMaxNV]
  double *cdiag;         // [G] (k-1)/rho
  double *infl;          // [G] parm_infl handed to letkf_core
  int *nobsl;            // [G]
  int *idx;              // [G][maxl]
  double *rdiag, *rloc;  // [G][maxl]
  // solver
  int *dims;             // [G] matrix dimension of the Newton-Schulz solve (n8 or round_up(p, 8))
  int *kd;               // [G] round_up(p, 16): K extent of the Gram
  int *state;            // [G] 0 active, 1 last iteration in flight, 2 done / nothing to solve
  int *zsel;             // [G] which of the two Z buffers holds the result
  int *iters;            // [G]
  int *fail;             // [G]
  double *snorm;         // [G] ||A||_1
  unsigned long long *snorm_bits;   // [G] atomicMax target
  double *brk;           // [G] lower end of the eigenvalue bracket
  double *h0, *h1;       // [G] T = h1 M + h0 I
  unsigned long long *res;   // [G] residual bits
  double *misc;          // [G][4]: sum w dep^2, trace(Yr^T Y), snorm of the second solve, spare
  int *nactive;          // [1]
  int *adims;            // [G] M extent of the apply GEMM: n8 for solved points, 0 otherwise
  int *solved_any;       // [G] >= 1 group of the point had local observations
  int nmax;              // leading dimension / item stride root of the solver matrices
  double *E;             // [G][n8][pK] (primal) or [G][pK][kK] (dual Yt)
  double *bZ[2], *bY[2], *mT, *mS;   // [G][nmax][nmax]: ping-pong iterates, T, dual: S then the kept Z1
  double *dw;            // [G][pK] dual: sqrt(w) dep ; [G][pK] sqrt(w) depd follows at + G*pK
  double *U;             // dual: [3][G][kMaxNV][pK] skinny work matrices U1T, U2T, U3T
  double *E2;            // dual: [G][n8][pK] = Yt^T (primal gather layout)
};

__device__ __forceinline__ size_t tl_gaddr(const DasParams &P, int vv, int m, int ij, size_t pbase, size_t sl) {
  return (vv < P.nv3d) ? pbase + ((size_t)m + (size_t)vv * P.nens) * sl
                       : (size_t)ij + ((size_t)m + (size_t)(vv - P.nv3d) * P.nens) * P.nij1;
}

// ---- tl_load: one CTA per point -------------------------------------------------------------------
__global__ void __launch_bounds__(256) tl_load_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.x, tid = threadIdx.x, k = P.k, n8 = B.n8;
  const long long wp = B.wp0 + g;
  const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
  const int nvtot = P.nv3d + (il == 0 ? P.nv2d : 0);
  const size_t sl = (size_t)P.nij1 * P.nlev, pbase = (size_t)ij + (size_t)il * P.nij1;
  __shared__ double xm[kMaxNV];
  double *csc = B.colsc + (size_t)g * 8 * kMaxNV;
  const double ri = P.rig1[ij], rj = P.rjg1[ij], rz = P.hgt1[pbase];
  double beta = 1.0;   // relax_beta (letkf_tools.f90:1911-1948)
  if (P.radar_only && rz > P.zcut) {
    beta = 0.0;
  } else if (P.BOUNDARY_BUFFER_WIDTH > 0.0) {
    const double dist_bdy = fmin(fmin(ri - P.IHALO, P.nlon + P.IHALO + 1 - ri) * P.DX,
                                 fmin(rj - P.JHALO, P.nlat + P.JHALO + 1 - rj) * P.DY) / P.BOUNDARY_BUFFER_WIDTH;
    if (dist_bdy < 1.0) beta = fmax(dist_bdy, 0.0);
  }
  if (tid < kMaxNV) {
    double m = 0.0, d = 0.0, infl = P.INFL_MUL;
    if (tid < nvtot) {
      const double *src = (tid < P.nv3d) ? P.gues3d : P.gues2d;
      m = src[tl_gaddr(P, tid, k, ij, pbase, sl)];
      d = P.det ? src[tl_gaddr(P, tid, k + 1, ij, pbase, sl)] : 0.0;
      if (P.infl_from_field && tid < P.nv3d) infl = P.infl3d[pbase + (size_t)tid * sl];
      if (P.INFL_MUL_MIN > 0.0) infl = fmax(infl, P.INFL_MUL_MIN);
    }
    xm[tid] = m;
    csc[0 * kMaxNV + tid] = m;
    csc[1 * kMaxNV + tid] = d;
    csc[6 * kMaxNV + tid] = infl;
    csc[7 * kMaxNV + tid] = P.RELAX_TO_INFLATED_PRIOR ? infl : 1.0;
  }
  __syncthreads();
  double *X = B.X + (size_t)g * kMaxNV * n8;
  for (int idx = tid; idx < kMaxNV * n8; idx += blockDim.x) {
    const int vv = idx / n8, m = idx - vv * n8;
    double pert = 0.0;
    if (vv < nvtot && m < k) {
      double *src = (vv < P.nv3d) ? P.gues3d : P.gues2d;
      const size_t ad = tl_gaddr(P, vv, m, ij, pbase, sl);
      pert = src[ad] - xm[vv];
      src[ad] = pert;   // gues3d is INTENT(INOUT) "destroyed": perturbations (letkf_tools.f90:209-230)
      if (beta == 0.0) {   // (letkf_tools.f90:333-359)
        double *dst = (vv < P.nv3d) ? P.anal3d : P.anal2d;
        dst[ad] = xm[vv] + pert;
      }
    }
    X[idx] = pert;
  }
  if (beta == 0.0 && P.det && tid < nvtot) {
    double *dst = (tid < P.nv3d) ? P.anal3d : P.anal2d;
    dst[tl_gaddr(P, tid, k + 1, ij, pbase, sl)] = csc[1 * kMaxNV + tid];
  }
  if (tid == 0) {
    B.beta[g] = beta;
    B.solved_any[g] = 0;
    B.skip[g] = beta == 0.0 ? 1 : 0;
    B.ri[g] = ri;
    B.rj[g] = rj;
    B.rz[g] = rz;
    B.lp[g] = P.logp ? P.logp[pbase] : log(xm[P.iv3d_p - 1]);
    atomicAdd(&P.counters[1], 1ull);
  }
}

// ---- tl_group: unmasked columns of group vg, masked stores, inflation, solver reset -----------------
__global__ void __launch_bounds__(128) tl_group_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.x, tid = threadIdx.x, k = P.k, n8 = B.n8;
  const long long wp = B.wp0 + g;
  const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
  const int nvtot = P.nv3d + (il == 0 ? P.nv2d : 0);
  const size_t sl = (size_t)P.nij1 * P.nlev, pbase = (size_t)ij + (size_t)il * P.nij1;
  const double *csc = B.colsc + (size_t)g * 8 * kMaxNV;
  const double *X = B.X + (size_t)g * kMaxNV * n8;
  __shared__ int s_cols[kMaxNV];
  __shared__ int s_nc;
  if (tid == 0) {
    int nc = 0;
    if (!B.skip[g]) {
      const double pmean = csc[P.iv3d_p - 1];
      for (int vv = 0; vv < nvtot; ++vv) {
        if (P.vgroup[vv] != B.vg) continue;
        const bool masked = (vv < P.nv3d) && pmean < P.Q_UPDATE_TOP && (vv + 1) >= P.iv3d_q && (vv + 1) <= P.iv3d_qg;
        s_cols[nc++] = masked ? -(vv + 1) : vv;
      }
    }
    s_nc = nc;
  }
  __syncthreads();
  int nc = 0;
  for (int c = 0; c < s_nc; ++c) {
    const int e = s_cols[c];
    if (e < 0) {   // masked moisture variable above the lid (letkf_tools.f90:371-385)
      const int vv = -e - 1;
      double *dst = P.anal3d;
      for (int m = tid; m < k; m += blockDim.x) dst[tl_gaddr(P, vv, m, ij, pbase, sl)] = csc[vv] + X[(size_t)vv * n8 + m];
      if (tid == 0) {
        if (P.det) dst[tl_gaddr(P, vv, k + 1, ij, pbase, sl)] = csc[1 * kMaxNV + vv];
        if (P.infl3d) P.infl3d[pbase + (size_t)vv * sl] = csc[6 * kMaxNV + vv];
      }
    } else {
      if (tid == 0) B.cols[(size_t)g * kMaxNV + nc] = e;
      ++nc;
    }
  }
  if (tid == 0) {
    B.ncols[g] = nc;
    double infl = 1.0;
    if (nc > 0) {
      int first = -1;
      for (int c = 0; c < s_nc && first < 0; ++c)
        if (s_cols[c] >= 0) first = s_cols[c];
      infl = csc[6 * kMaxNV + first];
    }
    B.infl[g] = infl;
    B.cdiag[g] = (double)(k - 1) / infl;
    B.nobsl[g] = 0;
    B.state[g] = 2;
    B.zsel[g] = 0;
    B.iters[g] = 0;
    B.fail[g] = 0;
    B.snorm_bits[g] = 0ull;
    B.res[g] = 0ull;
    for (int i = 0; i < 4; ++i) B.misc[(size_t)g * 4 + i] = 0.0;
  }
}

// ---- after the search: decide what each point solves --------------------------------------------------
__global__ void tl_plan_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B.G) return;
  int p = B.nobsl[g];
  const bool active = !B.skip[g] && B.ncols[g] > 0;
  if (!active) p = 0;
  if (p < 0) {   // local list overflow
    atomicAdd(&P.counters[5], 1ull);
    p = 0;
  }
  B.nobsl[g] = p;
  B.kd[g] = round_up(p, 16);
  B.state[g] = (active && p > 0) ? 0 : 2;
  B.adims[g] = (active && p > 0) ? B.n8 : 0;
  if (active && p > 0) B.solved_any[g] = 1;
  if (active) {
    if (P.nobsl_out && B.vg == 0) {
      const long long wp = B.wp0 + g;
      P.nobsl_out[wp] = p;   // pbase == wp (ilev-major, ij fastest)
    }
    atomicAdd(&P.counters[4], (unsigned long long)p);
  }
}

// matrix dimension of the solve once the form (primal / dual) of the batch is known
__global__ void tl_dims_kernel(const TiledParams B) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B.G) return;
  B.dims[g] = B.dual ? round_up(B.nobsl[g], 8) : B.n8;
}

// ---- tl_gather (primal): E[m][o] = row_o[m] sqrt(w_o); rows k, k+1 = dep, depd ----------------------
// grid (pK / 32, G), block (32, 8)
__global__ void __launch_bounds__(256) tl_gather_primal_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.y;
  if (B.state[g] >= 2) return;
  const int p = B.nobsl[g], o0 = blockIdx.x * 32;
  if (o0 >= B.kd[g]) return;
  __shared__ double tile[32][33];
  __shared__ double sw[32];
  __shared__ int siob[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (ty == 0) {
    const int o = o0 + tx;
    double w = 0.0;
    int iob = 0;
    if (o < p) {
      iob = B.idx[(size_t)g * B.maxl + o];
      w = sqrt(1.0 / B.rdiag[(size_t)g * B.maxl + o]);
    }
    sw[tx] = w;
    siob[tx] = iob;
  }
  __syncthreads();
  double *E = B.E + (size_t)g * B.n8 * B.pK;
  for (int m0 = 0; m0 < B.n8; m0 += 32) {
    for (int oo = ty; oo < 32; oo += 8) {   // read: obs row oo, members m0 + tx (coalesced along the row)
      const int m = m0 + tx;
      double v = 0.0;
      if (sw[oo] != 0.0 && m < B.n8) v = P.ensval[(size_t)siob[oo] * P.ldens + m] * sw[oo];
      tile[oo][tx] = v;
    }
    __syncthreads();
    for (int mm = ty; mm < 32; mm += 8) {   // write: member row m0 + mm, obs o0 + tx
      const int m = m0 + mm;
      if (m < B.n8 && o0 + tx < B.pK) E[(size_t)m * B.pK + o0 + tx] = tile[tx][mm];
    }
    __syncthreads();
  }
}

// ---- tl_rowsum: s = ||A||_1 of A = Gram + cdiag I, b / bd, adaptive-inflation statistics ---------------
// grid (n8 / 8, G), block 256: one warp per row
__global__ void __launch_bounds__(256) tl_rowsum_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.y;
  if (B.state[g] >= 2) return;
  const int k = P.k, n8 = B.n8, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n8) return;
  const double *Mrow = B.bZ[1] + (size_t)g * n8 * n8 + (size_t)row * n8;
  double *X = B.X + (size_t)g * kMaxNV * n8;
  if (row < k) {
    double rs = 0.0, fs = 0.0;
    for (int c = lane; c < k; c += 32) {
      rs += fabs(Mrow[c]);
      fs = fma(Mrow[c], Mrow[c], fs);
    }
    rs = warp_sum(rs);
    fs = warp_sum(fs);
    if (lane == 0) {
      atomicMax(&B.snorm_bits[g], (unsigned long long)__double_as_longlong(rs));   // ||G||_1
      atomicAdd(&B.misc[(size_t)g * 4 + 2], fs);                                   // ||G||_F^2
      X[(size_t)(kMaxNV - 2) * n8 + row] = Mrow[k];                    // b  = Yr^T dep
      X[(size_t)(kMaxNV - 1) * n8 + row] = P.det ? Mrow[k + 1] : 0.0;  // bd = Yr^T depd
      atomicAdd(&B.misc[(size_t)g * 4 + 1], Mrow[row]);                // trace(Yr^T Y)
    }
  } else {
    if (lane == 0) {
      X[(size_t)(kMaxNV - 2) * n8 + row] = 0.0;
      X[(size_t)(kMaxNV - 1) * n8 + row] = 0.0;
      if (row == k) B.misc[(size_t)g * 4 + 0] = Mrow[k];               // sum w dep^2
    }
  }
}

// ---- tl_scale: Y0 = (src + shift I) / s on the leading nact x nact block, identity on the padding ------
// src may alias dst.  grid (ceil(n*n / 1024), G), block 256.  mode 0: primal (nact = k, shift = cdiag,
// s = snorm_bits); mode 1: dual first solve (nact = p, same); mode 2: dual second solve, src = D already
// shifted (shift = 0, lower bound 4 c0).
__global__ void __launch_bounds__(256) tl_scale_kernel(const DasParams P, const TiledParams B, const double *src,
                                                        double *dst, int nmax, int mode) {
  const int g = blockIdx.y;
  if (B.state[g] >= 2) return;
  const int n = B.dims[g];
  const int nact = mode == 0 ? P.k : B.nobsl[g];
  double s = __longlong_as_double((long long)B.snorm_bits[g]);
  if (mode == 0)   // lambda_max(A) <= c0 + min(||G||_1, ||G||_F)  (tl_rowsum_kernel)
    s = B.cdiag[g] + fmin(s, sqrt(B.misc[(size_t)g * 4 + 2]) * (1.0 + 1.0e-12));
  const double is = 1.0 / s, shift = mode == 2 ? 0.0 : B.cdiag[g];
  const size_t base = (size_t)g * nmax * nmax;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n * n, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int row = e / n, col = e - row * n;
    double v;
    const size_t ad = base + (size_t)row * nmax + col;
    if (row < nact && col < nact) v = (src[ad] + (row == col ? shift : 0.0)) * is;
    else v = row == col ? 1.0 : 0.0;
    dst[ad] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double lo = mode == 2 ? 4.0 * B.cdiag[g] : B.cdiag[g];
    B.brk[g] = lo * is;
    if (mode == 2) B.misc[(size_t)g * 4 + 2] = s; else B.snorm[g] = s;
    B.res[g] = 0ull;
    B.iters[g] = 0;
    // mtx_eigen zeroes eigenvalues below lambda_max sqrt(eps) (common_mtx.f90:69) and letkf_core would
    // then divide by zero; ||A||_1 <= n lambda_max bounds the same condition.
    if (mode != 2 && !(B.cdiag[g] * (double)P.k >= s * 1.4901161193847656e-08)) B.fail[g] = 1;
    // The tiled path still runs the COUPLED Newton-Schulz iteration (Z <- T Z, Y <- T Y on symmetric storage), whose
    // rounding errors grow like sqrt(cond)/4 per step: accurate to ~2e-11 up to lambda_max/c0 = 2e3, 1e-9 class at
    // 5e3 (tools/ns_model.py).  Beyond a bound of 4e3 the point is flagged (LETKF_B200_EEIGEN) instead of returning
    // a silently degraded analysis; the one-CTA-per-point solver (MEMBER <= 102) uses the stable product form.
    if (lo * is < 2.5e-4) B.fail[g] = 1;
  }
}

// ---- tl_res: residual max |I - M| of the first iteration (M = Y0) ---------------------------------------
__global__ void __launch_bounds__(256) tl_res_kernel(const TiledParams B, const double *Mat, int nmax) {
  const int g = blockIdx.y;
  if (B.state[g] >= 2) return;
  const int n = B.dims[g];
  const size_t base = (size_t)g * nmax * nmax;
  double r = 0.0;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n * n, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int row = e / n, col = e - row * n;
    r = fmax(r, fabs((row == col ? 1.0 : 0.0) - Mat[base + (size_t)row * nmax + col]));
  }
  r = warp_max(r);
  if ((threadIdx.x & 31) == 0 && r > 0.0) atomicMax(&B.res[g], (unsigned long long)__double_as_longlong(r));
}

// ---- tl_step: per-point iteration control (the scalar part of newton_schulz_invsqrt) -------------------
// parity: index of the Z buffer the GEMMs of THIS iteration write.  allow_finish: once the residual is below
// 2e-3 one third- / fourth-order step (states 3 / 4, see ns_solver.cuh) finishes the point.
// nactive[0..2]: points still in flight, in state 3 or 4, in state 4.
__global__ void tl_step_kernel(const TiledParams B, int parity, int allow_finish) {
  __shared__ int s_act[3];
  if (threadIdx.x < 3) s_act[threadIdx.x] = 0;
  __syncthreads();
  for (int g = threadIdx.x; g < B.G; g += blockDim.x) {
    int st = B.state[g];
    if (st == 1 || st == 3 || st == 4) st = 2;   // the finishing GEMMs were issued by the previous iteration
    if (st == 0) {
      const double res = __longlong_as_double((long long)B.res[g]);
      const int it = B.iters[g] + 1;
      const bool conv = res < 1.0e-7;
      const bool last = conv || it >= B.max_iter;
      const bool fin = allow_finish && !last && it > 1 && res < 2.0e-3;
      double a = B.brk[g], c = 1.0;
      if (!last && !fin && (1.0 - a) > 1.0e-3) c = 3.0 / (a + sqrt(a) + 1.0);
      const double sc = sqrt(c);
      B.h0[g] = 1.5 * sc;
      B.h1[g] = -0.5 * c * sc;
      const double t = c * a;
      B.brk[g] = t * (3.0 - t) * (3.0 - t) * 0.25;
      B.iters[g] = it;
      B.res[g] = 0ull;
      if (last || fin) {
        st = last ? 1 : (res < 2.0e-4 ? 3 : 4);
        B.zsel[g] = parity;
        if (last && !conv) B.fail[g] = 1;
      }
    }
    B.state[g] = st;
    if (st != 2) atomicAdd(&s_act[0], 1);
    if (st == 3 || st == 4) atomicAdd(&s_act[1], 1);
    if (st == 4) atomicAdd(&s_act[2], 1);
  }
  __syncthreads();
  if (threadIdx.x < 3) B.nactive[threadIdx.x] = s_act[threadIdx.x];
}

// ---- tl_poly: T = h1 M + h0 I (and Z1 = T on the first iteration) --------------------------------------
__global__ void __launch_bounds__(256) tl_poly_kernel(const TiledParams B, const double *Mat, double *T, double *Zfirst,
                                                       int nmax) {
  const int g = blockIdx.y;
  if (B.state[g] == 2) return;
  const int n = B.dims[g];
  const bool fin = B.state[g] >= 3;              // finishing step: T holds E = I - M first
  const double h0 = fin ? 1.0 : B.h0[g], h1 = fin ? -1.0 : B.h1[g];
  const size_t base = (size_t)g * nmax * nmax;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n * n, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int row = e / n, col = e - row * n;
    const size_t ad = base + (size_t)row * nmax + col;
    const double v = fma(h1, Mat[ad], row == col ? h0 : 0.0);
    T[ad] = v;
    if (Zfirst) Zfirst[ad] = v;
  }
}

// ---- tl_finish_poly: T = I + E/2 + 3 E^2/8 (+ 5 E^3/16) for the points in state 3 (4); E in T (in place) ----
__global__ void __launch_bounds__(256) tl_finish_poly_kernel(const TiledParams B, double *T, const double *E2, const double *E3,
                                                              int nmax) {
  const int g = blockIdx.y;
  const int st = B.state[g];
  if (st != 3 && st != 4) return;
  const int n = B.dims[g];
  const size_t base = (size_t)g * nmax * nmax;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n * n, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int row = e / n, col = e - row * n;
    const size_t ad = base + (size_t)row * nmax + col;
    double v = fma(0.375, E2[ad], fma(0.5, T[ad], row == col ? 1.0 : 0.0));
    if (st == 4) v = fma(0.3125, E3[ad], v);
    T[ad] = v;
  }
}

// ---- tl_update: scalars, relaxation, update, stores (letkf_tools.f90:457-513) ---------------------------
// One CTA per point.  Ts is member-major [n8][kMaxNV].  dual == 0: Ts = Z [dX|b|bd] with Z = (A/s)^-1/2;
// dual == 1: Ts already holds t_c = A^-1/2 x_c (s = 1).
__global__ void __launch_bounds__(256) tl_update_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, k = P.k, n8 = B.n8;
  if (B.skip[g]) return;
  const int nc = B.ncols[g];
  if (nc == 0) return;
  const long long wp = B.wp0 + g;
  const int il = (int)(wp / P.nij1), ij = (int)(wp - (long long)il * P.nij1);
  const size_t sl = (size_t)P.nij1 * P.nlev, pbase = (size_t)ij + (size_t)il * P.nij1;
  __shared__ double csc[8 * kMaxNV];
  __shared__ double red[kMaxWarps];
  __shared__ int cols[kMaxNV];
  __shared__ double s_inflv;
  for (int i = tid; i < 8 * kMaxNV; i += blockDim.x) csc[i] = B.colsc[(size_t)g * 8 * kMaxNV + i];
  if (tid < kMaxNV) cols[tid] = tid < nc ? B.cols[(size_t)g * kMaxNV + tid] : 0;
  __syncthreads();
  double *xm = csc, *xdet = csc + kMaxNV, *varg = csc + 2 * kMaxNV, *vara = csc + 3 * kMaxNV;
  double *ssum = csc + 4 * kMaxNV, *sdsum = csc + 5 * kMaxNV, *inflv = csc + 6 * kMaxNV, *parmv = csc + 7 * kMaxNV;
  const double *X = B.X + (size_t)g * kMaxNV * n8;
  double *Ts = B.Ts + (size_t)g * n8 * kMaxNV;
  const int p = B.nobsl[g];
  const double beta = B.beta[g], infl = B.infl[g];
  const int vtrig = cols[0];
  double wscale, pscale;
  if (p > 0) {
    const double s = B.dual ? 1.0 : B.snorm[g];
    wscale = sqrt((double)(k - 1) / s);
    pscale = 1.0 / s;
    if (P.INFL_MUL_ADAPTIVE) {   // (common_letkf.f90:229-254)
      double p3 = 0.0;
      for (int o = tid; o < p; o += blockDim.x) p3 += B.rloc[(size_t)g * B.maxl + o];
      const double parm3 = block_sum(p3, red);
      const double parm1 = B.misc[(size_t)g * 4 + 0];
      const double parm2 = B.misc[(size_t)g * 4 + 1] / (double)(k - 1);
      const double parm4 = (parm1 - parm3) / parm2 - infl;
      const double tq = (infl * parm2 + parm3) / parm2;
      const double sigma_o = 2.0 / parm3 * (tq * tq);
      const double gain = 0.04 * 0.04 / (sigma_o + 0.04 * 0.04);
      if (tid == 0) s_inflv = infl + gain * parm4;
      __syncthreads();
      if (tid == 0) inflv[vtrig] = s_inflv;
      __syncthreads();
    }
  } else {   // nobsl == 0 (common_letkf.f90:89-107): W = sqrt(infl) I, wbar = 0, Pa = infl/(k-1) I
    wscale = sqrt(infl);
    pscale = infl / (double)(k - 1);
    for (int idx = tid; idx < n8 * kMaxNV; idx += blockDim.x) {
      const int m = idx / kMaxNV, vv = idx - m * kMaxNV;
      Ts[idx] = (vv < kMaxNV - 2) ? X[(size_t)vv * n8 + m] : 0.0;
    }
    __syncthreads();
  }
  {
    const int nw = blockDim.x >> 5;
    for (int c = w; c < nc; c += nw) {
      const int vv = cols[c];
      const double *x = X + (size_t)vv * n8;
      double vg_ = 0.0, va_ = 0.0, s_ = 0.0, sdv_ = 0.0;
      for (int m = lane; m < k; m += 32) {
        const double xv = x[m], tv = Ts[(size_t)m * kMaxNV + vv];
        vg_ = fma(xv, xv, vg_);
        va_ = fma(tv, tv, va_);
        s_ = fma(tv, Ts[(size_t)m * kMaxNV + kMaxNV - 2], s_);
        sdv_ = fma(tv, Ts[(size_t)m * kMaxNV + kMaxNV - 1], sdv_);
      }
      vg_ = warp_sum(vg_);
      va_ = warp_sum(va_);
      s_ = warp_sum(s_);
      sdv_ = warp_sum(sdv_);
      if (lane == 0) {
        varg[c] = vg_;
        vara[c] = va_ * pscale;
        ssum[c] = s_ * pscale;
        sdsum[c] = P.det ? sdv_ * pscale : 0.0;
      }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < nc * k; idx += blockDim.x) {
    const int c = idx / k, m = idx - c * k;
    const int vv = cols[c];
    const double x = X[(size_t)vv * n8 + m];
    const double z = wscale * Ts[(size_t)m * kMaxNV + vv];   // (W dx)_m
    const double parm = parmv[vv];
    double wx;
    if (P.RELAX_ALPHA != 0.0) {
      wx = (1.0 - P.RELAX_ALPHA) * z + P.RELAX_ALPHA * sqrt(parm) * x;
    } else if (P.RELAX_ALPHA_SPREAD != 0.0) {
      double f = 1.0;
      if (varg[c] > 0.0 && vara[c] > 0.0)
        f = P.RELAX_ALPHA_SPREAD * sqrt(varg[c] * parm / (vara[c] * (double)(k - 1))) - P.RELAX_ALPHA_SPREAD + 1.0;
      wx = f * z;
    } else {
      wx = z;
    }
    Ts[(size_t)m * kMaxNV + vv] = xm[vv] + (wx + ssum[c]) * beta + (1.0 - beta) * x;
  }
  __syncthreads();
  if (P.Q_SPRD_MAX > 0.0) {   // (letkf_tools.f90:500-513)
    for (int c = 0; c < nc; ++c) {
      if (cols[c] != P.iv3d_q - 1) continue;
      const int vv = cols[c];
      double part = 0.0;
      for (int m = tid; m < k; m += blockDim.x) part += Ts[(size_t)m * kMaxNV + vv];
      const double q_mean = block_sum(part, red) / (double)k;
      part = 0.0;
      for (int m = tid; m < k; m += blockDim.x) {
        const double d = Ts[(size_t)m * kMaxNV + vv] - q_mean;
        part = fma(d, d, part);
      }
      const double q_sprd = sqrt(block_sum(part, red) / (double)(k - 1)) / q_mean;
      if (q_sprd > P.Q_SPRD_MAX) {
        for (int m = tid; m < k; m += blockDim.x) {
          const double d = Ts[(size_t)m * kMaxNV + vv] - q_mean;
          Ts[(size_t)m * kMaxNV + vv] = q_mean + d * P.Q_SPRD_MAX / q_sprd;
        }
      }
      __syncthreads();
    }
  }
  for (int idx = tid; idx < nc * k; idx += blockDim.x) {
    const int c = idx / k, m = idx - c * k;
    const int vv = cols[c];
    double *dst = (vv < P.nv3d) ? P.anal3d : P.anal2d;
    dst[tl_gaddr(P, vv, m, ij, pbase, sl)] = Ts[(size_t)m * kMaxNV + vv];
  }
  if (tid < nc) {
    const int vv = cols[tid];
    if (P.det) {
      double *dst = (vv < P.nv3d) ? P.anal3d : P.anal2d;
      dst[tl_gaddr(P, vv, k + 1, ij, pbase, sl)] = xdet[vv] + sdsum[tid] * beta;   // (:489-497)
    }
    if (P.rtps_out && vv < P.nv3d) {
      double f = 1.0;
      if (P.RELAX_ALPHA == 0.0 && P.RELAX_ALPHA_SPREAD != 0.0 && varg[tid] > 0.0 && vara[tid] > 0.0)
        f = P.RELAX_ALPHA_SPREAD * sqrt(varg[tid] * parmv[vv] / (vara[tid] * (double)(k - 1))) - P.RELAX_ALPHA_SPREAD + 1.0;
      P.rtps_out[pbase + (size_t)vv * sl] = f;
    }
    if (P.infl3d && vv < P.nv3d) {
      const double v = (vv == vtrig || P.INFL_MUL_ADAPTIVE) ? inflv[P.INFL_MUL_ADAPTIVE ? P.vfirst[vv] : vv] : inflv[vv];
      P.infl3d[pbase + (size_t)vv * sl] = v;
    }
  }
  if (tid == 0) {
    if (p > 0) {
      atomicAdd(&P.counters[6], (unsigned long long)B.iters[g]);
      if (B.fail[g]) atomicAdd(&P.counters[3], 1ull);
    }
  }
}

// points with >= 1 solved group (das_stats "solved")
__global__ void tl_count_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B.G) return;
  if (B.solved_any[g]) atomicAdd(&P.counters[2], 1ull);
}

// ================================================================================================
// Dual (observation-space) form
// ================================================================================================

// Yt[o][m] = ensval[iob_o][m] sqrt(w_o) (m < k; zero beyond), dw = sqrt(w) dep, dwd = sqrt(w) depd.
// grid (pK / 8, G), block 256 (one warp per obs row)
__global__ void __launch_bounds__(256) tl_gather_dual_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.y;
  if (B.state[g] >= 2) return;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, k = P.k;
  if (o >= B.kd[g]) return;
  const int p = B.nobsl[g];
  double *row = B.E + ((size_t)g * B.pK + o) * B.kK;
  double w = 0.0;
  const double *src = nullptr;
  if (o < p) {
    const int iob = B.idx[(size_t)g * B.maxl + o];
    w = sqrt(1.0 / B.rdiag[(size_t)g * B.maxl + o]);
    src = P.ensval + (size_t)iob * P.ldens;
  }
  for (int m = lane; m < B.kK; m += 32) row[m] = (src && m < k) ? src[m] * w : 0.0;
  if (lane == 0) {
    B.dw[(size_t)g * B.pK + o] = src ? src[k] * w : 0.0;
    B.dw[(size_t)(B.G + g) * B.pK + o] = (src && P.det) ? src[k + 1] * w : 0.0;
  }
}

// s = ||B||_1 with B = S + c0 I over the leading p x p block; also sum w dep^2, trace(S) = trace(Yr^T Y).
// grid (pK / 8, G), block 256: one warp per row.  mode 2: src is D (no shift).
__global__ void __launch_bounds__(256) tl_rowsum_dual_kernel(const TiledParams B, const double *src, int nmax, int mode) {
  const int g = blockIdx.y;
  if (B.state[g] >= 2) return;
  const int p = B.nobsl[g], lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p) return;
  const double *Mrow = src + (size_t)g * nmax * nmax + (size_t)row * nmax;
  double rs = 0.0;
  for (int c = lane; c < p; c += 32) rs += fabs(Mrow[c]);
  rs = warp_sum(rs);
  if (lane == 0) {
    if (mode != 2) {
      rs += B.cdiag[g];
      const double d = B.dw[(size_t)g * B.pK + row];
      atomicAdd(&B.misc[(size_t)g * 4 + 0], d * d);
      atomicAdd(&B.misc[(size_t)g * 4 + 1], Mrow[row]);
    }
    atomicMax(&B.snorm_bits[g], (unsigned long long)__double_as_longlong(rs));
  }
}

// D = c0 I + 2 sqrt(c0) C + B on the leading p x p block, with C = sqrt(s) Yfin, B = S + c0 I.
// S is read from mS, Yfin from mY; D is written to mD (may alias neither).  Also resets snorm_bits.
__global__ void __launch_bounds__(256) tl_dual_d_kernel(const TiledParams B, double *D, int nmax) {
  const int g = blockIdx.y;
  if (B.adims[g] == 0) return;
  const double *S = B.mS, *Yfin = B.bY[B.zsel[g]];
  const int n = B.dims[g], p = B.nobsl[g];
  const double c0 = B.cdiag[g], sq = sqrt(B.snorm[g]), tc = 2.0 * sqrt(c0);
  const size_t base = (size_t)g * nmax * nmax;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n * n, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int row = e / n, col = e - row * n;
    const size_t ad = base + (size_t)row * nmax + col;
    double v = 0.0;
    if (row < p && col < p) v = S[ad] + tc * sq * Yfin[ad] + (row == col ? 2.0 * c0 : 0.0);
    else if (row == col) v = 1.0;
    D[ad] = v;
  }
}

// keep the result of the first solve: mS <- bZ[zsel] (S is no longer needed once D is formed)
__global__ void __launch_bounds__(256) tl_dual_keep_kernel(const TiledParams B, int nmax) {
  const int g = blockIdx.y;
  if (B.adims[g] == 0) return;
  const int n = B.dims[g];
  const size_t base = (size_t)g * nmax * nmax;
  const double *Z = B.bZ[B.zsel[g]];
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n * n, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int row = e / n, col = e - row * n;
    const size_t ad = base + (size_t)row * nmax + col;
    B.mS[ad] = Z[ad];
  }
}
// re-arm the solver state for the second solve of the dual form
__global__ void tl_dual_restart_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= B.G || B.adims[g] == 0) return;
  atomicAdd(&P.counters[6], (unsigned long long)B.iters[g]);
  B.state[g] = 0;
  B.snorm_bits[g] = 0ull;
  B.res[g] = 0ull;
}

// Dual apply as skinny GEMMs (tl_gemm) around two small kernels:
//   U1T = X Yt^T                 (16 x p)   u_c = Yt x_c
//   U2T = U1T Z2 ; rows 14, 15 <- dw, dwd   (tl_dual_setb)
//   U3T = U2T Z1                 (16 x p)
//   Ts  = E U3T^T                (n8 x 16)  with E = Yt^T (gathered in the primal layout)
//   tl_dual_fin:  t_c = x_c / sqrt(c0) - fv Ts_c  (variables),  t_b = fb Ts_b   (b, bd)
// Z1 = (B/s1)^-1/2 = sqrt(s1) C^-1,  Z2 = (D/s2)^-1/2 = sqrt(s2) (sqrt(c0) I + C)^-1.
__global__ void __launch_bounds__(256) tl_dual_setb_kernel(const TiledParams B, double *U2T) {
  const int g = blockIdx.y;
  if (B.adims[g] == 0) return;
  const int o = blockIdx.x * 256 + threadIdx.x;
  if (o >= B.pK) return;
  const int p = B.nobsl[g];
  double *u = U2T + (size_t)g * kMaxNV * B.pK;
  u[(size_t)(kMaxNV - 2) * B.pK + o] = o < p ? B.dw[(size_t)g * B.pK + o] : 0.0;
  u[(size_t)(kMaxNV - 1) * B.pK + o] = o < p ? B.dw[(size_t)(B.G + g) * B.pK + o] : 0.0;
}
__global__ void __launch_bounds__(256) tl_dual_fin_kernel(const DasParams P, const TiledParams B) {
  const int g = blockIdx.y;
  if (B.adims[g] == 0) return;
  const int n8 = B.n8, k = P.k;
  const double c0 = B.cdiag[g], s1 = B.snorm[g], s2 = B.misc[(size_t)g * 4 + 2];
  const double isc0 = 1.0 / sqrt(c0), fv = isc0 / (sqrt(s1) * sqrt(s2)), fb = 1.0 / sqrt(s1);
  const double *X = B.X + (size_t)g * kMaxNV * n8;
  double *Ts = B.Ts + (size_t)g * n8 * kMaxNV;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min(n8 * kMaxNV, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int m = e / kMaxNV, c = e - m * kMaxNV;
    const double acc = Ts[e];
    double v;
    if (c < kMaxNV - 2) v = (m < k) ? X[(size_t)c * n8 + m] * isc0 - fv * acc : 0.0;
    else v = (m < k) ? fb * acc : 0.0;
    Ts[e] = v;
  }
}

// ================================================================================================
// letkf_core twin for large ensembles (common/common_letkf.f90:52-257, ne > 128): explicit trans / transm /
// pao from the same tiled machinery (primal form).
// ================================================================================================
struct CoreTiledParams {
  int ne, nobs, npts, rdiag_wloc, infl_update;
  int pt0;               // first point of the batch
  const int *nobsl;
  const double *hdxb, *rdiag, *rloc, *dep, *depd;
  double *parm_infl, *trans, *transm, *pao, *transmd;
  double *sw;            // [G][pK] sqrt(R^-1 weight)
  unsigned long long *counters;
};

// per-point solver state + weights; grid (ceil(pK / 256), G)
__global__ void __launch_bounds__(256) tlc_init_kernel(const CoreTiledParams C, const TiledParams B) {
  const int g = blockIdx.y, pt = C.pt0 + g, o = blockIdx.x * 256 + threadIdx.x;
  const int p = C.nobsl[pt];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double infl = C.parm_infl[pt];
    B.nobsl[g] = p;
    B.kd[g] = round_up(p, 16);
    B.dims[g] = B.n8;
    B.state[g] = p > 0 ? 0 : 2;
    B.adims[g] = p > 0 ? B.n8 : 0;
    B.skip[g] = 0;
    B.infl[g] = infl;
    B.cdiag[g] = (double)(C.ne - 1) / infl;
    B.zsel[g] = 0;
    B.iters[g] = 0;
    B.fail[g] = 0;
    B.snorm_bits[g] = 0ull;
    B.res[g] = 0ull;
    for (int i = 0; i < 3; ++i) B.misc[(size_t)g * 4 + i] = 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 32) B.misc[(size_t)g * 4 + 3] = 0.0;
  if (o >= B.pK) return;
  double w = 0.0;
  if (o < p) {
    const size_t a = (size_t)pt * C.nobs + o;
    const double winv = C.rdiag_wloc ? 1.0 / C.rdiag[a] : C.rloc[a] / C.rdiag[a];   // (:111-123)
    w = sqrt(winv);
  }
  C.sw[(size_t)g * B.pK + o] = w;
}
// sum of rloc (adaptive inflation, :229-254); grid G, block 256 -- after tlc_init
__global__ void __launch_bounds__(256) tlc_rlocsum_kernel(const CoreTiledParams C, const TiledParams B) {
  __shared__ double red[kMaxWarps];
  const int g = blockIdx.x, pt = C.pt0 + g, p = C.nobsl[pt];
  double s = 0.0;
  for (int o = threadIdx.x; o < p; o += blockDim.x) s += C.rloc[(size_t)pt * C.nobs + o];
  s = block_sum(s, red);
  if (threadIdx.x == 0) B.misc[(size_t)g * 4 + 3] = s;
}
// E[m][o] = hdxb(o, m) sqrt(w_o); rows ne, ne+1: dep, depd; grid (n8 / 8, G), block 256 (one warp per row)
__global__ void __launch_bounds__(256) tlc_gather_kernel(const CoreTiledParams C, const TiledParams B) {
  const int g = blockIdx.y, pt = C.pt0 + g;
  if (B.state[g] >= 2) return;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (m >= B.n8) return;
  const int kd = B.kd[g];
  const double *sw = C.sw + (size_t)g * B.pK;
  double *row = B.E + ((size_t)g * B.n8 + m) * B.pK;
  const double *src = nullptr;
  if (m < C.ne) src = C.hdxb + ((size_t)pt * C.ne + m) * C.nobs;   // hdxb(nobs, ne) column-major: row m contiguous
  else if (m == C.ne) src = C.dep + (size_t)pt * C.nobs;
  else if (m == C.ne + 1 && C.depd) src = C.depd + (size_t)pt * C.nobs;
  for (int o = lane; o < kd; o += 32) {
    const double w = sw[o];
    row[o] = (src && w != 0.0) ? src[o] * w : 0.0;
  }
}
// transm = Pa b, transmd = Pa bd with Pa = Z Z / s in mT (unscaled Z Z); grid G, block 256
__global__ void __launch_bounds__(256) tlc_transm_kernel(const CoreTiledParams C, const TiledParams B) {
  const int g = blockIdx.x, pt = C.pt0 + g, ne = C.ne, n8 = B.n8;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double *wm = B.Ts + (size_t)g * n8 * kMaxNV;          // scratch: [0..n8) transm, [n8..2 n8) transmd
  if (B.nobsl[g] == 0) {
    for (int i = threadIdx.x; i < ne; i += blockDim.x) {
      wm[i] = 0.0;
      wm[n8 + i] = 0.0;
      if (C.transm) C.transm[(size_t)pt * ne + i] = 0.0;
      if (C.transmd) C.transmd[(size_t)pt * ne + i] = 0.0;
    }
    return;
  }
  const double is = 1.0 / B.snorm[g];
  const double *ZZ = B.mT + (size_t)g * n8 * n8;
  const double *b = B.X + (size_t)g * kMaxNV * n8 + (size_t)(kMaxNV - 2) * n8, *bd = b + n8;
  for (int i = w; i < ne; i += nw) {
    double s1 = 0.0, s2 = 0.0;
    for (int j = lane; j < ne; j += 32) {
      const double z = ZZ[(size_t)i * n8 + j];
      s1 = fma(z, b[j], s1);
      s2 = fma(z, bd[j], s2);
    }
    s1 = warp_sum(s1) * is;
    s2 = warp_sum(s2) * is;
    if (lane == 0) {
      wm[i] = s1;
      wm[n8 + i] = s2;
      if (C.transm) C.transm[(size_t)pt * ne + i] = s1;
      if (C.transmd && C.depd) C.transmd[(size_t)pt * ne + i] = s2;
    }
  }
  if (C.infl_update && threadIdx.x == 0) {   // (:229-254)
    const double infl = B.infl[g];
    const double parm1 = B.misc[(size_t)g * 4 + 0], parm2 = B.misc[(size_t)g * 4 + 1] / (double)(ne - 1);
    const double parm3 = B.misc[(size_t)g * 4 + 3];
    const double parm4 = (parm1 - parm3) / parm2 - infl;
    const double tq = (infl * parm2 + parm3) / parm2;
    const double sigma_o = 2.0 / parm3 * (tq * tq);
    const double gain = 0.04 * 0.04 / (sigma_o + 0.04 * 0.04);
    C.parm_infl[pt] = infl + gain * parm4;
  }
  if (threadIdx.x == 0 && B.fail[g]) atomicAdd(&C.counters[3], 1ull);
}
// trans = sqrt((ne-1)/s) Z (+ transm on every column when transm is absent, :218-226), pao = Z Z / s;
// grid (ceil(ne^2 / 1024), G), block 256; outputs column-major (ne, ne)
__global__ void __launch_bounds__(256) tlc_out_kernel(const CoreTiledParams C, const TiledParams B) {
  const int g = blockIdx.y, pt = C.pt0 + g, ne = C.ne, n8 = B.n8;
  const size_t k2 = (size_t)ne * ne;
  const int p = B.nobsl[g];
  const double infl = B.infl[g];
  const double *Z = B.bZ[B.zsel[g]] + (size_t)g * n8 * n8, *ZZ = B.mT + (size_t)g * n8 * n8;
  const double *wm = B.Ts + (size_t)g * n8 * kMaxNV;
  const double s = p > 0 ? B.snorm[g] : 1.0;
  const double ft = sqrt((double)(ne - 1) / s), is = 1.0 / s;
  for (int e = blockIdx.x * 1024 + threadIdx.x; e < min((int)k2, (int)(blockIdx.x + 1) * 1024); e += 256) {
    const int j = e / ne, i = e - j * ne;   // column j, row i
    double t, pa;
    if (p == 0) {   // (:89-107)
      t = i == j ? sqrt(infl) : 0.0;
      pa = i == j ? infl / (double)(ne - 1) : 0.0;
    } else {
      t = ft * Z[(size_t)i * n8 + j];
      pa = ZZ[(size_t)i * n8 + j] * is;
      if (!C.transm) t += wm[i];
    }
    C.trans[(size_t)pt * k2 + e] = t;
    if (C.pao) C.pao[(size_t)pt * k2 + e] = pa;
  }
}

}  // namespace letkf
