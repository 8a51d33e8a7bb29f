// ns_solver.cuh -- tensor-core (FP64 DMMA) weight solve of the LETKF analysis, one CTA per grid
// point, replacing mtx_eigen / EISPACK rs (common/common_mtx.f90:41, common/netlibrs.f:21) and the
// four dgemm calls of letkf_core (common/common_letkf.f90:127,156,169,205).
//
// letkf_core only ever uses functions of A = Yr^T Y + (k-1)/rho I:
//     Pa = A^-1,   trans = sqrt(k-1) A^-1/2,   transm = Pa Yr^T d
// so no eigenvectors are needed.  Z = A^-1/2 is computed with the coupled Newton-Schulz iteration
// (Higham, Functions of Matrices, eq. 6.35), interval-scaled each step:
//     M = Z Y;  T = sqrt(c) (3 I - c M) / 2;  Z <- T Z;  Y <- T Y;      Y0 = A / s, Z0 = I
// with s = ||A||_1 >= lambda_max, the eigenvalues of M bracketed by [a, b] (a0 = c0/s with the known
// lowest eigenvalue bound c0 = (k-1)/rho, b0 = 1) and c = 3 / (a + sqrt(ab) + b), the scaling that
// maps both interval ends onto the same image.  Every iterate is a polynomial in A, so all
// matrices are symmetric and commute: the iteration is numerically stable and quadratically
// convergent even with the (k-p)-fold degenerate eigenvalue c0 (p < k).  Once the residual ||I - Z Y|| is below
// 2e-3 a single third- or fourth-order step (Z <- (I + E/2 + 3E^2/8 [+ 5E^3/16]) Z) finishes the solve: 3.9-4.7
// iterations per solve on the BASELINE shapes instead of 5.3-6.4 with quadratic steps only.
//
// All products are GEMMs on the FP64 tensor cores: mma.sync.aligned.m8n8k4.f64 (DMMA; tcgen05/TMEM
// has no FP64 kind).  Because every matrix is symmetric, only the lower triangle of 8x8 tiles is
// stored (packed, XOR-swizzled so that both the direct and the transposed fragment patterns are
// bank-conflict free) and only one tile of each symmetric pair is computed: with NB (odd) row
// blocks, warp w computes the "circulant" tiles (w, (w+d) mod NB), d = 0..(NB-1)/2, which covers
// every unordered pair exactly once with perfect load balance.
#pragma once
#include "common.cuh"

namespace letkf {

template <int NB_>
struct NsCfg {
  static_assert(NB_ % 2 == 1, "the circulant tile assignment needs an odd number of row blocks");
  static constexpr int NB = NB_;                   // 8-row blocks
  static constexpr int H = (NB_ - 1) / 2;          // warp w owns tiles (w, w+d mod NB), d = 0..H
  static constexpr int KP = 8 * NB_;               // padded ensemble size (>= k + 2)
  static constexpr int LD = KP + 4;                // row stride of row-major staging / vector blocks
  static constexpr int NT = 32 * NB_;              // one warp per row block
  static constexpr int NTILE = NB_ * (NB_ + 1) / 2;
  static constexpr int PSZ = NTILE * 64;           // doubles per packed symmetric matrix
  static constexpr int CR = (PSZ / LD) & ~3;       // obs rows per staging chunk (three chunks fit in 3 PSZ); CR == 4 NB
  // resident CTAs per SM the register allocation is sized for
  static constexpr int MINB = NB_ <= 3 ? 8 : NB_ <= 5 ? 5 : NB_ <= 7 ? 4 : NB_ <= 9 ? 2 : 1;
  // (registers: each of the 4 SM sub-partitions holds 16 K registers and ceil(NB MINB / 4) warps, which
  // is what __launch_bounds__(NT, MINB) makes ptxas budget for -- 128 for NB = 13, 80 for NB = 7)
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// ---- packed symmetric tile storage ---------------------------------------------------------------
// tile (bi, bj), bi >= bj, at ((bi (bi+1))/2 + bj) * 64; element (r, c) of a stored tile at
// r*8 + (c ^ ((r & 2) << 1)).
__host__ __device__ __forceinline__ int ptile(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * 64; }
__host__ __device__ __forceinline__ int pelem(int r, int c) { return r * 8 + (c ^ ((r & 2) << 1)); }
// address of logical element (row, col) of a packed symmetric matrix
__host__ __device__ __forceinline__ int paddr(int row, int col) {
  const int bi = row >> 3, bj = col >> 3;
  return (bi >= bj) ? ptile(bi, bj) + pelem(row & 7, col & 7) : ptile(bj, bi) + pelem(col & 7, row & 7);
}

struct LaneOfs {   // per-lane offsets of the two fragment patterns, k-halves h = 0, 1
  int p1[2];       // stored (r, 4h+q): direct A operand / transposed B operand
  int p2[2];       // stored (4h+q, r): transposed A operand / direct B operand
  int st_n;        // accumulator store, direct:     stored (r, 2q) [double2]
  int st_t[2];     // accumulator store, transposed: stored (2q+e, r)
};
__device__ __forceinline__ LaneOfs lane_offsets(int lane) {
  const int r = lane >> 2, q = lane & 3;
  LaneOfs o;
  o.p1[0] = pelem(r, q);
  o.p1[1] = pelem(r, 4 + q);
  o.p2[0] = pelem(q, r);
  o.p2[1] = pelem(4 + q, r);
  o.st_n = pelem(r, 2 * q);
  o.st_t[0] = pelem(2 * q, r);
  o.st_t[1] = pelem(2 * q + 1, r);
  return o;
}
// A operand: logical block (bi, bl) of packed M, k-half h
__device__ __forceinline__ double afrag(const double *M, int bi, int bl, int h, const LaneOfs &o) {
  return (bi >= bl) ? M[ptile(bi, bl) + o.p1[h]] : M[ptile(bl, bi) + o.p2[h]];
}
// B operand: logical block (bl, bj) of packed M, k-half h
__device__ __forceinline__ double bfrag(const double *M, int bl, int bj, int h, const LaneOfs &o) {
  return (bl >= bj) ? M[ptile(bl, bj) + o.p2[h]] : M[ptile(bj, bl) + o.p1[h]];
}
// store / load one accumulator tile of logical block (bi, bj)
__device__ __forceinline__ void store_tile(double *M, int bi, int bj, double c0, double c1, const LaneOfs &o) {
  if (bi >= bj) {
    *reinterpret_cast<double2 *>(M + ptile(bi, bj) + o.st_n) = make_double2(c0, c1);
  } else {
    double *t = M + ptile(bj, bi);
    t[o.st_t[0]] = c0;
    t[o.st_t[1]] = c1;
  }
}
__device__ __forceinline__ void load_tile(const double *M, int bi, int bj, double &c0, double &c1,
                                          const LaneOfs &o) {
  if (bi >= bj) {
    const double2 v = *reinterpret_cast<const double2 *>(M + ptile(bi, bj) + o.st_n);
    c0 = v.x;
    c1 = v.y;
  } else {
    const double *t = M + ptile(bj, bi);
    c0 = t[o.st_t[0]];
    c1 = t[o.st_t[1]];
  }
}

// Operand fragments of one l-step of the circulant half-GEMM: A = X(w, l) (both k-halves) and
// B = W(l, jd) for the warp's H + 1 column blocks.
template <int NB>
struct SymmFrag {
  double a[2];
  double b[(NB + 1) / 2][2];
};
// Running tile offsets (in doubles) of the operands.  Packed tile (bi, bj), bi >= bj, sits at
// (bi (bi + 1) / 2 + bj) * 64, so walking l = 0, 1, ... along block-row x of a symmetric matrix adds
// 64 per step up to the diagonal (stored row x) and (l + 1) * 64 per step below it (stored column x):
// one compare + select + add per operand and step instead of re-deriving the tile address.
template <int NB>
struct SymmWalk {
  int ta, tb[(NB + 1) / 2], j[(NB + 1) / 2];
};
template <int NB>
__device__ __forceinline__ void symm_load(SymmFrag<NB> &f, SymmWalk<NB> &K, const double *X, const double *W,
                                          int w, int l, const LaneOfs &o) {
  {
    const bool col = l > w;   // below the diagonal of block-row w: stored column w, transposed pattern
    const double *t = X + K.ta;
    f.a[0] = t[col ? o.p2[0] : o.p1[0]];
    f.a[1] = t[col ? o.p2[1] : o.p1[1]];
    K.ta += (l < w) ? 64 : (l + 1) * 64;
  }
#pragma unroll
  for (int d = 0; d <= (NB - 1) / 2; ++d) {
    const bool col = l >= K.j[d];
    const double *t = W + K.tb[d];
    f.b[d][0] = t[col ? o.p2[0] : o.p1[0]];
    f.b[d][1] = t[col ? o.p2[1] : o.p1[1]];
    K.tb[d] += (l < K.j[d]) ? 64 : (l + 1) * 64;
  }
}
template <int NB>
__device__ __forceinline__ void symm_mma(double (&acc)[(NB + 1) / 2][2], const SymmFrag<NB> &f) {
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int d = 0; d <= (NB - 1) / 2; ++d) dmma884(acc[d][0], acc[d][1], f.a[h], f.b[d][h]);
}

// acc[d] (+)= sum_l X(w, l) W(l, jd)   for the warp's circulant tiles jd = (w + d) mod NB.
// Register double buffering: the fragments of step l + 1 are in flight while the DMMAs of step l issue.
template <int NB>
__device__ __forceinline__ void symm_gemm(double (&acc)[(NB + 1) / 2][2], const double *X, const double *W,
                                          int w, const LaneOfs &o) {
  constexpr int H = (NB - 1) / 2;
  SymmWalk<NB> K;
  K.ta = (w * (w + 1) / 2) * 64;
#pragma unroll
  for (int d = 0; d <= H; ++d) {
    int j = w + d;
    if (j >= NB) j -= NB;
    K.j[d] = j;
    K.tb[d] = (j * (j + 1) / 2) * 64;
  }
  SymmFrag<NB> f0, f1;
  symm_load<NB>(f0, K, X, W, w, 0, o);
#pragma unroll 1
  for (int l = 0; l < NB - 1; l += 2) {   // NB is odd: pairs (l, l + 1), then the last step
    symm_load<NB>(f1, K, X, W, w, l + 1, o);
    symm_mma<NB>(acc, f0);
    symm_load<NB>(f0, K, X, W, w, l + 2, o);
    symm_mma<NB>(acc, f1);
  }
  symm_mma<NB>(acc, f0);
}

template <int NB>
__device__ __forceinline__ void store_circ(const double (&acc)[(NB + 1) / 2][2], double *M, int w,
                                           const LaneOfs &o) {
#pragma unroll
  for (int d = 0; d <= (NB - 1) / 2; ++d) {
    int j = w + d;
    if (j >= NB) j -= NB;
    store_tile(M, w, j, acc[d][0], acc[d][1], o);
  }
}

// Gram of a staged chunk: raw obs rows Ys[o][m] (row-major, leading dimension LD, nrows4 rows, a
// multiple of 4) with per-row weights wv[o] (0 for padding rows):
//   acc[d] += sum_o wv[o] Ys[o][w-block]^T Ys[o][jd-block]
template <int NB, int LD>
__device__ __forceinline__ void gram_circ(double (&acc)[(NB + 1) / 2][2], const double *Ys, const double *wv,
                                          int nrows4, int w, int lane) {
  constexpr int H = (NB - 1) / 2;
  const int r = lane >> 2, q = lane & 3;
  const double *pa = Ys + (size_t)q * LD + w * 8 + r;
  const double *pb = Ys + (size_t)q * LD + r;
  int jo[H + 1];
#pragma unroll
  for (int d = 0; d <= H; ++d) {
    int j = w + d;
    if (j >= NB) j -= NB;
    jo[d] = j * 8;
  }
#pragma unroll 4
  for (int o = 0; o < nrows4; o += 4) {
    const double a = pa[(size_t)o * LD] * wv[o + q];
    const double *pbk = pb + (size_t)o * LD;
#pragma unroll
    for (int d = 0; d <= H; ++d) dmma884(acc[d][0], acc[d][1], a, pbk[jo[d]]);
  }
}

// Coupled, interval-scaled Newton-Schulz on packed symmetric matrices.  On entry Yp holds
// Y0 = A / s (leading k x k block; identity on the padding rows), c0s = c0 / s the lower eigenvalue
// bound.  On exit Zp holds Z ~= (A/s)^-1/2.  Tp: scratch.  `red`: >= 32 doubles of shared scratch.
// Returns the number of iterations, negative if the residual test was not met within max_iter.
// All NT = 32 NB threads of the CTA must call.
template <int NB>
__device__ __forceinline__ int newton_schulz_invsqrt(double *Yp, double *Zp, double *Tp, double c0s,
                                                     double *red, int max_iter) {
  constexpr int H = (NB - 1) / 2;
  const int lane = threadIdx.x & 31;
  const int w = __shfl_sync(LETKF_FULL_MASK, threadIdx.x >> 5, 0);
  const int r = lane >> 2, q = lane & 3;
  const LaneOfs o = lane_offsets(lane);
  double a = c0s, b = 1.0;   // eigenvalue bracket of M = Z Y
  double acc[H + 1][2], az[H + 1][2];
  int it = 0;
  bool first = true, done = false;
  while (!done) {
    ++it;
    if (first) {   // M = Z0 Y0 = Y0
#pragma unroll
      for (int d = 0; d <= H; ++d) {
        int j = w + d;
        if (j >= NB) j -= NB;
        load_tile(Yp, w, j, acc[d][0], acc[d][1], o);
      }
    } else {
#pragma unroll
      for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
      symm_gemm<NB>(acc, Zp, Yp, w, o);
    }
    // residual ||I - M||_max (the warp's diagonal tile is d = 0)
    double res = 0.0;
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double dg = (d == 0 && r == 2 * q + e) ? 1.0 : 0.0;
        res = fmax(res, fabs(dg - acc[d][e]));
      }
    }
    res = warp_max(res);
    if (lane == 0) red[w] = res;
    __syncthreads();
    res = red[0];
#pragma unroll
    for (int i = 1; i < NB; ++i) res = fmax(res, red[i]);
    const bool last = (res < 1.0e-7) || (it >= max_iter);
    if (!last && !first && res < 2.0e-3) {
      // Close to convergence one higher-order step finishes the job.  With E = I - Z Y (||E|| = res):
      //   res < 2e-4:  Z <- (I + E/2 + 3 E^2/8) Z              residual ~ (5/16) res^3   <= 2.5e-12
      //   res < 2e-3:  Z <- (I + E/2 + 3 E^2/8 + 5 E^3/16) Z   residual ~ (35/128) res^4 <= 4.4e-12
      // i.e. 2 (3) half-GEMMs instead of the 3 + 2 (3 + 3 + 2) of further quadratic steps plus the final one.
      // Y is no longer needed, its storage holds E^2.
      const bool order4 = res >= 2.0e-4;
#pragma unroll
      for (int d = 0; d <= H; ++d) {
#pragma unroll
        for (int e = 0; e < 2; ++e) acc[d][e] = ((d == 0 && r == 2 * q + e) ? 1.0 : 0.0) - acc[d][e];
      }
      store_circ<NB>(acc, Tp, w, o);   // E
      __syncthreads();
#pragma unroll
      for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
      symm_gemm<NB>(az, Tp, Tp, w, o);   // E^2
      if (order4) store_circ<NB>(az, Yp, w, o);
#pragma unroll
      for (int d = 0; d <= H; ++d) {
#pragma unroll
        for (int e = 0; e < 2; ++e)
          az[d][e] = fma(0.375, az[d][e], fma(0.5, acc[d][e], (d == 0 && r == 2 * q + e) ? 1.0 : 0.0));
      }
      __syncthreads();   // E^2 visible; (order 3: all reads of E done)
      if (order4) {
#pragma unroll
        for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
        symm_gemm<NB>(acc, Yp, Tp, w, o);   // E^3
#pragma unroll
        for (int d = 0; d <= H; ++d) {
#pragma unroll
          for (int e = 0; e < 2; ++e) az[d][e] = fma(0.3125, acc[d][e], az[d][e]);
        }
        __syncthreads();   // all reads of E done
      }
      store_circ<NB>(az, Tp, w, o);
      __syncthreads();
#pragma unroll
      for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
      symm_gemm<NB>(az, Tp, Zp, w, o);   // Z' = T Z
      __syncthreads();
      store_circ<NB>(az, Zp, w, o);
      __syncthreads();
      return it;
    }
    double c = 1.0;
    if (!last && (b - a) > 1.0e-3) c = 3.0 / (a + sqrt(a * b) + b);
    const double sc = sqrt(c), h0 = 1.5 * sc, h1 = -0.5 * c * sc;
    // T = sqrt(c) (3 I - c M) / 2
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[d][e] = fma(h1, acc[d][e], (d == 0 && r == 2 * q + e) ? h0 : 0.0);
    }
    if (first) {
      // Z1 = T ; Y1 = T Y0
      store_circ<NB>(acc, Zp, w, o);
      store_circ<NB>(acc, Tp, w, o);
      __syncthreads();
#pragma unroll
      for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
      symm_gemm<NB>(az, Tp, Yp, w, o);
      __syncthreads();   // all reads of Y0 done (also protects red)
      store_circ<NB>(az, Yp, w, o);
      __syncthreads();
    } else {
      store_circ<NB>(acc, Tp, w, o);
      __syncthreads();
#pragma unroll
      for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
      symm_gemm<NB>(az, Tp, Zp, w, o);   // Z' = T Z
      if (!last) {
#pragma unroll
        for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
        symm_gemm<NB>(acc, Tp, Yp, w, o);   // Y' = T Y
      }
      __syncthreads();   // all reads of Z, Y, T done
      store_circ<NB>(az, Zp, w, o);
      if (!last) store_circ<NB>(acc, Yp, w, o);
      __syncthreads();
    }
    const double t = c * a;   // image of the bracket under x -> c x (3 - c x)^2 / 4 is [g(c a), 1]
    a = t * (3.0 - t) * (3.0 - t) * 0.25;
    b = 1.0;
    first = false;
    done = last;
    if (last && !(res < 1.0e-7)) it = -it;
  }
  return it;
}

}  // namespace letkf
