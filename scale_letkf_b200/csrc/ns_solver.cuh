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
// matrices commute and the iteration is numerically stable and quadratically convergent, even
// with the (k-p)-fold degenerate eigenvalue c0 (p < k).  6-7 iterations for cond(A) <= 100.
//
// All products are k x k x k GEMMs on the FP64 tensor cores: mma.sync.aligned.m8n8k4.f64 (DMMA;
// tcgen05/TMEM has no FP64 kind).  Warp w owns the 8-row block w of every product; operands are
// read from shared memory with leading dimension LD = KP + 4 (== 4 or 12 mod 16 doubles), which
// makes both fragment patterns bank-conflict free; T stays in registers and is converted from the
// accumulator layout to the A-operand layout with warp shuffles, so only Y and Z live in smem.
#pragma once
#include "common.cuh"

namespace letkf {

template <int NB_>
struct NsCfg {
  static constexpr int NB = NB_;          // 8-row blocks
  static constexpr int KP = 8 * NB_;      // padded ensemble size
  static constexpr int LD = KP + 4;       // leading dimension of every smem matrix
  static constexpr int NT = 32 * NB_;     // one warp per row block
  // resident CTAs per SM the register allocation is sized for (shared memory allows as many)
  static constexpr int MINB = NB_ <= 3 ? 6 : NB_ <= 5 ? 4 : NB_ <= 7 ? 3 : NB_ <= 8 ? 2 : 1;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// accumulator tile (lane holds C[r][2q], C[r][2q+1]; r = lane>>2, q = lane&3) -> A fragment of
// its k-step h (columns 4h..4h+3): A[r][4h + q]
__device__ __forceinline__ double acc_to_afrag(double c0, double c1, int h, int lane) {
  const int src = (lane & ~3) | (2 * h + ((lane & 3) >> 1));
  const double v0 = __shfl_sync(LETKF_FULL_MASK, c0, src);
  const double v1 = __shfl_sync(LETKF_FULL_MASK, c1, src);
  return (lane & 1) ? v1 : v0;
}

// acc[j] += sum over k-blocks kb in [0, nkb) of  Afrag(kb) * B[kb*4.., j*8..]
// A from a row-major smem matrix (rows w*8..): A[m][kk] = Am[m*LD + kk]
// B from a row-major smem matrix:              B[kk][n] = Bm[kk*LD + n]
template <int NB, int LD>
__device__ __forceinline__ void gemm_rowblock_ss(double (&acc)[NB][2], const double *Am, const double *Bm,
                                                 int w, int nkb, int lane) {
  const int r = lane >> 2, q = lane & 3;
  const double *pa = Am + (size_t)(w * 8 + r) * LD + q;
  const double *pb = Bm + (size_t)q * LD + r;
#pragma unroll 2
  for (int kb = 0; kb < nkb; ++kb) {
    const double a = pa[kb * 4];
    const double *pbk = pb + (size_t)kb * 4 * LD;
#pragma unroll
    for (int j = 0; j < NB; ++j) dmma884(acc[j][0], acc[j][1], a, pbk[j * 8]);
  }
}

// acc[j] += T(w, :) * B   with the row block of T held in registers (accumulator layout)
template <int NB, int LD>
__device__ __forceinline__ void gemm_rowblock_rs(double (&acc)[NB][2], const double (&T)[NB][2],
                                                 const double *Bm, int lane) {
  const int r = lane >> 2, q = lane & 3;
  const double *pb = Bm + (size_t)q * LD + r;
#pragma unroll
  for (int l = 0; l < NB; ++l) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double a = acc_to_afrag(T[l][0], T[l][1], h, lane);
      const double *pbk = pb + (size_t)(l * 8 + h * 4) * LD;
#pragma unroll
      for (int j = 0; j < NB; ++j) dmma884(acc[j][0], acc[j][1], a, pbk[j * 8]);
    }
  }
}

// Gram of a staged chunk: rows o of Ys (row-major [o][m], leading dimension LD, `nrows4` rows, a
// multiple of 4, zero padded) -> acc(w-block, :) += Ys^T Ys
template <int NB, int LD>
__device__ __forceinline__ void gram_rowblock(double (&acc)[NB][2], const double *Ys, int nrows4, int w,
                                              int lane) {
  const int r = lane >> 2, q = lane & 3;
  const double *pa = Ys + (size_t)q * LD + w * 8 + r;
  const double *pb = Ys + (size_t)q * LD + r;
#pragma unroll 2
  for (int o = 0; o < nrows4; o += 4) {
    const double a = pa[(size_t)o * LD];
    const double *pbk = pb + (size_t)o * LD;
#pragma unroll
    for (int j = 0; j < NB; ++j) dmma884(acc[j][0], acc[j][1], a, pbk[j * 8]);
  }
}

// store the row block held in accumulator layout: element (w*8 + r, j*8 + 2q + e) -> M[row*LD + col]
template <int NB, int LD>
__device__ __forceinline__ void store_rowblock(const double (&acc)[NB][2], double *M, int w, int lane) {
  const int r = lane >> 2, q = lane & 3;
  double *p = M + (size_t)(w * 8 + r) * LD + 2 * q;
#pragma unroll
  for (int j = 0; j < NB; ++j) *reinterpret_cast<double2 *>(p + j * 8) = make_double2(acc[j][0], acc[j][1]);
}

template <int NB>
__device__ __forceinline__ void zero_rowblock(double (&acc)[NB][2]) {
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j][0] = acc[j][1] = 0.0;
}

// Coupled, interval-scaled Newton-Schulz.  On entry warp w holds the row block w of A (including
// the (k-1)/rho diagonal; rows/cols >= k are zero) in `acc`.  On exit Zb holds Z ~= sqrt(s) A^-1/2
// (row-major, LD) for the leading k x k block and *s_out = s.  Ybuf/Zbuf: KP x LD doubles each.
// `red`: >= 32 doubles of shared scratch.  Returns the number of iterations, negative if the
// iteration did not converge within max_iter.  All threads of the CTA (NT = 32 NB) must call.
template <int NB>
__device__ __forceinline__ int newton_schulz_invsqrt(double (&acc)[NB][2], double *Yb, double *Zb, int k,
                                                     double c0, double *red, int max_iter, double *s_out) {
  using C = NsCfg<NB>;
  constexpr int LD = C::LD;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = lane >> 2, q = lane & 3;
  // ---- s = ||A||_1 (max absolute row sum; A symmetric) -----------------------------------------
  double rs = 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) rs += fabs(acc[j][0]) + fabs(acc[j][1]);
  rs += __shfl_xor_sync(LETKF_FULL_MASK, rs, 1);
  rs += __shfl_xor_sync(LETKF_FULL_MASK, rs, 2);
  rs = fmax(rs, __shfl_xor_sync(LETKF_FULL_MASK, rs, 4));
  rs = fmax(rs, __shfl_xor_sync(LETKF_FULL_MASK, rs, 8));
  rs = fmax(rs, __shfl_xor_sync(LETKF_FULL_MASK, rs, 16));
  __syncthreads();
  if (lane == 0) red[w] = rs;
  __syncthreads();
  double s = red[0];
#pragma unroll
  for (int i = 1; i < NB; ++i) s = fmax(s, red[i]);
  *s_out = s;
  const double is = 1.0 / s;
  // ---- Y0 = A / s (padding rows: identity), M = Z0 Y0 = Y0 ---------------------------------------
  const int row = w * 8 + r;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = j * 8 + 2 * q + e;
      double v = acc[j][e] * is;
      if (row >= k || col >= k) v = (row == col) ? 1.0 : 0.0;
      acc[j][e] = v;
    }
  }
  store_rowblock<NB, LD>(acc, Yb, w, lane);
  double a = c0 * is, b = 1.0;   // eigenvalue bracket of M
  double T[NB][2];
  int it = 0;
  bool first = true, done = false;
  while (!done) {
    ++it;
    if (!first) {   // M = Z Y
      zero_rowblock<NB>(acc);
      gemm_rowblock_ss<NB, LD>(acc, Zb, Yb, w, 2 * NB, lane);
    }
    // residual ||I - M||_max and the scaled T = sqrt(c) (3 I - c M) / 2
    double res = 0.0;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = j * 8 + 2 * q + e;
        res = fmax(res, fabs(((row == col) ? 1.0 : 0.0) - acc[j][e]));
      }
    }
    res = warp_max(res);
    __syncthreads();   // every warp is done reading Y, Z for M
    if (lane == 0) red[w] = res;
    __syncthreads();
    res = red[0];
#pragma unroll
    for (int i = 1; i < NB; ++i) res = fmax(res, red[i]);
    const bool last = (res < 1.0e-7) || (it >= max_iter);
    double c = 1.0;
    if (!last && (b - a) > 1.0e-3) c = 3.0 / (a + sqrt(a * b) + b);
    const double sc = sqrt(c), h0 = 1.5 * sc, h1 = -0.5 * c * sc;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = j * 8 + 2 * q + e;
        T[j][e] = fma(h1, acc[j][e], (row == col) ? h0 : 0.0);
      }
    }
    if (first) {
      // Z1 = T, Y1 = T Y0
      zero_rowblock<NB>(acc);
      gemm_rowblock_rs<NB, LD>(acc, T, Yb, lane);
      store_rowblock<NB, LD>(T, Zb, w, lane);
      __syncthreads();   // all reads of Y0 done
      store_rowblock<NB, LD>(acc, Yb, w, lane);
      __syncthreads();
    } else {
      zero_rowblock<NB>(acc);
      gemm_rowblock_rs<NB, LD>(acc, T, Zb, lane);   // Z' = T Z
      __syncthreads();                               // all reads of Z done
      store_rowblock<NB, LD>(acc, Zb, w, lane);
      if (!last) {
        zero_rowblock<NB>(acc);
        gemm_rowblock_rs<NB, LD>(acc, T, Yb, lane);   // Y' = T Y
        __syncthreads();
        store_rowblock<NB, LD>(acc, Yb, w, lane);
      }
      __syncthreads();
    }
    if (c != 1.0) {
      const double t = c * a;
      a = t * (3.0 - t) * (3.0 - t) * 0.25;
      b = 1.0;
    } else {
      // unscaled step: x -> x (3 - x)^2 / 4 keeps [a, 1]
      a = a * (3.0 - a) * (3.0 - a) * 0.25;
    }
    first = false;
    done = last;
    if (last && !(res < 1.0e-7)) it = -it;
  }
  return it;
}

}  // namespace letkf
