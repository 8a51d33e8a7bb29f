// solver.cuh -- per-CTA dense algebra of the LETKF weight solve, fp64, shared-memory resident.
//
// One CTA solves one grid point (the reference solves one point per OpenMP thread,
// scale/letkf/letkf_tools.f90:319-320, common/common_letkf.f90:52):
//   A  = Yr^T Y + (k-1)/rho I          register-tiled SYRK over obs chunks staged in smem
//   A  = L L^T                         left-looking Cholesky in smem
//   L V = U S                          one-sided (Hestenes) Jacobi on the Cholesky factor
//                                      (Veselic-Hari): A = U S^2 U^T, no V accumulation
// which replaces mtx_eigen / EISPACK rs (common/common_mtx.f90:41, common/netlibrs.f:21).
// With G = U S (orthogonal columns g_j, lambda_j = |g_j|^2) and the known lowest eigenvalue
// c0 = (k-1)/rho (A = c0 I + Yr^T Y), every f(A) is evaluated in shifted form
//   f(A) = f(c0) I + sum_j (f(lambda_j) - f(c0)) g_j g_j^T / lambda_j
//   Pa    : f = 1/lambda                                 (common_letkf.f90:151-157)
//   trans : f = sqrt((k-1)/lambda)                       (common_letkf.f90:199-206)
// so that columns inside the degenerate c0 cluster (p < k) drop out.
// Only f(A) is ever used, so the eigenvector sign / order / basis inside degenerate
// eigenspaces is immaterial (SURVEY.md section 8c "acceptance").
#pragma once
#include "common.cuh"

namespace letkf {

// Compile-time size classes: KC >= k.  NT threads, RJ rows per Jacobi thread, R Gram tiles
// per thread (4x4 tiles of the lower triangle).
template <int KC>
struct SizeClass {
  static constexpr int kPairs = (KC + 1) / 2;
  static constexpr int NT = ((4 * kPairs + 31) / 32) * 32;
  static constexpr int RJ = (KC + 3) / 4;
  static constexpr int NT4 = (KC + 3) / 4;
  static constexpr int kTiles = NT4 * (NT4 + 1) / 2;
  static constexpr int R = (kTiles + NT - 1) / NT;
};

constexpr int kChunk = 32;   // observations staged per Gram chunk

// leading dimension of G (column-major k x 2*ceil(k/2)): ld == 2 (mod 4) makes the quad
// access pattern of the Jacobi (4 lanes x 4 column pairs per 16-lane phase) conflict-free.
__host__ __device__ __forceinline__ int ld_of(int k) {
  int ld = k;
  while ((ld & 3) != 2) ++ld;
  return ld;
}
// row stride of a staged obs chunk: multiple of 4 (16-byte LDS.128 of 4 members)
__host__ __device__ __forceinline__ int ldk_of(int k) { return round_up(k, 4); }

__device__ __forceinline__ void tile_coords(int t, int &ta, int &tb) {
  int a = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (a * (a + 1) / 2 > t) --a;
  while ((a + 1) * (a + 2) / 2 <= t) ++a;
  ta = a;
  tb = t - a * (a + 1) / 2;
}

// acc += sum over `nrows` staged rows of (row[4ta..4ta+3]) x (row[4tb..4tb+3]).
// Rows are pre-scaled by sqrt(weight), so A = sum_i (sqrt(w_i) y_i)(sqrt(w_i) y_i)^T.
template <int R>
__device__ __forceinline__ void gram_accumulate(double (&acc)[R][16], const double *Ys, int nrows,
                                                int ldk, int ntiles) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = threadIdx.x + r * blockDim.x;
    if (t >= ntiles) continue;
    int ta, tb;
    tile_coords(t, ta, tb);
    const double *pa = Ys + 4 * ta, *pb = Ys + 4 * tb;
#pragma unroll 4
    for (int o = 0; o < nrows; ++o) {
      const double2 a01 = *reinterpret_cast<const double2 *>(pa + o * ldk);
      const double2 a23 = *reinterpret_cast<const double2 *>(pa + o * ldk + 2);
      const double2 b01 = *reinterpret_cast<const double2 *>(pb + o * ldk);
      const double2 b23 = *reinterpret_cast<const double2 *>(pb + o * ldk + 2);
      const double a[4] = {a01.x, a01.y, a23.x, a23.y};
      const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][i * 4 + j] = fma(a[i], b[j], acc[r][i * 4 + j]);
    }
  }
}

template <int R>
__device__ __forceinline__ void gram_zero(double (&acc)[R][16]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[r][i] = 0.0;
}

// Write the lower triangle (row >= col) of the accumulated tiles into column-major C
// (leading dimension ldc); `mirror` also fills the upper triangle (symmetric output).
template <int R>
__device__ __forceinline__ void gram_store(const double (&acc)[R][16], double *C, int ldc, int k,
                                           int ntiles, bool mirror) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int t = threadIdx.x + r * blockDim.x;
    if (t >= ntiles) continue;
    int ta, tb;
    tile_coords(t, ta, tb);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = 4 * ta + i, col = 4 * tb + j;
        if (row < k && col < k && row >= col) {
          C[(size_t)col * ldc + row] = acc[r][i * 4 + j];
          if (mirror && row != col) C[(size_t)row * ldc + col] = acc[r][i * 4 + j];
        }
      }
  }
}

// Left-looking Cholesky of the lower triangle of G (k x k, column-major, ld).  On exit the
// lower triangle holds L, the strict upper triangle and the padding column are zero.
// Returns false (to every thread) on a non-positive / non-finite pivot.
__device__ __forceinline__ bool cholesky_lower(double *G, int k, int ld, int ncols, double *piv) {
  const int a = threadIdx.x;
  bool ok = true;
  for (int j = 0; j < k; ++j) {
    double s = 0.0, s2 = 0.0;
    if (a >= j && a < k) {
      const double *ga = G + a, *gj = G + j;
      int c = 0;
      for (; c + 1 < j; c += 2) {
        s = fma(ga[(size_t)c * ld], gj[(size_t)c * ld], s);
        s2 = fma(ga[(size_t)(c + 1) * ld], gj[(size_t)(c + 1) * ld], s2);
      }
      if (c < j) s = fma(ga[(size_t)c * ld], gj[(size_t)c * ld], s);
      s = G[(size_t)j * ld + a] - (s + s2);
      if (a == j) *piv = s;
    }
    __syncthreads();
    const double p = *piv;
    if (!(p > 0.0) || !isfinite(p)) ok = false;
    if (a >= j && a < k) {
      const double d = sqrt(p);
      G[(size_t)j * ld + a] = (a == j) ? d : s / d;
    }
    __syncthreads();
  }
  // zero the strict upper triangle and the padding columns
  for (int idx = threadIdx.x; idx < ncols * k; idx += blockDim.x) {
    const int col = idx / k, row = idx - col * k;
    if (col >= k || row < col) G[(size_t)col * ld + row] = 0.0;
  }
  __syncthreads();
  return ok;
}

// One-sided Jacobi on the columns of G (k rows, 2*m column slots; slot pairs (2j, 2j+1)).
// Round-robin ordering realised by physically moving columns between slots after every
// step, so each thread quad always works on the adjacent slots (2j, 2j+1).
// Returns the number of sweeps used; *converged tells whether the stop rule was met.
template <int RJ>
__device__ __forceinline__ int jacobi_onesided(double *G, int k, int ld, int m, double *red,
                                               int max_sweeps, double c0, bool *converged) {
  const int tid = threadIdx.x;
  const int j = tid >> 2, t = tid & 3;
  const bool act = j < m;
  // destination slots of this pair's (top, bottom) columns for the next step
  int dtop = 2 * j, dbot = 2 * j + 1;
  if (m > 1) {
    if (j == 0) {
      dtop = 0;
      dbot = 2;
    } else {
      dtop = (j == m - 1) ? 2 * (m - 1) + 1 : 2 * (j + 1);
      dbot = 2 * (j - 1) + 1;
    }
  }
  const int nsteps = (m > 1) ? 2 * m - 1 : 1;
  const double *ptop = G + (size_t)(2 * j) * ld, *pbot = G + (size_t)(2 * j + 1) * ld;
  double *qtop = G + (size_t)dtop * ld, *qbot = G + (size_t)dbot * ld;
  int sweep = 0;
  *converged = false;
  while (sweep < max_sweeps) {
    double maxc = 0.0;
    for (int step = 0; step < nsteps; ++step) {
      double gp[RJ], gq[RJ];
      double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
      for (int i = 0; i < RJ; ++i) {
        const int r = t + 4 * i;
        if (act && r < k) {
          gp[i] = ptop[r];
          gq[i] = pbot[r];
        } else {
          gp[i] = 0.0;
          gq[i] = 0.0;
        }
        a = fma(gp[i], gp[i], a);
        b = fma(gq[i], gq[i], b);
        c = fma(gp[i], gq[i], c);
      }
      a += __shfl_xor_sync(LETKF_FULL_MASK, a, 1);
      b += __shfl_xor_sync(LETKF_FULL_MASK, b, 1);
      c += __shfl_xor_sync(LETKF_FULL_MASK, c, 1);
      a += __shfl_xor_sync(LETKF_FULL_MASK, a, 2);
      b += __shfl_xor_sync(LETKF_FULL_MASK, b, 2);
      c += __shfl_xor_sync(LETKF_FULL_MASK, c, 2);
      const double ab = a * b;
      if (ab > 0.0) {
        const double cosang = fabs(c) * rsqrt(ab);
        // Columns that both sit on the known lowest eigenvalue c0 = (k-1)/rho (exactly (k-p)-fold
        // degenerate when p < k) carry weight f(lambda) - f(c0) ~ 0 in the shifted formulas of
        // the consumers, so their mutual angle is irrelevant once it is small; rotating them
        // would only re-diagonalise rounding noise inside the cluster (large-angle rotations,
        // linear convergence).
        const bool cluster = cosang < 1.0e-6 && fabs(a - c0) <= 1.0e-9 * c0 && fabs(b - c0) <= 1.0e-9 * c0;
        if (!cluster) maxc = fmax(maxc, cosang);
        if (!cluster && cosang > 1.0e-15) {
          const double zeta = (b - a) / (2.0 * c);
          const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
          const double cs = rsqrt(fma(tt, tt, 1.0));
          const double sn = cs * tt;
#pragma unroll
          for (int i = 0; i < RJ; ++i) {
            const double p = gp[i], q = gq[i];
            gp[i] = cs * p - sn * q;
            gq[i] = fma(sn, p, cs * q);
          }
        }
      }
      __syncthreads();   // everybody has read its pair
      if (act) {
#pragma unroll
        for (int i = 0; i < RJ; ++i) {
          const int r = t + 4 * i;
          if (r < k) {
            qtop[r] = gp[i];
            qbot[r] = gq[i];
          }
        }
      }
      __syncthreads();
    }
    ++sweep;
    maxc = block_max(maxc, red);
    // quadratic convergence: the rotations of a sweep whose largest |cos| was < 1e-8 leave
    // off-diagonal cosines of order 1e-16 / (relative eigenvalue gap).
    if (maxc < 1.0e-8) {
      *converged = true;
      break;
    }
  }
  return sweep;
}

// lam[j] = |g_j|^2 for the first k column slots holding real columns.  A padding slot (odd
// k) keeps a zero column; `slot_of` is not needed because zero columns get lam = 0 and are
// skipped by every consumer.
__device__ __forceinline__ void column_norms(const double *G, int k, int ld, int ncols, double *lam) {
  for (int j = threadIdx.x; j < ncols; j += blockDim.x) {
    const double *g = G + (size_t)j * ld;
    double s = 0.0, s2 = 0.0;
    int r = 0;
    for (; r + 1 < k; r += 2) {
      s = fma(g[r], g[r], s);
      s2 = fma(g[r + 1], g[r + 1], s2);
    }
    if (r < k) s = fma(g[r], g[r], s);
    lam[j] = s + s2;
  }
  __syncthreads();
}

// T[j][v] = sum_a G[a][j] * X[a][v]     (G^T X)   j < ncols, v < 4*nvb
// X, T are [row][NVP] with v fastest (NVP multiple of 4).  Work item = (j, block of 4 v).
__device__ __forceinline__ void gemm_gt_x(const double *G, int k, int ld, int ncols, const double *X,
                                          double *T, int nvb, int NVP) {
  const int items = ncols * nvb;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int vb = it / ncols, j = it - vb * ncols;
    const double *g = G + (size_t)j * ld;
    const double *x = X + 4 * vb;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int a = 0; a < k; ++a) {
      const double ga = g[a];
      const double2 x01 = *reinterpret_cast<const double2 *>(x + (size_t)a * NVP);
      const double2 x23 = *reinterpret_cast<const double2 *>(x + (size_t)a * NVP + 2);
      t0 = fma(ga, x01.x, t0);
      t1 = fma(ga, x01.y, t1);
      t2 = fma(ga, x23.x, t2);
      t3 = fma(ga, x23.y, t3);
    }
    double *o = T + (size_t)j * NVP + 4 * vb;
    o[0] = t0; o[1] = t1; o[2] = t2; o[3] = t3;
  }
}
// Z[a][v] = sum_j G[a][j] * U[j][v]     (G U)     a < k
__device__ __forceinline__ void gemm_g_u(const double *G, int k, int ld, int ncols, const double *U,
                                         double *Z, int nvb, int NVP) {
  const int items = k * nvb;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int vb = it / k, a = it - vb * k;
    const double *u = U + 4 * vb;
    double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
    for (int j = 0; j < ncols; ++j) {
      const double ga = G[(size_t)j * ld + a];
      const double2 u01 = *reinterpret_cast<const double2 *>(u + (size_t)j * NVP);
      const double2 u23 = *reinterpret_cast<const double2 *>(u + (size_t)j * NVP + 2);
      z0 = fma(ga, u01.x, z0);
      z1 = fma(ga, u01.y, z1);
      z2 = fma(ga, u23.x, z2);
      z3 = fma(ga, u23.y, z3);
    }
    double *o = Z + (size_t)a * NVP + 4 * vb;
    o[0] = z0; o[1] = z1; o[2] = z2; o[3] = z3;
  }
}

}  // namespace letkf
