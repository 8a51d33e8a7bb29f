// ns_solver.cuh -- tensor-core (FP64 DMMA) weight solve of the LETKF analysis, one CTA per grid
// point, replacing mtx_eigen / EISPACK rs (common/common_mtx.f90:41, common/netlibrs.f:21) and the
// four dgemm calls of letkf_core (common/common_letkf.f90:127,156,169,205).
//
// letkf_core only ever uses functions of A = Yr^T Y + (k-1)/rho I:
//     Pa = A^-1,   trans = sqrt(k-1) A^-1/2,   transm = Pa Yr^T d
// so no eigenvectors are needed.  Z = A^-1/2 comes from a Newton-Schulz iteration in PRODUCT FORM:
//     M_0 = A / s,  Z_0 = I;     T_j = sqrt(c_j) (3 I - c_j M_j) / 2;
//     M_{j+1} = T_j M_j T_j,     Z_{j+1} = T_j Z_j                      (invariant M_j = Z_j (A/s) Z_j)
// with s >= lambda_max(A), the eigenvalues of M_j bracketed by [a_j, 1] (a_0 = c0/s with the known lowest
// eigenvalue c0 = (k-1)/rho) and c_j = 3 / (a_j + sqrt(a_j) + 1), the scaling that maps both ends of the
// bracket onto the same image.  M_{j+1} is a function of the single matrix M_j (T_j is a polynomial in
// M_j), so the iteration on M is self-correcting whatever the conditioning of A, and Z only accumulates the
// factors: errors are never amplified.  (The textbook coupled form Z <- T Z, Y <- T Y with M = Z Y needs
// Y <- Y T to be stable; with symmetric storage that cannot be expressed and rounding errors that do not
// commute with A grow by ~sqrt(cond(A))/4 per step -- tools/ns_model.py)
// Once ||I - M||_F < 1.5e-3 a single third- or fourth-order step (Z <- (I + E/2 + 3E^2/8 [+ 5E^3/16]) Z,
// E = I - M) finishes the solve.
//
// All products are GEMMs on the FP64 tensor cores: mma.sync.aligned.m8n8k4.f64 (DMMA; tcgen05/TMEM
// has no FP64 kind).  Every matrix is symmetric, so only one 8x8 tile of each symmetric pair is stored and
// computed: with NB (odd) row blocks, warp w owns the "circulant" tiles (w, (w+d) mod NB), d = 0..H,
// H = (NB-1)/2, which covers every unordered pair exactly once with perfect load balance.  Tile (w, d)
// lives at ((H+1) w + d) * 64 doubles, its elements in FRAGMENT ORDER: (r, c) at (c>>2)*32 + r*4 + (c&3),
// so that a DMMA operand fragment of a stored tile is one bank-conflict-free, lane-contiguous load and the
// fragment of its transpose is conflict-free as well.  Summing the inner index in circulant order
// l = (w+e) mod NB makes every "stored or mirrored?" decision a function of (d, e) only -- resolved at compile
// time in the fully unrolled product, leaving loads with immediate offsets and DMMAs in the inner loop.
#pragma once
#include "common.cuh"

// LETKF_EXP_TRACE builds only (profiling aid): (tag, clock64) pairs of CTA 0 / thread 0, 16000 entries
#ifdef LETKF_EXP_TRACE
namespace letkf {
__device__ long long *g_trace_buf;
__device__ int g_trace_n;
}
#define LETKF_TRACE(tag)                                                                       \
  do {                                                                                         \
    if (threadIdx.x == 0 && blockIdx.x == 0 && letkf::g_trace_buf && letkf::g_trace_n < 16000) { \
      letkf::g_trace_buf[2 * letkf::g_trace_n] = (tag);                                        \
      letkf::g_trace_buf[2 * letkf::g_trace_n + 1] = clock64();                                \
      ++letkf::g_trace_n;                                                                      \
    }                                                                                          \
  } while (0)
#else
#define LETKF_TRACE(tag) do { } while (0)
#endif

namespace letkf {

template <int NB_>
struct NsCfg {
  static_assert(NB_ % 2 == 1, "the circulant tile assignment needs an odd number of row blocks");
  static constexpr int NB = NB_;                   // 8-row blocks
  static constexpr int H = (NB_ - 1) / 2;          // warp w owns tiles (w, w+d mod NB), d = 0..H
  static constexpr int KP = 8 * NB_;               // padded ensemble size (>= k + 2)
  static constexpr int LD = KP + 4;                // row stride of row-major staging / vector blocks
  // Warps of the CTA.  Normally one per row block, owning the H + 1 circulant tiles of that block.  NB = 13 runs 16
  // warps, four per SM sub-partition: with 13 the sub-partitions would issue 28 / 21 / 21 / 21 of the 91 tile products
  // of every GEMM and the first one sets the pace (measured: every DMMA phase of the k = 100 solver ran at the pipe
  // limit of that sub-partition).  Warps 0..11 own their whole row block (compile-time addressing, as everywhere); the
  // seven tiles of row block 12 are shared out to warps 12..15 (2 + 2 + 2 + 1, run-time addressing): 23/23/23/22.
  // The last warp also carries the vector rows of block 12.
#ifdef LETKF_NOSPLIT13   // A/B aid: 13 warps, one per row block
  static constexpr int NW = NB_;
#else
  static constexpr int NW = NB_ == 13 ? 16 : NB_;
#endif
  static constexpr bool SPLIT = NW != NB_;
  static constexpr int NT = 32 * NW;
  static constexpr int NTILE = NB_ * (NB_ + 1) / 2;
  static constexpr int PSZ = NTILE * 64;           // doubles per stored symmetric matrix
  static constexpr int RS = (H + 1) * 64;          // doubles per warp-owned row of tiles
  static constexpr int CR = ((3 * PSZ) / (4 * LD)) & ~3;   // obs rows per staging chunk: FOUR chunk buffers share the 3 PSZ doubles of Y, Z, T
  static constexpr int NSTEP = CR / 4;             // four-row DMMA steps of a full chunk
  static constexpr int RPW = (CR + NW - 1) / NW;   // rows of a chunk one warp requests (row = warp + NW * r, r < RPW)
  // resident CTAs per SM the register allocation is sized for
#ifndef LETKF_MINB7
#define LETKF_MINB7 4
#endif
  static constexpr int MINB = NB_ <= 3 ? 8 : NB_ <= 5 ? 5 : NB_ <= 7 ? LETKF_MINB7 : NB_ <= 9 ? 2 : 1;
};

// The tiles of its row block w that the calling warp owns (accumulates, updates, stores): (w, dbase + s), s < nd,
// held in acc[s].  Without the split: all H + 1 of them, dbase = 0 (everything below folds at compile time).
struct TileOwn {
  int dbase, nd;
};
template <int NB>
__device__ __forceinline__ TileOwn tile_own(int warp) {
  constexpr int H = (NB - 1) / 2;
  if constexpr (!NsCfg<NB>::SPLIT) {
    return TileOwn{0, H + 1};
  } else {
    if (warp < NB - 1) return TileOwn{0, H + 1};
    const int i = warp - (NB - 1);
    return TileOwn{2 * i, (H + 1 - 2 * i) < 2 ? (H + 1 - 2 * i) : 2};
  }
}
template <int NB>
__device__ __forceinline__ bool has(const TileOwn &o, int s) {
  if constexpr (!NsCfg<NB>::SPLIT) return true;
  else return s < o.nd;
}
template <int NB>
__device__ __forceinline__ int dact(const TileOwn &o, int s) {   // circulant offset of the tile in acc[s]
  if constexpr (!NsCfg<NB>::SPLIT) return s;
  else return o.dbase + s;
}
template <int NB>
__device__ __forceinline__ bool owns_row(const TileOwn &o) {    // the whole row block: compile-time addressed code
  if constexpr (!NsCfg<NB>::SPLIT) return true;
  else return o.nd == (NB + 1) / 2;
}
template <int NB>
__device__ __forceinline__ bool carries_vectors(int warp) {      // the warp that holds the vector rows of its row block
  if constexpr (!NsCfg<NB>::SPLIT) return true;
  else return warp < NB - 1 || warp == NsCfg<NB>::NW - 1;
}

// DMMA and the operand loads of the hot loops are volatile asm: the issue order is the source order, which is
// written as an explicit software pipeline (fragments of step e + 1 are requested before the DMMAs of step e issue,
// independent accumulator chains interleaved).  Left to itself ptxas placed every LDS right in front of the
// DMMA that consumes it (one register recycled for all B fragments): latency-bound, 3x slower in the Gram.
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// shared-memory load at a 32-bit shared address + compile-time byte offset (folded into the LDS immediate)
template <int OFF>
__device__ __forceinline__ double lds_imm(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// ---- circulant, fragment-ordered tile storage -------------------------------------------------------
__host__ __device__ __forceinline__ int fpos(int r, int c) { return (c >> 2) * 32 + r * 4 + (c & 3); }
// address of logical element (row, col) of a stored symmetric matrix (element-wise passes only)
template <int NB>
__host__ __device__ __forceinline__ int caddr(int row, int col) {
  constexpr int H = (NB - 1) / 2;
  const int bi = row >> 3, bj = col >> 3;
  int g = bj - bi;
  if (g < 0) g += NB;
  return (g <= H) ? (bi * (H + 1) + g) * 64 + fpos(row & 7, col & 7)
                  : (bj * (H + 1) + (NB - g)) * 64 + fpos(col & 7, row & 7);
}

struct LaneFrag {   // per-lane offsets inside a tile; lane = 4 r + q
  int dir;          // element (r, 4h + q) at dir + 32 h   (operand fragment of the stored tile)
  int trn;          // element (4h + q, r) at trn + 16 h   (operand fragment of its transpose)
  int st;           // accumulator pair (r, 2q), (r, 2q + 1) at st, st + 1
};
__device__ __forceinline__ LaneFrag lane_frag(int lane) {
  const int r = lane >> 2, q = lane & 3;
  LaneFrag o;
  o.dir = lane;
  o.trn = (r >> 2) * 32 + q * 4 + (r & 3);
  o.st = (q >> 1) * 32 + r * 4 + 2 * (q & 1);
  return o;
}

// acc[d] += sum_l X(w, l) W(l, jd)   for the warp's circulant tiles jd = (w + d) mod NB, X and W symmetric.
// l runs in circulant order l = (w + e) mod NB.  With g = (d - e) mod NB:
//   X(w, w+e):    e <= H: stored tile (w, e);                      else transpose of stored tile (w+e, NB-e)
//   W(w+e, w+d):  g <= H: stored tile (w+e, g) [B operand = its "transposed" fragment pattern];
//                 else transpose of stored tile (w+d, NB-g)
template <int NB>
struct SymmAddr {   // 32-bit shared addresses (bytes) of the lane's element in the first tile of block-row (w + x) mod NB
  unsigned xd0;              // X, direct pattern, block-row w
  unsigned xt[NB];           // X, transposed pattern, block-row w + x   (used for x > H)
  unsigned wt[NB];           // W, transposed pattern, block-row w + x
  unsigned wd[(NB + 1) / 2]; // W, direct pattern, block-row w + d
};
template <int NB>
struct SymmFrag {
  double a[2], b[(NB + 1) / 2][2];
};
template <int NB, int E, int D>
__device__ __forceinline__ void symm_load_b(SymmFrag<NB> &f, const SymmAddr<NB> &A);
template <int NB, int E>
__device__ __forceinline__ void symm_load(SymmFrag<NB> &f, const SymmAddr<NB> &A) {
  constexpr int H = (NB - 1) / 2;
  if constexpr (E <= H) {
    f.a[0] = lds_imm<E * 512>(A.xd0);
    f.a[1] = lds_imm<E * 512 + 256>(A.xd0);
  } else {
    f.a[0] = lds_imm<(NB - E) * 512>(A.xt[E]);
    f.a[1] = lds_imm<(NB - E) * 512 + 128>(A.xt[E]);
  }
  symm_load_b<NB, E, 0>(f, A);
}
template <int NB, int E, int D>
__device__ __forceinline__ void symm_load_b(SymmFrag<NB> &f, const SymmAddr<NB> &A) {
  constexpr int H = (NB - 1) / 2, G = (D - E + NB) % NB;
  if constexpr (G <= H) {
    f.b[D][0] = lds_imm<G * 512>(A.wt[E]);
    f.b[D][1] = lds_imm<G * 512 + 128>(A.wt[E]);
  } else {
    f.b[D][0] = lds_imm<(NB - G) * 512>(A.wd[D]);
    f.b[D][1] = lds_imm<(NB - G) * 512 + 256>(A.wd[D]);
  }
  if constexpr (D < H) symm_load_b<NB, E, D + 1>(f, A);
}
template <int NB, int E>
__device__ __forceinline__ void symm_steps(double (&acc)[(NB + 1) / 2][2], SymmFrag<NB> &cur, SymmFrag<NB> &nxt,
                                           const SymmAddr<NB> &A) {
  constexpr int H = (NB - 1) / 2;
  if constexpr (E + 1 < NB) symm_load<NB, E + 1>(nxt, A);
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int d = 0; d <= H; ++d) dmma884(acc[d][0], acc[d][1], cur.a[h], cur.b[d][h]);
  if constexpr (E + 1 < NB) symm_steps<NB, E + 1>(acc, nxt, cur, A);
}
template <int NB>
__device__ __forceinline__ void symm_gemm_row(double (&acc)[(NB + 1) / 2][2], const double *X, const double *W,
                                              int w, const LaneFrag &lf) {
  constexpr int H = (NB - 1) / 2, RS = (H + 1) * 64;
  const unsigned xs = smem_addr(X), ws = smem_addr(W);
  SymmAddr<NB> A;
#pragma unroll
  for (int x = 0; x < NB; ++x) {
    int j = w + x;
    if (j >= NB) j -= NB;
    const unsigned rb = (unsigned)(j * RS);
    if (x == 0) A.xd0 = xs + 8u * (rb + lf.dir);
    A.xt[x] = xs + 8u * (rb + lf.trn);
    A.wt[x] = ws + 8u * (rb + lf.trn);
    if (x <= H) A.wd[x] = ws + 8u * (rb + lf.dir);
  }
  SymmFrag<NB> f0, f1;
  symm_load<NB, 0>(f0, A);
  symm_steps<NB, 0>(acc, f0, f1, A);
}

// A operand fragments of block-row w of a stored symmetric matrix, inner block (w + e) mod NB
template <int NB>
__device__ __forceinline__ void symm_afrag(double (&a)[2], const double *X, int w, int e, int rb_e,
                                           const LaneFrag &lf) {
  constexpr int H = (NB - 1) / 2, RS = (H + 1) * 64;
  if (e <= H) {
    const double *t = X + w * RS + e * 64 + lf.dir;
    a[0] = t[0];
    a[1] = t[32];
  } else {
    const double *t = X + rb_e + (NB - e) * 64 + lf.trn;
    a[0] = t[0];
    a[1] = t[16];
  }
}

// The same product for a warp that owns only the tiles (w, dbase + s), s < nd <= 2 of its row block: run-time
// addressing (the split of NB = 13; these warps issue 4 instead of 14 DMMAs per inner block).
template <int NB>
__device__ __forceinline__ void symm_gemm_part(double (&acc)[(NB + 1) / 2][2], const double *X, const double *W, int w,
                                               const LaneFrag &lf, const TileOwn &own) {
  constexpr int H = (NB - 1) / 2, RS = (H + 1) * 64;
#pragma unroll
  for (int e = 0; e < NB; ++e) {
    int j = w + e;
    if (j >= NB) j -= NB;
    double a[2];
    symm_afrag<NB>(a, X, w, e, j * RS, lf);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      if (s < own.nd) {
        const int d = own.dbase + s;
        int g = d - e, jd = w + d;
        if (g < 0) g += NB;
        if (jd >= NB) jd -= NB;
        const bool stored = g <= H;   // W(w+e, w+d) is the stored tile (w+e, g); else the transpose of tile (w+d, NB-g)
        const double *t = stored ? W + j * RS + g * 64 + lf.trn : W + jd * RS + (NB - g) * 64 + lf.dir;
        const double b0 = t[0], b1 = t[stored ? 16 : 32];
        dmma884(acc[s][0], acc[s][1], a[0], b0);
        dmma884(acc[s][0], acc[s][1], a[1], b1);
      }
    }
  }
}
template <int NB>
__device__ __forceinline__ void symm_gemm(double (&acc)[(NB + 1) / 2][2], const double *X, const double *W,
                                          int w, const LaneFrag &lf, const TileOwn &own) {
  if (owns_row<NB>(own)) symm_gemm_row<NB>(acc, X, W, w, lf);
  else symm_gemm_part<NB>(acc, X, W, w, lf, own);
}

// the warp's tiles <-> registers (accumulator layout)
template <int NB>
__device__ __forceinline__ void store_circ(const double (&acc)[(NB + 1) / 2][2], double *M, int w,
                                           const LaneFrag &lf, const TileOwn &own) {
  constexpr int H = (NB - 1) / 2, RS = (H + 1) * 64;
  double *t = M + w * RS + lf.st;
#pragma unroll
  for (int d = 0; d <= H; ++d)
    if (has<NB>(own, d)) *reinterpret_cast<double2 *>(t + dact<NB>(own, d) * 64) = make_double2(acc[d][0], acc[d][1]);
}
template <int NB>
__device__ __forceinline__ void load_circ(double (&acc)[(NB + 1) / 2][2], const double *M, int w,
                                          const LaneFrag &lf, const TileOwn &own) {
  constexpr int H = (NB - 1) / 2, RS = (H + 1) * 64;
  const double *t = M + w * RS + lf.st;
#pragma unroll
  for (int d = 0; d <= H; ++d) {
    if (!has<NB>(own, d)) continue;
    const double2 v = *reinterpret_cast<const double2 *>(t + dact<NB>(own, d) * 64);
    acc[d][0] = v.x;
    acc[d][1] = v.y;
  }
}

// Gram of a staged chunk: raw obs rows Ys[o][m] (row-major, leading dimension LD, nrows4 rows, a
// multiple of 4) with per-row weights wv[o] (0 for padding rows):
//   acc[s] += sum_o wv[o] Ys[o][w-block]^T Ys[o][jd-block],   jd = w + dbase + s, s < nd
// (run-time loop: the partial last chunk of a list, and every chunk of a warp that owns only part of its row block)
template <int NB, int LD>
__device__ __forceinline__ void gram_circ(double (&acc)[(NB + 1) / 2][2], const double *Ys, const double *wv,
                                          int nrows4, int w, int lane, const TileOwn &own) {
  constexpr int H = (NB - 1) / 2;
  const int r = lane >> 2, q = lane & 3;
  const double *pa = Ys + (size_t)q * LD + w * 8 + r;
  const double *pw = wv + q;
  if (owns_row<NB>(own)) {
    const double *pb[H + 1];
#pragma unroll
    for (int d = 0; d <= H; ++d) {
      int j = w + d;
      if (j >= NB) j -= NB;
      pb[d] = Ys + (size_t)q * LD + j * 8 + r;
    }
#pragma unroll 4
    for (int o = 0; o < nrows4; o += 4) {
      const double a = pa[(size_t)o * LD] * pw[o];
#pragma unroll
      for (int d = 0; d <= H; ++d) dmma884(acc[d][0], acc[d][1], a, pb[d][(size_t)o * LD]);
    }
  } else {
    int j0 = w + own.dbase, j1 = w + own.dbase + 1;
    if (j0 >= NB) j0 -= NB;
    if (j1 >= NB) j1 -= NB;
    const double *pb0 = Ys + (size_t)q * LD + j0 * 8 + r, *pb1 = Ys + (size_t)q * LD + j1 * 8 + r;
    const bool two = own.nd > 1;
#pragma unroll 2
    for (int o = 0; o < nrows4; o += 4) {
      const double a = pa[(size_t)o * LD] * pw[o];
      dmma884(acc[0][0], acc[0][1], a, pb0[(size_t)o * LD]);
      if (two) dmma884(acc[1][0], acc[1][1], a, pb1[(size_t)o * LD]);
    }
  }
}

// The same for a FULL chunk of 4 NSTEP rows and a warp that owns its whole row block, as an explicit software pipeline
// over its four-row steps (operands of step s + 1 requested before the DMMAs of step s issue; all offsets are LDS
// immediates).
template <int NB>
struct GramFrag {
  double a, wt, b[(NB + 1) / 2];
};
template <int NB, int LD, int S, int D>
__device__ __forceinline__ void gram_load_b(GramFrag<NB> &f, const unsigned (&pb)[(NB + 1) / 2]) {
  f.b[D] = lds_imm<S * 4 * LD * 8>(pb[D]);
  if constexpr (D < (NB - 1) / 2) gram_load_b<NB, LD, S, D + 1>(f, pb);
}
template <int NB, int LD, int S>
__device__ __forceinline__ void gram_load(GramFrag<NB> &f, unsigned pa, unsigned pw, const unsigned (&pb)[(NB + 1) / 2]) {
  f.a = lds_imm<S * 4 * LD * 8>(pa);
  f.wt = lds_imm<S * 4 * 8>(pw);
  gram_load_b<NB, LD, S, 0>(f, pb);
}
template <int NB, int LD, int NSTEP, int S>
__device__ __forceinline__ void gram_steps(double (&acc)[(NB + 1) / 2][2], GramFrag<NB> &cur, GramFrag<NB> &nxt, unsigned pa,
                                           unsigned pw, const unsigned (&pb)[(NB + 1) / 2]) {
  constexpr int H = (NB - 1) / 2;
  if constexpr (S + 1 < NSTEP) gram_load<NB, LD, S + 1>(nxt, pa, pw, pb);
  const double a = cur.a * cur.wt;
#pragma unroll
  for (int d = 0; d <= H; ++d) dmma884(acc[d][0], acc[d][1], a, cur.b[d]);
  if constexpr (S + 1 < NSTEP) gram_steps<NB, LD, NSTEP, S + 1>(acc, nxt, cur, pa, pw, pb);
}
template <int NB, int LD, int NSTEP>
__device__ __forceinline__ void gram_circ_full(double (&acc)[(NB + 1) / 2][2], const double *Ys, const double *wv, int w,
                                               int lane) {
  constexpr int H = (NB - 1) / 2;
  const int r = lane >> 2, q = lane & 3;
  const unsigned ys = smem_addr(Ys) + 8u * (unsigned)(q * LD + r);
  const unsigned pa = ys + 64u * (unsigned)w, pw = smem_addr(wv) + 8u * (unsigned)q;
  unsigned pb[H + 1];
#pragma unroll
  for (int d = 0; d <= H; ++d) {
    int j = w + d;
    if (j >= NB) j -= NB;
    pb[d] = ys + 64u * (unsigned)j;
  }
  GramFrag<NB> f0, f1;
  gram_load<NB, LD, 0>(f0, pa, pw, pb);
  gram_steps<NB, LD, NSTEP, 0>(acc, f0, f1, pa, pw, pb);
}

// Frobenius norm^2 contribution of (I - acc) held in the warp's tiles (off-diagonal tiles count twice)
template <int NB>
__device__ __forceinline__ double resid_fro2(const double (&acc)[(NB + 1) / 2][2], int lane, const TileOwn &own) {
  const int r = lane >> 2, q = lane & 3;
  double s = 0.0;
#pragma unroll
  for (int d = 0; d <= (NB - 1) / 2; ++d) {
    if (!has<NB>(own, d)) continue;
    const bool dg = dact<NB>(own, d) == 0;   // the diagonal tile (the others stand for two tiles of the full matrix)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double v = ((dg && r == 2 * q + e) ? 1.0 : 0.0) - acc[d][e];
      s = fma(dg ? v : 2.0 * v, v, s);
    }
  }
  return s;
}

// Product-form, interval-scaled Newton-Schulz on stored symmetric matrices.  On entry `acc` holds the warp's
// tiles of M0 = A / s (leading k x k block; identity on the padding rows) and the same tiles have been stored
// to Mp (no barrier needed in between); c0s = c0 / s is the lower eigenvalue bound.  On exit Zp holds
// Z ~= (A/s)^-1/2 (visible to all threads).  Tp: scratch.  `red`: >= NB doubles of shared scratch.
// Returns the number of iterations, negative if max_iter was reached before the residual test was met.
// All NT = 32 NB threads of the CTA must call.
template <int NB>
__device__ __forceinline__ int newton_schulz_invsqrt(double (&acc)[(NB + 1) / 2][2], double *Mp, double *Zp,
                                                     double *Tp, double c0s, double *red, int max_iter) {
  constexpr int H = (NB - 1) / 2;
  const int lane = threadIdx.x & 31;
  const int wid = __shfl_sync(LETKF_FULL_MASK, threadIdx.x >> 5, 0);   // warp of the CTA
  const int w = wid < NB ? wid : NB - 1;                               // its row block
  const TileOwn own = tile_own<NB>(wid);
  const int r = lane >> 2, q = lane & 3;
  const LaneFrag lf = lane_frag(lane);
  double a = c0s;            // eigenvalues of M in [a, 1]
  double res = 1.0 - a;      // ||I - M||_2 <= 1 - a; later min(bound, measured Frobenius norm)
  double az[H + 1][2];
  int it = 0;
  bool have_z = false;       // Z0 = I is not stored
  for (;;) {
    ++it;
    if (res < 1.5e-3 || it >= max_iter) break;
    double c = 1.0;
    if (1.0 - a > 1.0e-3) c = 3.0 / (a + sqrt(a) + 1.0);
    const double sc = sqrt(c), h0 = 1.5 * sc, h1 = -0.5 * c * sc;
    // T = sqrt(c) (3 I - c M) / 2
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[d][e] = fma(h1, acc[d][e], (dact<NB>(own, d) == 0 && r == 2 * q + e) ? h0 : 0.0);
    }
    store_circ<NB>(acc, Tp, w, lf, own);
    if (!have_z) store_circ<NB>(acc, Zp, w, lf, own);   // Z1 = T0
    __syncthreads();                               // T (and M, Z) visible
#pragma unroll
    for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
    symm_gemm<NB>(acc, Tp, Mp, w, lf, own);             // U = T M
    if (have_z) {
#pragma unroll
      for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
      symm_gemm<NB>(az, Tp, Zp, w, lf, own);            // Z' = T Z
    }
    __syncthreads();                               // all reads of M and Z done
    store_circ<NB>(acc, Mp, w, lf, own);
    if (have_z) store_circ<NB>(az, Zp, w, lf, own);
    __syncthreads();
#pragma unroll
    for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
    symm_gemm<NB>(acc, Mp, Tp, w, lf, own);             // M' = U T
    const double f2 = warp_sum(resid_fro2<NB>(acc, lane, own));
    if (lane == 0) red[wid] = f2;
    __syncthreads();                               // all reads of U and T done; red visible
    store_circ<NB>(acc, Mp, w, lf, own);                // (visible after the next barrier)
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < NsCfg<NB>::NW; ++i) fro2 += red[i];
    const double t = c * a;   // image of the bracket under x -> c x (3 - c x)^2 / 4 is [g(c a), 1]
    a = t * (3.0 - t) * (3.0 - t) * 0.25;
    res = fmin(1.0 - a, sqrt(fro2));
    have_z = true;
  }
  const bool ok = res < 1.5e-3;
  // Finishing step.  With E = I - M (||E||_2 <= res):
  //   res < 1e-8  :  Z <- (I + E/2) Z                          residual ~ (3/8) res^2    < 4e-17
  //   res < 1.5e-4:  Z <- (I + E/2 + 3 E^2/8) Z                residual ~ (5/16) res^3   < 1.1e-12
  //   else        :  Z <- (I + E/2 + 3 E^2/8 + 5 E^3/16) Z     residual ~ (35/128) res^4 < 1.4e-12
  const int order = res < 1.0e-8 ? 2 : res < 1.5e-4 ? 3 : 4;
#pragma unroll
  for (int d = 0; d <= H; ++d) {
#pragma unroll
    for (int e = 0; e < 2; ++e) acc[d][e] = ((dact<NB>(own, d) == 0 && r == 2 * q + e) ? 1.0 : 0.0) - acc[d][e];   // E
  }
  if (order == 2) {
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e) az[d][e] = fma(0.5, acc[d][e], (dact<NB>(own, d) == 0 && r == 2 * q + e) ? 1.0 : 0.0);
    }
  } else {
    store_circ<NB>(acc, Tp, w, lf, own);   // E
    __syncthreads();
#pragma unroll
    for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
    symm_gemm<NB>(az, Tp, Tp, w, lf, own);   // E^2
    if (order == 4) store_circ<NB>(az, Mp, w, lf, own);   // M is dead (nobody reads Mp between the last barrier and here)
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e)
        az[d][e] = fma(0.375, az[d][e], fma(0.5, acc[d][e], (dact<NB>(own, d) == 0 && r == 2 * q + e) ? 1.0 : 0.0));
    }
    __syncthreads();   // E^2 visible; (order 3: all reads of E done)
    if (order == 4) {
#pragma unroll
      for (int d = 0; d <= H; ++d) acc[d][0] = acc[d][1] = 0.0;
      symm_gemm<NB>(acc, Mp, Tp, w, lf, own);   // E^3
#pragma unroll
      for (int d = 0; d <= H; ++d) {
#pragma unroll
        for (int e = 0; e < 2; ++e) az[d][e] = fma(0.3125, acc[d][e], az[d][e]);
      }
      __syncthreads();   // all reads of E done
    }
  }
  if (!have_z) {   // Z = T (A/s was already close to the identity)
    store_circ<NB>(az, Zp, w, lf, own);
    __syncthreads();
  } else {
    store_circ<NB>(az, Tp, w, lf, own);
    __syncthreads();
#pragma unroll
    for (int d = 0; d <= H; ++d) az[d][0] = az[d][1] = 0.0;
    symm_gemm<NB>(az, Tp, Zp, w, lf, own);   // Z' = T Z
    __syncthreads();
    store_circ<NB>(az, Zp, w, lf, own);
    __syncthreads();
  }
  return ok ? it : -it;
}

// ---- the same iteration APPLIED to a block of vectors (the normal path of das_ns_kernel) -----------------------------
// das_letkf never needs Z = (A/s)^-1/2 itself, only Z [dX | b | bd] (at most 16 columns).  Z is the product of the
// factors T_j = h0_j I + h1_j M_j, so the factors are applied to the vector block as they appear,
//     V_j = h0_j V_{j-1} + h1_j (M_j V_{j-1}),
// one skinny product (4 NB DMMAs per warp) instead of the symmetric half-GEMM Z <- T Z ((NB + 1) NB DMMAs), T is never
// stored (U = T M = h0 M + h1 M M, M' = U T = h0 U + h1 U M), and the finishing polynomial is evaluated on the vectors
// in Horner form (order - 1 skinny products instead of E^2, E^3 and T Z).  Per iteration: two half-GEMMs, one skinny
// product, three CTA barriers (product form on Z: three half-GEMMs, four barriers).
//
// Vector blocks are [16][LD] (column c of the block contiguous over the members); a warp owns the 8 members of its
// row block in every column.
template <int LD>
__device__ __forceinline__ void load_vown(double (&v)[2][2], const double *V, int w, int lane) {
  const int r = lane >> 2, q = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) v[nt][e] = V[(size_t)(nt * 8 + 2 * q + e) * LD + w * 8 + r];
}
template <int LD>
__device__ __forceinline__ void store_vown(const double (&v)[2][2], double *V, int w, int lane) {
  const int r = lane >> 2, q = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) V[(size_t)(nt * 8 + 2 * q + e) * LD + w * 8 + r] = v[nt][e];
}
// a2[nt] += Mat(row block w, :) V(:, column tile nt), Mat a stored symmetric matrix
template <int NB, int LD>
__device__ __forceinline__ void skinny_mv(double (&a2)[2][2], const double *Mat, const double *V, int w,
                                          const LaneFrag &lf, int lane) {
  constexpr int RS = ((NB - 1) / 2 + 1) * 64;
  const int r = lane >> 2, q = lane & 3;
  const double *pb = V + (size_t)r * LD + q;
  double a[2], an[2];
  symm_afrag<NB>(a, Mat, w, 0, w * RS, lf);
#pragma unroll
  for (int e = 0; e < NB; ++e) {
    int j = w + e;
    if (j >= NB) j -= NB;
    if (e + 1 < NB) {   // A fragments one step ahead
      int jn = j + 1;
      if (jn >= NB) jn -= NB;
      symm_afrag<NB>(an, Mat, w, e + 1, jn * RS, lf);
    }
    const double b00 = pb[j * 8], b01 = pb[j * 8 + 4], b10 = pb[(size_t)8 * LD + j * 8], b11 = pb[(size_t)8 * LD + j * 8 + 4];
    dmma884(a2[0][0], a2[0][1], a[0], b00);
    dmma884(a2[1][0], a2[1][1], a[0], b10);
    dmma884(a2[0][0], a2[0][1], a[1], b01);
    dmma884(a2[1][0], a2[1][1], a[1], b11);
    a[0] = an[0];
    a[1] = an[1];
  }
}

// On entry `acc` holds the warp's tiles of M0 = A / s, also stored to Ma; V0: the vectors.  On exit V = (A/s)^-1/2 V0
// (visible to all threads).  Mb: matrix scratch; Eb: where the finishing step stores E = I - M (may be Mb); tmp: one
// vector block of scratch (may overlay Ma, and Mb when Eb != Mb).  V0 is only read.  All threads of the CTA must call.
// Returns the number of iterations, negative if max_iter was reached before the residual test was met.
template <int NB, int LD>
__device__ __forceinline__ int newton_schulz_apply(double (&acc)[(NB + 1) / 2][2], double *Ma, double *Mb, const double *V0,
                                                   double *V, double *Eb, double *tmp, double c0s, double *red,
                                                   int max_iter) {
  constexpr int H = (NB - 1) / 2;
  const int lane = threadIdx.x & 31;
  const int wid = __shfl_sync(LETKF_FULL_MASK, threadIdx.x >> 5, 0);   // warp of the CTA
  const int w = wid < NB ? wid : NB - 1;                               // its row block
  const TileOwn own = tile_own<NB>(wid);
  const int r = lane >> 2, q = lane & 3;
  const LaneFrag lf = lane_frag(lane);
  double a = c0s;            // eigenvalues of M in [a, 1]
  double res = 1.0 - a;      // ||I - M||_2 <= 1 - a; later min(bound, measured Frobenius norm)
  double au[H + 1][2];
  const bool rowown = carries_vectors<NB>(wid);   // this warp also carries the vector rows of its row block
  const double *Vs = V0;
  int it = 0;
  __syncthreads();           // M0 (stored by the caller) visible: the first product reads every tile
  LETKF_TRACE(20);
  for (;;) {
    ++it;
    if (res < 1.5e-3 || it >= max_iter) break;
    double c = 1.0;
    if (1.0 - a > 1.0e-3) c = 3.0 / (a + sqrt(a) + 1.0);
    const double sc = sqrt(c), h0 = 1.5 * sc, h1 = -0.5 * c * sc;   // T = h0 I + h1 M
#pragma unroll
    for (int d = 0; d <= H; ++d) au[d][0] = au[d][1] = 0.0;
    symm_gemm<NB>(au, Ma, Ma, w, lf, own);                // M M
    double pv[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, vo[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    if (rowown) {
      skinny_mv<NB, LD>(pv, Ma, Vs, w, lf, lane);    // M V
      load_vown<LD>(vo, Vs, w, lane);
    }
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[d][e] = fma(h1, au[d][e], h0 * acc[d][e]);   // U = T M
    }
    store_circ<NB>(acc, Mb, w, lf, own);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) pv[nt][e] = fma(h1, pv[nt][e], h0 * vo[nt][e]);   // V' = T V
    LETKF_TRACE(21);
    __syncthreads();                                 // U visible; all reads of V done
    LETKF_TRACE(22);
    if (rowown) store_vown<LD>(pv, V, w, lane);
#pragma unroll
    for (int d = 0; d <= H; ++d) au[d][0] = au[d][1] = 0.0;
    symm_gemm<NB>(au, Mb, Ma, w, lf, own);                // U M
#pragma unroll
    for (int d = 0; d <= H; ++d) {
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[d][e] = fma(h1, au[d][e], h0 * acc[d][e]);   // M' = U T
    }
    const double f2 = warp_sum(resid_fro2<NB>(acc, lane, own));
    if (lane == 0) red[wid] = f2;
    LETKF_TRACE(23);
    __syncthreads();                                 // all reads of M and U done; red visible
    LETKF_TRACE(24);
    store_circ<NB>(acc, Ma, w, lf, own);
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < NsCfg<NB>::NW; ++i) fro2 += red[i];
    __syncthreads();                                 // M' and V' visible (red may be rewritten)
    LETKF_TRACE(25);
    const double t = c * a;   // image of the bracket under x -> c x (3 - c x)^2 / 4 is [g(c a), 1]
    a = t * (3.0 - t) * (3.0 - t) * 0.25;
    res = fmin(1.0 - a, sqrt(fro2));
    Vs = V;
  }
  const bool ok = res < 1.5e-3;
  // Finishing step (same orders and thresholds as newton_schulz_invsqrt), Horner form on the vectors:
  //   V <- Vs + E (1/2 Vs + E (3/8 Vs + E (5/16 Vs)))
  const int order = res < 1.0e-8 ? 2 : res < 1.5e-4 ? 3 : 4;
#pragma unroll
  for (int d = 0; d <= H; ++d) {
#pragma unroll
    for (int e = 0; e < 2; ++e) acc[d][e] = ((dact<NB>(own, d) == 0 && r == 2 * q + e) ? 1.0 : 0.0) - acc[d][e];   // E
  }
  store_circ<NB>(acc, Eb, w, lf, own);
  __syncthreads();
  LETKF_TRACE(26);
  const double *ts = Vs;
  for (int jj = order - 1; jj >= 1; --jj) {
    double pv[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, vo[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    if (rowown) {
      skinny_mv<NB, LD>(pv, Eb, ts, w, lf, lane);
      load_vown<LD>(vo, Vs, w, lane);
    }
    const double cm = jj == 1 ? 1.0 : jj == 2 ? 0.5 : 0.375;                           // coefficient of Vs at this level
    const double sp = jj != order - 1 ? 1.0 : jj == 1 ? 0.5 : jj == 2 ? 0.375 : 0.3125;   // E (c Vs) = c (E Vs)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) pv[nt][e] = fma(sp, pv[nt][e], cm * vo[nt][e]);
    double *dst = jj == 1 ? V : tmp;
    if (dst == ts) __syncthreads();   // in place: every warp has read the block
    if (rowown) store_vown<LD>(pv, dst, w, lane);
    __syncthreads();
    ts = tmp;
  }
  return ok ? it : -it;
}

}  // namespace letkf
