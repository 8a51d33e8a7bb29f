// oracle_core.cpp -- CPU restatement of letkf_core and the numerics below it.
//
// TEST INFRASTRUCTURE ONLY (see letkf_oracle.h).  PARITY UNPINNED by the reference:
// no golden vectors exist and the Fortran cannot be built here; pinned by KATs,
// invariants and a LAPACK cross-check in tests/.
//
// Build: g++ -O2 -ffp-contract=off -fopenmp (no FMA contraction, like gfortran on
// x86-64 without -march flags).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "letkf_oracle.h"

namespace {
inline double dsign(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }
}  // namespace

extern "C" {

// common/netlibrs.f:1-20 -- sqrt(a^2+b^2) by the Moler-Morrison iteration.
double oracle_pythag(double a, double b) {
  double p = std::max(std::fabs(a), std::fabs(b));
  if (p == 0.0) return p;
  if (!(a == a) || !(b == b)) return a + b;   // NaN in: the reference's loop below would never end; the checker must
  double q = std::min(std::fabs(a), std::fabs(b)) / p;
  double r = q * q;
  for (;;) {
    double t = 4.0 + r;
    if (t == 4.0) break;
    double s = r / t;
    double u = 1.0 + 2.0 * s;
    p = u * p;
    double su = s / u;
    r = su * su * r;
  }
  return p;
}

// common/netlibrs.f:520-683 -- Householder tridiagonalisation accumulating Q in z.
// 1-based accessors keep the index algebra of the published algorithm checkable.
void oracle_tred2(int nm, int n, const double *a, double *d_, double *e_, double *z_) {
#define Z(i, j) z_[((i)-1) + (size_t)((j)-1) * nm]
#define AA(i, j) a[((i)-1) + (size_t)((j)-1) * nm]
#define D(i) d_[(i)-1]
#define E(i) e_[(i)-1]
  for (int i = 1; i <= n; ++i) {
    for (int j = i; j <= n; ++j) Z(j, i) = AA(j, i);   // only the lower triangle is read
    D(i) = AA(n, i);
  }
  if (n > 1) {
    for (int i = n; i >= 2; --i) {
      const int l = i - 1;
      double h = 0.0, scale = 0.0;
      bool skip = (l < 2);
      if (!skip) {
        for (int k = 1; k <= l; ++k) scale += std::fabs(D(k));
        skip = (scale == 0.0);
      }
      if (skip) {
        E(i) = D(l);
        for (int j = 1; j <= l; ++j) {
          D(j) = Z(l, j);
          Z(i, j) = 0.0;
          Z(j, i) = 0.0;
        }
      } else {
        for (int k = 1; k <= l; ++k) {
          D(k) /= scale;
          h += D(k) * D(k);
        }
        double f = D(l);
        double g = -dsign(std::sqrt(h), f);
        E(i) = scale * g;
        h -= f * g;
        D(l) = f - g;
        for (int j = 1; j <= l; ++j) E(j) = 0.0;
        for (int j = 1; j <= l; ++j) {   // form A*u with fused dot + axpy down column j
          f = D(j);
          Z(j, i) = f;
          g = E(j) + Z(j, j) * f;
          for (int k = j + 1; k <= l; ++k) {
            g += Z(k, j) * D(k);
            E(k) += Z(k, j) * f;
          }
          E(j) = g;
        }
        f = 0.0;
        for (int j = 1; j <= l; ++j) {   // form p
          E(j) /= h;
          f += E(j) * D(j);
        }
        const double hh = f / (h + h);
        for (int j = 1; j <= l; ++j) E(j) -= hh * D(j);   // form q
        for (int j = 1; j <= l; ++j) {   // rank-2 update of the reduced matrix
          f = D(j);
          g = E(j);
          for (int k = j; k <= l; ++k) Z(k, j) = Z(k, j) - f * E(k) - g * D(k);
          D(j) = Z(l, j);
          Z(i, j) = 0.0;
        }
      }
      D(i) = h;
    }
    for (int i = 2; i <= n; ++i) {   // accumulate the transformations
      const int l = i - 1;
      Z(n, l) = Z(l, l);
      Z(l, l) = 1.0;
      const double h = D(i);
      if (h != 0.0) {
        for (int k = 1; k <= l; ++k) D(k) = Z(k, i) / h;
        for (int j = 1; j <= l; ++j) {
          double g = 0.0;
          for (int k = 1; k <= l; ++k) g += Z(k, i) * Z(k, j);
          for (int k = 1; k <= l; ++k) Z(k, j) -= g * D(k);
        }
      }
      for (int k = 1; k <= l; ++k) Z(k, i) = 0.0;
    }
  }
  for (int i = 1; i <= n; ++i) {
    D(i) = Z(n, i);
    Z(n, i) = 0.0;
  }
  Z(n, n) = 1.0;
  E(1) = 0.0;
#undef AA
}

// common/netlibrs.f:215-384 -- implicit-shift QL with eigenvectors; <=30 iterations per
// eigenvalue (:300); ascending selection sort of (d, z) at the end (:356-377).
int oracle_tql2(int nm, int n, double *d_, double *e_, double *z_) {
  if (n == 1) return 0;
  for (int i = 2; i <= n; ++i) E(i - 1) = E(i);
  double f = 0.0, tst1 = 0.0;
  E(n) = 0.0;
  for (int l = 1; l <= n; ++l) {
    int j = 0;
    double h = std::fabs(D(l)) + std::fabs(E(l));
    if (tst1 < h) tst1 = h;
    int m;
    for (m = l; m <= n; ++m) {
      double tst2 = tst1 + std::fabs(E(m));
      if (tst2 == tst1) break;   // e(n) == 0 guarantees termination
    }
    if (m != l) {
      for (;;) {
        if (j == 30) return l;
        ++j;
        const int l1 = l + 1, l2 = l1 + 1;
        double g = D(l);
        double p = (D(l1) - g) / (2.0 * E(l));
        double r = oracle_pythag(p, 1.0);
        D(l) = E(l) / (p + dsign(r, p));
        D(l1) = E(l) * (p + dsign(r, p));
        const double dl1 = D(l1);
        h = g - D(l);
        for (int i = l2; i <= n; ++i) D(i) -= h;
        f += h;
        // QL sweep from m-1 down to l
        p = D(m);
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = E(l1);
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * E(i);
          h = c * p;
          r = oracle_pythag(p, E(i));
          E(i + 1) = s * r;
          s = E(i) / r;
          c = p / r;
          p = c * D(i) - s * g;
          D(i + 1) = h + s * (c * g + s * D(i));
          for (int k = 1; k <= n; ++k) {   // rotate the eigenvector columns i, i+1
            h = Z(k, i + 1);
            Z(k, i + 1) = s * Z(k, i) + c * h;
            Z(k, i) = c * Z(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * E(l) / dl1;
        E(l) = s * p;
        D(l) = c * p;
        double tst2 = tst1 + std::fabs(E(l));
        if (!(tst2 > tst1)) break;
      }
    }
    D(l) = D(l) + f;
  }
  for (int ii = 2; ii <= n; ++ii) {   // order eigenvalues and eigenvectors
    const int i = ii - 1;
    int k = i;
    double p = D(i);
    for (int j = ii; j <= n; ++j) {
      if (D(j) < p) {
        k = j;
        p = D(j);
      }
    }
    if (k != i) {
      D(k) = D(i);
      D(i) = p;
      for (int j = 1; j <= n; ++j) std::swap(Z(j, i), Z(j, k));
    }
  }
  return 0;
#undef Z
#undef D
#undef E
}

// common/netlibrs.f:21-79 with matz != 0: tred2 then tql2; w ascending.
int oracle_rs(int nm, int n, const double *a, double *w, double *z) {
  if (n > nm) return 10 * n;
  std::vector<double> fv1(n);
  oracle_tred2(nm, n, a, w, fv1.data(), z);
  return oracle_tql2(nm, n, w, fv1.data(), z);
}

// common/common_mtx.f90:41-99.  Eigenvalues returned in DESCENDING order; values below
// |lambda_max|*sqrt(eps) are zeroed together with their vectors (:66-74).  The reorder
// branch (:81-91) writes eivec instead of eivec8 in the reference (a latent defect); it is
// unreachable for the LETKF matrix (A >= (k-1)/rho I) and is restated as written.
int oracle_mtx_eigen(int n, const double *a, double *eival, double *eivec) {
  std::vector<double> a8(a, a + (size_t)n * n), eival8(n), eivec8((size_t)n * n, 0.0);
  int ierr = oracle_rs(n, n, a8.data(), eival8.data(), eivec8.data());
  if (ierr != 0) return -1;
  int nrank_eff = n;
  if (eival8[n - 1] > 0) {
    const double thr = std::fabs(eival8[n - 1]) * std::sqrt(std::numeric_limits<double>::epsilon());
    for (int i = 0; i < n; ++i) {
      if (eival8[i] < thr) {
        --nrank_eff;
        eival8[i] = 0.0;
        std::fill(eivec8.begin() + (size_t)i * n, eivec8.begin() + (size_t)(i + 1) * n, 0.0);
      }
    }
  } else {
    return -2;
  }
  if (nrank_eff < n && eival8[0] != 0) {
    int j = 0;
    for (int i = n; i >= 1; --i) {
      if (eival8[i - 1] == 0) {
        const int src = n - nrank_eff - j;   // 1-based
        eival8[i - 1] = eival8[src - 1];
        std::copy(eivec8.begin() + (size_t)(src - 1) * n, eivec8.begin() + (size_t)src * n,
                  eivec + (size_t)(i - 1) * n);
        eival8[src - 1] = 0.0;
        std::fill(eivec8.begin() + (size_t)(src - 1) * n, eivec8.begin() + (size_t)src * n, 0.0);
        ++j;
      }
    }
  }
  for (int i = 1; i <= n; ++i) {
    eival[i - 1] = eival8[n - i];
    std::copy(eivec8.begin() + (size_t)(n - i) * n, eivec8.begin() + (size_t)(n - i + 1) * n,
              eivec + (size_t)(i - 1) * n);
  }
  return nrank_eff;
}

}  // extern "C"

namespace {
// common/netlibblas.f:493-505 -- C(m,n) = A(kk,m)^T * B(kk,n): dot-product form.
void dgemm_tn(int m, int n, int kk, const double *A, int lda, const double *B, int ldb, double *C,
              int ldc) {
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < m; ++i) {
      double temp = 0.0;
      for (int l = 0; l < kk; ++l) temp += A[l + (size_t)i * lda] * B[l + (size_t)j * ldb];
      C[i + (size_t)j * ldc] = temp;
    }
}
// common/netlibblas.f:512-530 -- C(m,n) = A(m,kk) * B(n,kk)^T: axpy form, skips B(j,l)==0.
void dgemm_nt(int m, int n, int kk, const double *A, int lda, const double *B, int ldb, double *C,
              int ldc) {
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) C[i + (size_t)j * ldc] = 0.0;
    for (int l = 0; l < kk; ++l) {
      const double b = B[j + (size_t)l * ldb];
      if (b != 0.0) {
        for (int i = 0; i < m; ++i) C[i + (size_t)j * ldc] += b * A[i + (size_t)l * lda];
      }
    }
  }
}
}  // namespace

extern "C" {

// common/common_letkf.f90:52-257.
int oracle_letkf_core(int ne, int nobs, int nobsl, const double *hdxb, const double *rdiag,
                      const double *rloc, const double *dep, double *parm_infl, double *trans,
                      double *transm, double *pao, int rdiag_wloc, int infl_update,
                      const double *depd, double *transmd) {
  const size_t k2 = (size_t)ne * ne;
  if (nobsl == 0) {   // :89-107
    std::fill(trans, trans + k2, 0.0);
    for (int i = 0; i < ne; ++i) trans[i + (size_t)i * ne] = std::sqrt(*parm_infl);
    if (transm) std::fill(transm, transm + ne, 0.0);
    if (transmd) std::fill(transmd, transmd + ne, 0.0);
    if (pao) {
      std::fill(pao, pao + k2, 0.0);
      for (int i = 0; i < ne; ++i) pao[i + (size_t)i * ne] = *parm_infl / (double)(ne - 1);
    }
    return 0;
  }
  std::vector<double> hdxb_rinv((size_t)nobsl * ne), hdxb_c((size_t)nobsl * ne), eivec(k2),
      eival(ne), pa(k2), work1(k2), work2((size_t)ne * nobsl), work3(ne);
  // hdxb Rinv (:111-123); hdxb(1:nobsl,:) section copy (:127)
  for (int j = 0; j < ne; ++j)
    for (int i = 0; i < nobsl; ++i) {
      const double h = hdxb[i + (size_t)j * nobs];
      hdxb_c[i + (size_t)j * nobsl] = h;
      hdxb_rinv[i + (size_t)j * nobsl] = rdiag_wloc ? h / rdiag[i] : h / rdiag[i] * rloc[i];
    }
  // hdxb^T Rinv hdxb (:127-128)
  dgemm_tn(ne, ne, nobsl, hdxb_rinv.data(), nobsl, hdxb_c.data(), nobsl, work1.data(), ne);
  // + (m-1) I / rho (:140-143)
  double rho = 1.0 / *parm_infl;
  for (int i = 0; i < ne; ++i) work1[i + (size_t)i * ne] += (double)(ne - 1) * rho;
  // eigen-decomposition (:147)
  int nrank = oracle_mtx_eigen(ne, work1.data(), eival.data(), eivec.data());
  if (nrank < 0) return nrank;
  // Pa = V D^-1 V^T (:151-157)
  for (int j = 0; j < ne; ++j)
    for (int i = 0; i < ne; ++i) work1[i + (size_t)j * ne] = eivec[i + (size_t)j * ne] / eival[j];
  dgemm_nt(ne, ne, ne, work1.data(), ne, eivec.data(), ne, pa.data(), ne);
  // Pa hdxb_rinv^T (:169-170)
  dgemm_nt(ne, nobsl, ne, pa.data(), ne, hdxb_rinv.data(), nobsl, work2.data(), ne);
  // Pa hdxb_rinv^T dep (:182-195)
  for (int i = 0; i < ne; ++i) {
    double w = work2[i] * dep[0];
    for (int j = 1; j < nobsl; ++j) w += work2[i + (size_t)j * ne] * dep[j];
    work3[i] = w;
  }
  if (depd && transmd) {
    for (int i = 0; i < ne; ++i) {
      double w = work2[i] * depd[0];
      for (int j = 1; j < nobsl; ++j) w += work2[i + (size_t)j * ne] * depd[j];
      transmd[i] = w;
    }
  }
  // T = sqrt[(m-1) Pa] (:199-206)
  for (int j = 0; j < ne; ++j) {
    rho = std::sqrt((double)(ne - 1) / eival[j]);
    for (int i = 0; i < ne; ++i) work1[i + (size_t)j * ne] = eivec[i + (size_t)j * ne] * rho;
  }
  dgemm_nt(ne, ne, ne, work1.data(), ne, eivec.data(), ne, trans, ne);
  // T + Pa hdxb_rinv^T dep (:218-227)
  if (transm) {
    std::copy(work3.begin(), work3.end(), transm);
  } else {
    for (int j = 0; j < ne; ++j)
      for (int i = 0; i < ne; ++i) trans[i + (size_t)j * ne] += work3[i];
  }
  if (pao) std::copy(pa.begin(), pa.end(), pao);
  if (!infl_update) return 0;
  // adaptive inflation estimate (:229-254)
  double parm[4] = {0.0, 0.0, 0.0, 0.0};
  const double sigma_b = 0.04;
  for (int i = 0; i < nobsl; ++i)
    parm[0] += rdiag_wloc ? dep[i] * dep[i] / rdiag[i] : dep[i] * dep[i] / rdiag[i] * rloc[i];
  for (int j = 0; j < ne; ++j)
    for (int i = 0; i < nobsl; ++i)
      parm[1] += hdxb_rinv[i + (size_t)j * nobsl] * hdxb_c[i + (size_t)j * nobsl];
  parm[1] = parm[1] / (double)(ne - 1);
  for (int i = 0; i < nobsl; ++i) parm[2] += rloc[i];
  parm[3] = (parm[0] - parm[2]) / parm[1] - *parm_infl;
  const double t = (*parm_infl * parm[1] + parm[2]) / parm[1];
  const double sigma_o = 2.0 / parm[2] * (t * t);
  const double gain = sigma_b * sigma_b / (sigma_o + sigma_b * sigma_b);
  *parm_infl = *parm_infl + gain * parm[3];
  return 0;
}

int oracle_core_batch(int ne, int nobs, int npts, const int32_t *nobsl, const double *hdxb,
                      const double *rdiag, const double *rloc, const double *dep,
                      double *parm_infl, double *trans, double *transm, double *pao,
                      int rdiag_wloc, int infl_update, const double *depd, double *transmd,
                      int nthreads) {
  int status = 0;
  const size_t k2 = (size_t)ne * ne;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
  for (int i = 0; i < npts; ++i) {
    int r = oracle_letkf_core(ne, nobs, nobsl[i], hdxb + (size_t)i * nobs * ne,
                              rdiag + (size_t)i * nobs, rloc + (size_t)i * nobs,
                              dep + (size_t)i * nobs, parm_infl + i, trans + i * k2,
                              transm ? transm + (size_t)i * ne : nullptr,
                              pao ? pao + i * k2 : nullptr, rdiag_wloc, infl_update,
                              depd ? depd + (size_t)i * nobs : nullptr,
                              transmd ? transmd + (size_t)i * ne : nullptr);
    if (r != 0) {
#ifdef _OPENMP
#pragma omp critical
#endif
      status = r;
    }
  }
  return status;
}

}  // extern "C"

// ---- common/common_sort.f90 -----------------------------------------------------------
namespace {
struct Asc {
  bool operator()(double a, double b) const { return a < b; }
};
struct Desc {
  bool operator()(double a, double b) const { return a > b; }
};
// X and A are addressed 1-based exactly like the Fortran (X(j) is a 1-based index into A).
template <class Cmp>
struct Select {
  const double *A;
  int32_t *X;
  Cmp lt;
  double key(int pos) const { return A[X[pos - 1] - 1]; }
  // partition_arg / partition_desc_arg (:64-86 / :115-137): Lomuto, strict comparison
  int partition(int left, int right, int pivot) {
    const double a_pivot = key(pivot);
    std::swap(X[pivot - 1], X[right - 1]);
    int store = left;
    for (int idx = left; idx <= right - 1; ++idx) {
      if (lt(key(idx), a_pivot)) {
        std::swap(X[store - 1], X[idx - 1]);
        ++store;
      }
    }
    std::swap(X[right - 1], X[store - 1]);
    return store;
  }
  // median_of_three_arg (:168-192); the same decision tree is used for both orders
  int median3(int i1, int i2, int i3) const {
    if (key(i1) < key(i2)) {
      if (key(i2) < key(i3)) return i2;
      if (key(i1) < key(i3)) return i3;
      return i1;
    }
    if (key(i1) < key(i3)) return i1;
    if (key(i2) < key(i3)) return i3;
    return i2;
  }
  // sample_second_min_arg / sample_second_max_arg (:224-249 / :281-306)
  int sample_second(int left, int right, int K) const {
    int i = left, i2 = left;
    double best = lt(0.0, 1.0) ? std::numeric_limits<double>::max()
                               : -std::numeric_limits<double>::max();
    double best2 = best;
    for (int j = left; j <= right; j += K) {
      if (lt(key(j), best)) {
        best2 = best;
        best = key(j);
        i2 = i;
        i = j;
      } else if (lt(key(j), best2)) {
        best2 = key(j);
        i2 = j;
      }
    }
    return i2;
  }
  // QUICKSELECT_arg / QUICKSELECT_desc_arg (:341-369 / :404-432); tail recursion unrolled
  void run(int left, int right, int K) {
    while (left < right) {
      int pivot;
      if ((right - left) / K >= 2) {
        pivot = sample_second(left, right, K);
      } else {
        pivot = median3(left, (left + right) / 2, right);
      }
      pivot = partition(left, right, pivot);
      if (K < pivot) {
        right = pivot - 1;
      } else if (K > pivot) {
        left = pivot + 1;
      } else {
        return;
      }
    }
  }
};
}  // namespace

extern "C" {
void oracle_quickselect_arg(const double *A, int32_t *X, int left, int right, int K) {
  Select<Asc> s{A, X, Asc()};
  s.run(left, right, K);
}
void oracle_quickselect_desc_arg(const double *A, int32_t *X, int left, int right, int K) {
  Select<Desc> s{A, X, Desc()};
  s.run(left, right, K);
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}  // extern "C"
