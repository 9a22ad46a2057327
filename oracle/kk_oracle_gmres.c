/*
 * kk_oracle_gmres.c -- CPU restatement of KokkosSparse::Experimental::gmres.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c).
 *
 * Follows GmresWrap::gmres, sparse/impl/KokkosSparse_gmres_impl.hpp:58-327, for real scalars on a CrsMatrix, with the
 * optional right preconditioner in the one form the reference ships and tests: MatrixPrec, whose apply is an spmv with a
 * given matrix (sparse/src/KokkosSparse_MatrixPrec.hpp:79-83).  Restart length m, tolerance tol, max_restart, ortho
 * (0 = CGS2, 1 = MGS) and the three results (num_iters, end_rel_res, conv_flag: 0 Conv, 1 NoConv, 2 LOA) are the
 * GMRESHandle's (sparse/src/KokkosSparse_gmres_handle.hpp:76-110,175).
 *   initial residual and the zero-rhs special cases                      :112-133
 *   Arnoldi step: (prec,) spmv, MGS (:150-156) or CGS2 (:157-171), norm, new basis vector (:176-181)
 *   Givens rotations after Demmel et al. (:186-204), shortcut residual (:205), breakdown / NaN throws (:211-218) -> -1 / -2
 *   least squares by back substitution on the rotated H (:221-235), solution update (:237-249), true residual (:250-261)
 * BLAS pieces in their serial orders: dot / nrm2 = one accumulator (nrm2 = sqrt of it), gemv "C" = one dot per column,
 * gemv "N" = axpy per column.  V is column-major n x (m+1).  Not pinned bit for bit on reference code (needs real Kokkos);
 * pinned by the unit test's acceptance (sparse/unit_test/Test_Sparse_gmres.hpp:109-170): true relative residual below the
 * tolerance and flag Conv, on its own problem family (tests/test_oracle_gmres.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OKK_API __attribute__((visibility("default")))

#define DEF_GMRES(NAME, T, SQRT, FABS)                                                                          \
  static void NAME##_spmv(int n, const int* rm, const int* ci, const T* v, const T* x, T* y) {                  \
    for (int i = 0; i < n; ++i) {                                                                               \
      T sum = (T)0;                                                                                             \
      for (int j = rm[i]; j < rm[i + 1]; ++j) sum += v[j] * x[ci[j]];                                           \
      y[i] = sum;                                                                                               \
    }                                                                                                           \
  }                                                                                                             \
  static T NAME##_dot(int n, const T* a, const T* b) {                                                          \
    T s = (T)0;                                                                                                 \
    for (int i = 0; i < n; ++i) s += a[i] * b[i];                                                               \
    return s;                                                                                                   \
  }                                                                                                             \
  OKK_API int NAME(int n, const int* rm, const int* ci, const T* v, const int* prm, const int* pci,             \
                   const T* pv, const T* B, T* X, int m, T tol, int max_restart, int ortho, int* num_iters_out, \
                   T* end_rel_res_out, int* conv_flag_out) {                                                    \
    const size_t N = (size_t)(n > 0 ? n : 1);                                                                   \
    T* Xiter = (T*)calloc(N, sizeof(T));                                                                        \
    T* Res = (T*)calloc(N, sizeof(T));                                                                          \
    T* Wj = (T*)calloc(N, sizeof(T));                                                                           \
    T* Wj2 = (T*)calloc(N, sizeof(T));                                                                          \
    T* V = (T*)calloc(N * (size_t)(m + 1), sizeof(T));                                                          \
    T* H = (T*)calloc((size_t)(m + 1) * (size_t)m, sizeof(T)); /* H(i,j) at H[i + j*(m+1)] */                   \
    T* GVec = (T*)calloc((size_t)m + 1, sizeof(T));                                                             \
    T* Ls = (T*)calloc((size_t)m, sizeof(T));                                                                   \
    T* Cos = (T*)calloc((size_t)m, sizeof(T));                                                                  \
    T* Sin = (T*)calloc((size_t)m, sizeof(T));                                                                  \
    T* tmp = (T*)calloc((size_t)m, sizeof(T));                                                                  \
    int status = 0, converged = 0, cycle = 0, numIters = 0;                                                     \
    T nrmB, trueRes, relRes, shortRelRes;                                                                       \
    nrmB = SQRT(NAME##_dot(n, B, B));                                                                           \
    memcpy(Res, B, sizeof(T) * (size_t)n);                                                                      \
    NAME##_spmv(n, rm, ci, v, X, Wj);                                                                           \
    for (int i = 0; i < n; ++i) Res[i] += (T)-1 * Wj[i];                                                        \
    trueRes = SQRT(NAME##_dot(n, Res, Res));                                                                    \
    if (nrmB != (T)0) relRes = trueRes / nrmB;                                                                  \
    else if (trueRes == (T)0) relRes = trueRes;                                                                 \
    else {                                                                                                      \
      for (int i = 0; i < n; ++i) X[i] = (T)0;                                                                  \
      relRes = (T)0;                                                                                            \
    }                                                                                                           \
    shortRelRes = relRes;                                                                                       \
    if (relRes < tol) converged = 1;                                                                            \
    memcpy(Xiter, X, sizeof(T) * (size_t)n); /* the reference leaves Xiter = 0 until the first update; X is copied back   \
                                                from it only after an update or at a restart, which always follows one */ \
    while (!converged && cycle <= max_restart && shortRelRes >= (T)1e-14 && status == 0) {                      \
      GVec[0] = trueRes;                                                                                        \
      T* Vj = V;                                                                                                \
      for (int i = 0; i < n; ++i) Vj[i] = (1 / trueRes) * Res[i];                                               \
      for (int j = 0; j < m; j++) {                                                                             \
        if (prm) {                                                                                              \
          NAME##_spmv(n, prm, pci, pv, Vj, Wj2);                                                                \
          NAME##_spmv(n, rm, ci, v, Wj2, Wj);                                                                   \
        } else {                                                                                                \
          NAME##_spmv(n, rm, ci, v, Vj, Wj);                                                                    \
        }                                                                                                       \
        T* Hj = H + (size_t)j * (m + 1);                                                                        \
        if (ortho == 1) {                                                                                       \
          for (int i = 0; i <= j; i++) {                                                                        \
            const T* Vi = V + (size_t)i * N;                                                                    \
            Hj[i] = NAME##_dot(n, Vi, Wj);                                                                      \
            for (int q = 0; q < n; ++q) Wj[q] += -Hj[i] * Vi[q];                                                \
          }                                                                                                     \
        } else {                                                                                                \
          for (int i = 0; i <= j; i++) Hj[i] = NAME##_dot(n, V + (size_t)i * N, Wj);                            \
          for (int i = 0; i <= j; i++) {                                                                        \
            const T* Vi = V + (size_t)i * N;                                                                    \
            for (int q = 0; q < n; ++q) Wj[q] += ((T)-1 * Hj[i]) * Vi[q];                                       \
          }                                                                                                     \
          for (int i = 0; i <= j; i++) tmp[i] = NAME##_dot(n, V + (size_t)i * N, Wj);                           \
          for (int i = 0; i <= j; i++) {                                                                        \
            const T* Vi = V + (size_t)i * N;                                                                    \
            for (int q = 0; q < n; ++q) Wj[q] += ((T)-1 * tmp[i]) * Vi[q];                                      \
          }                                                                                                     \
          for (int i = 0; i <= j; i++) Hj[i] += tmp[i];                                                         \
        }                                                                                                       \
        const T tmpNrm = SQRT(NAME##_dot(n, Wj, Wj));                                                           \
        Hj[j + 1] = tmpNrm;                                                                                     \
        if (tmpNrm > (T)1e-14) {                                                                                \
          Vj = V + (size_t)(j + 1) * N;                                                                         \
          for (int q = 0; q < n; ++q) Vj[q] = (1 / Hj[j + 1]) * Wj[q];                                          \
        }                                                                                                       \
        for (int i = 0; i < j; i++) {                                                                           \
          const T tempVal = Cos[i] * Hj[i] + Sin[i] * Hj[i + 1];                                                \
          Hj[i + 1] = -Sin[i] * Hj[i] + Cos[i] * Hj[i + 1];                                                     \
          Hj[i] = tempVal;                                                                                      \
        }                                                                                                       \
        const T f = Hj[j], g = Hj[j + 1];                                                                       \
        const T f2 = f * f, g2 = g * g;                                                                         \
        T fg2 = f2 + g2;                                                                                        \
        const T D1 = 1 / SQRT(f2 * fg2);                                                                        \
        Cos[j] = f2 * D1;                                                                                       \
        fg2 = fg2 * D1;                                                                                         \
        Hj[j] = f * fg2;                                                                                        \
        Sin[j] = f * D1 * g;                                                                                    \
        Hj[j + 1] = (T)0;                                                                                       \
        GVec[j + 1] = GVec[j] * (-Sin[j]);                                                                      \
        GVec[j] = GVec[j] * Cos[j];                                                                             \
        shortRelRes = FABS(GVec[j + 1]) / nrmB;                                                                 \
        if (tmpNrm <= (T)1e-14 && shortRelRes >= tol) { status = -1; break; } /* lucky breakdown throw */       \
        if (shortRelRes != shortRelRes) { status = -2; break; }              /* NaN throw */                    \
        if (shortRelRes < tol || j == m - 1) {                                                                  \
          for (int i = 0; i < m; ++i) Ls[i] = GVec[i];                                                          \
          for (int i = j; i >= 0; --i) { /* upper-triangular solve, SerialTrsm L U N N */                       \
            T s = Ls[i];                                                                                        \
            for (int q = i + 1; q <= j; ++q) s -= H[i + (size_t)q * (m + 1)] * Ls[q];                           \
            Ls[i] = s / H[i + (size_t)i * (m + 1)];                                                             \
          }                                                                                                     \
          memcpy(Xiter, X, sizeof(T) * (size_t)n);                                                              \
          if (prm) {                                                                                            \
            for (int q = 0; q < n; ++q) Wj[q] = (T)0;                                                           \
            for (int i = 0; i <= j; ++i) {                                                                      \
              const T* Vi = V + (size_t)i * N;                                                                  \
              for (int q = 0; q < n; ++q) Wj[q] += Ls[i] * Vi[q];                                               \
            }                                                                                                   \
            NAME##_spmv(n, prm, pci, pv, Wj, Wj2);                                                              \
            for (int q = 0; q < n; ++q) Xiter[q] = Xiter[q] + Wj2[q];                                           \
          } else {                                                                                              \
            for (int i = 0; i <= j; ++i) {                                                                      \
              const T* Vi = V + (size_t)i * N;                                                                  \
              for (int q = 0; q < n; ++q) Xiter[q] += Ls[i] * Vi[q];                                            \
            }                                                                                                   \
          }                                                                                                     \
          NAME##_spmv(n, rm, ci, v, Xiter, Wj);                                                                 \
          memcpy(Res, B, sizeof(T) * (size_t)n);                                                                \
          for (int q = 0; q < n; ++q) Res[q] += (T)-1 * Wj[q];                                                  \
          trueRes = SQRT(NAME##_dot(n, Res, Res));                                                              \
          relRes = trueRes / nrmB;                                                                              \
          numIters = j + 1;                                                                                     \
          if (relRes < tol) {                                                                                   \
            converged = 1;                                                                                      \
            memcpy(X, Xiter, sizeof(T) * (size_t)n);                                                            \
            break;                                                                                              \
          } else if (shortRelRes < (T)1e-30) {                                                                  \
            break;                                                                                              \
          }                                                                                                     \
        }                                                                                                       \
      }                                                                                                         \
      if (status != 0) break;                                                                                   \
      cycle++;                                                                                                  \
      memcpy(X, Xiter, sizeof(T) * (size_t)n);                                                                  \
    }                                                                                                           \
    *end_rel_res_out = relRes;                                                                                  \
    *conv_flag_out = converged ? 0 : (shortRelRes < tol ? 2 : 1);                                               \
    *num_iters_out = cycle > 0 ? (cycle - 1) * m + numIters : 0;                                                \
    free(Xiter); free(Res); free(Wj); free(Wj2); free(V); free(H); free(GVec); free(Ls); free(Cos); free(Sin);  \
    free(tmp);                                                                                                  \
    return status;                                                                                              \
  }

DEF_GMRES(okk_gmres_f64, double, sqrt, fabs)
DEF_GMRES(okk_gmres_f32, float, sqrtf, fabsf)
