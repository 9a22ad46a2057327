/*
 * kk_oracle_sptrsv.c -- CPU restatement of the sparse triangular solve and of the classic (sptrsv) form of the reference's
 * two-stage Gauss-Seidel.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c).
 *
 * okk_sptrsv_*            x = T^{-1} b by the serial substitution loop, the definition every algorithm of
 *                         KokkosSparse::sptrsv_solve implements (sparse/src/KokkosSparse_sptrsv.hpp:290-480; the host loops
 *                         sparse/impl/KokkosSparse_trsv_impl.hpp:60-130 are the same recurrence): rows in dependency order,
 *                         acc = b_i; acc -= a_ij x_j over the off-diagonal entries in STORAGE order; x_i = acc / a_ii.
 *                         side = 0: T is triangular as given (an entry on the wrong side: return 1 + row);
 *                         side = 1 / 2: T is the lower / upper triangle of a general matrix (the other entries and the columns
 *                         >= n are skipped), dinv != NULL: the diagonal is 1 / dinv_i.
 * okk_gs2_classic_apply_* sparse/impl/KokkosSparse_twostage_gauss_seidel_impl.hpp apply (:778-925) with two_stage == false:
 *                         per sweep R = B - A x (compact: R = B - Ua x or La x; skipped on the first sweep when x was zeroed),
 *                         Z = (L + D)^{-1} R (forward) or (U + D)^{-1} R (backward) with D = a_ii, or 1 / the inverse diagonal
 *                         the caller supplied (numeric :446-456), x += Z (compact: x = Z); omega must be 1 (:886-893).
 *                         The residual product is the host loop order O1 (okk_spmv_serial_*), like kk_oracle_gs2.c.
 * Pinned on the reference's own fixtures and check for the level-scheduled algorithms (sparse/unit_test/Test_Sparse_sptrsv.hpp:
 * 64-118 the "ones" matrices, :140-157 rhs = A * ones, :212-225 sum(lhs) == nrows: exact), on the definition (T x == b to rounding),
 * on scipy.sparse.linalg.spsolve_triangular, and one classic forward sweep == textbook Gauss-Seidel in natural order
 * (tests/test_oracle_sptrsv.py).  Not pinned on reference OUTPUT BITS for general inputs: the reference's solver functors
 * (sparse/impl/KokkosSparse_sptrsv_solve_impl.hpp) pull KokkosBatched / KokkosBlas headers and cannot be compiled in place the way
 * oracle/kkref_spmv.cpp compiles the SpMV loops; their row recurrence (rhs_i - sum a_ij lhs_j in storage order, / diagonal) is the one
 * restated here.  Compiled with -ffp-contract=off.
 */
#include <stdint.h>
#include <stdlib.h>

#define OKK_API __attribute__((visibility("default")))

void okk_spmv_serial_f64(int nrow, const int* rm, const int* ci, const double* v, const double* x, double* y, double alpha,
                         double beta);
void okk_spmv_serial_f32(int nrow, const int* rm, const int* ci, const float* v, const float* x, float* y, float alpha,
                         float beta);

#define DEF_SPTRSV(NAME, T)                                                                                           \
  OKK_API int NAME(int n, const int* rm, const int* ci, const T* v, const T* b, T* x, int lower, int side, const T* dinv) { \
    for (int q = 0; q < n; ++q) {                                                                                     \
      const int i = lower ? q : n - 1 - q;                                                                            \
      T acc = b[i];                                                                                                   \
      T d = (T)1;                                                                                                     \
      for (int k = rm[i]; k < rm[i + 1]; ++k) {                                                                       \
        const int c = ci[k];                                                                                          \
        if (c == i) { d = v[k]; continue; }                                                                           \
        if (side == 0) {                                                                                              \
          if ((lower ? (c > i) : (c < i)) || c >= n) return 1 + i;                                                    \
        } else if (side == 1) {                                                                                       \
          if (c > i) continue;                                                                                        \
        } else {                                                                                                      \
          if (c < i || c >= n) continue;                                                                              \
        }                                                                                                             \
        const T prod = v[k] * x[c];                                                                                   \
        acc = acc - prod;                                                                                             \
      }                                                                                                               \
      x[i] = dinv ? acc / ((T)1 / dinv[i]) : acc / d;                                                                 \
    }                                                                                                                 \
    return 0;                                                                                                         \
  }

DEF_SPTRSV(okk_sptrsv_f64, double)
DEF_SPTRSV(okk_sptrsv_f32, float)

#define DEF_GS2C(NAME, T, SPMV, TRSV)                                                                                 \
  OKK_API int NAME(int n, int ncols, const int* rm, const int* ci, const T* v, const T* given_inverse_diagonal,       \
                   int compact, int outer_sweeps, T* x, const T* b, int init_zero_x, int num_iter, int direction) {   \
    const T one = (T)1, zero = (T)0;                                                                                  \
    /* compact form: Ua = upper entries incl. ghost columns, La = lower entries + ghost columns */                    \
    int nnz = rm[n];                                                                                                  \
    int* ra = (int*)calloc((size_t)n + 1, sizeof(int));                                                               \
    int* ea = (int*)malloc(sizeof(int) * (size_t)(nnz + 1));                                                          \
    T* va = (T*)malloc(sizeof(T) * (size_t)(nnz + 1));                                                                \
    T* R = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                   \
    T* Z = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                   \
    int sweeps = outer_sweeps > num_iter ? outer_sweeps : num_iter;                                                   \
    if (direction == 0) sweeps *= 2;                                                                                  \
    if (init_zero_x)                                                                                                  \
      for (int i = 0; i < ncols; ++i) x[i] = zero;                                                                    \
    for (int sweep = 0; sweep < sweeps; ++sweep) {                                                                    \
      const int forward = direction == 1 || (direction == 0 && sweep % 2 == 0);                                       \
      for (int i = 0; i < n; ++i) R[i] = one * b[i];                                                                  \
      if (sweep > 0 || !init_zero_x) {                                                                                \
        if (compact) {                                                                                                \
          int p = 0;                                                                                                  \
          for (int i = 0; i < n; ++i) {                                                                               \
            ra[i] = p;                                                                                                \
            for (int k = rm[i]; k < rm[i + 1]; ++k) {                                                                 \
              const int c = ci[k];                                                                                    \
              const int take = forward ? (c > i) : (c < i || c >= n);                                                 \
              if (take) { ea[p] = c; va[p++] = v[k]; }                                                                \
            }                                                                                                         \
          }                                                                                                           \
          ra[n] = p;                                                                                                  \
          SPMV(n, ra, ea, va, x, R, -one, one);                                                                       \
        } else {                                                                                                      \
          SPMV(n, rm, ci, v, x, R, -one, one);                                                                        \
        }                                                                                                             \
      }                                                                                                               \
      const int rc = TRSV(n, rm, ci, v, R, Z, forward, forward ? 1 : 2, given_inverse_diagonal);                      \
      if (rc) { free(ra); free(ea); free(va); free(R); free(Z); return rc; }                                          \
      if (compact)                                                                                                    \
        for (int i = 0; i < n; ++i) x[i] = one * Z[i];                                                                \
      else                                                                                                            \
        for (int i = 0; i < n; ++i) x[i] += one * Z[i];                                                               \
    }                                                                                                                 \
    free(ra); free(ea); free(va); free(R); free(Z);                                                                   \
    return 0;                                                                                                         \
  }

DEF_GS2C(okk_gs2_classic_apply_f64, double, okk_spmv_serial_f64, okk_sptrsv_f64)
DEF_GS2C(okk_gs2_classic_apply_f32, float, okk_spmv_serial_f32, okk_sptrsv_f32)
