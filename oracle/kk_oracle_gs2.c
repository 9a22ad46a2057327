/*
 * kk_oracle_gs2.c -- CPU restatement of the reference's TWO-STAGE Gauss-Seidel (GS_TWOSTAGE with inner Jacobi-Richardson
 * sweeps): the Gauss-Seidel variant that is a loop of SpMVs.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c).
 *
 * Follows sparse/impl/KokkosSparse_twostage_gauss_seidel_impl.hpp (paths relative to /root/reference):
 *   symbolic  :544-697 with the counting / filling functors :274-383 -- L = strictly lower entries, U = strictly upper entries
 *             with column < num_rows, both in the storage order of A; compact form also La ("complement of U+D": the lower
 *             entries and the columns >= num_rows) and Ua ("complement of L+D": the upper entries, all columns)
 *   numeric   :385-470 -- D = 1 / a_ii (or the inverse diagonal the caller supplied), L and U row-scaled by it
 *             (values(k) *= diags(i)), La / Ua unscaled, Da = a_ii
 *   apply     :778-1035 -- per sweep (symmetric = forward then backward, NumSweeps = max(outer sweeps, numIter), doubled):
 *               R = B;  classic: R -= A x;  compact: R -= (forward ? Ua : La) x, and if omega != 1: R += (1/omega - 1) Da.*x
 *               (both skipped on the first sweep when x was zeroed)
 *               no inner sweeps:  Z = D.*R (times gamma if gamma != 1)
 *               else              T = D.*R;  R = T (times gamma);  per inner sweep: Z = T;  Z += -omega (L or U) R;
 *                                 gamma != 1: Z = gamma Z + (1 - gamma) R;  R = Z unless last
 *               x = compact ? omega Z : x + omega Z            (rows 0 .. num_rows-1 of x; x has num_cols entries)
 *             every SpMV is KokkosSparse::spmv("N", alpha, M, x, one, y): restated with the host loop order O1
 *             (kk_oracle.c okk_spmv_serial_*: y = beta*y + alpha*sum), KokkosBlas::mult / scal / axpy as written (mult with
 *             beta = 0 overwrites).
 * The sptrsv variant (two_stage = false, "classic" in the unit test) is restated in kk_oracle_sptrsv.c.
 * Pinned by (a) the definition -- with enough inner sweeps the inner iteration converges to the triangular solve, so a
 * forward sweep equals textbook Gauss-Seidel to rounding (tests/test_oracle_gs2.py) -- and (b) the reference unit test's
 * acceptance (sparse/unit_test/Test_Sparse_gauss_seidel.hpp:236-241: error norm below the initial one).  Parity unpinned
 * against reference bits: the functor needs KokkosBlas and KokkosSparse::spmv around it.
 * Compiled with -ffp-contract=off.
 */
#include <stdint.h>
#include <stdlib.h>

#define OKK_API __attribute__((visibility("default")))

void okk_spmv_serial_f64(int nrow, const int* rm, const int* ci, const double* v, const double* x, double* y, double alpha,
                         double beta);
void okk_spmv_serial_f32(int nrow, const int* rm, const int* ci, const float* v, const float* x, float* y, float alpha,
                         float beta);

#define DEF_GS2(NAME, T, SPMV)                                                                                       \
  /* returns 0, or 1 + the first row without a diagonal entry */                                                     \
  OKK_API int NAME(int n, int ncols, const int* rm, const int* ci, const T* v, const T* given_inverse_diagonal,      \
                   int compact, int inner_sweeps, int outer_sweeps, T gamma, T* x, const T* b, int init_zero_x,      \
                   T omega, int num_iter, int direction) {                                                           \
    const T one = (T)1, zero = (T)0;                                                                                 \
    (void)ncols;                                                                                                     \
    int* rl = (int*)calloc((size_t)n + 1, sizeof(int));                                                              \
    int* ru = (int*)calloc((size_t)n + 1, sizeof(int));                                                              \
    int* rla = (int*)calloc((size_t)n + 1, sizeof(int));                                                             \
    int* rua = (int*)calloc((size_t)n + 1, sizeof(int));                                                             \
    for (int i = 0; i < n; ++i) {                                                                                    \
      int cl = 0, cu = 0, cla = 0, cua = 0, diag = 0;                                                                \
      for (int k = rm[i]; k < rm[i + 1]; ++k) {                                                                      \
        if (ci[k] < i) { ++cl; ++cla; }                                                                              \
        else if (ci[k] > i) {                                                                                        \
          if (ci[k] < n) { ++cu; ++cua; }                                                                            \
          else { ++cla; ++cua; }                                                                                     \
        } else diag = 1;                                                                                             \
      }                                                                                                              \
      if (!diag) { free(rl); free(ru); free(rla); free(rua); return 1 + i; }                                         \
      rl[i + 1] = rl[i] + cl; ru[i + 1] = ru[i] + cu; rla[i + 1] = rla[i] + cla; rua[i + 1] = rua[i] + cua;          \
    }                                                                                                                \
    int* el = (int*)malloc(sizeof(int) * (size_t)(rl[n] + 1));                                                       \
    int* eu = (int*)malloc(sizeof(int) * (size_t)(ru[n] + 1));                                                       \
    int* ela = (int*)malloc(sizeof(int) * (size_t)(rla[n] + 1));                                                     \
    int* eua = (int*)malloc(sizeof(int) * (size_t)(rua[n] + 1));                                                     \
    T* vl = (T*)malloc(sizeof(T) * (size_t)(rl[n] + 1));                                                             \
    T* vu = (T*)malloc(sizeof(T) * (size_t)(ru[n] + 1));                                                             \
    T* vla = (T*)malloc(sizeof(T) * (size_t)(rla[n] + 1));                                                           \
    T* vua = (T*)malloc(sizeof(T) * (size_t)(rua[n] + 1));                                                           \
    T* D = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                  \
    T* Da = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                 \
    for (int i = 0; i < n; ++i) {                                                                                    \
      int pl = rl[i], pu = ru[i], pla = rla[i], pua = rua[i];                                                        \
      for (int k = rm[i]; k < rm[i + 1]; ++k) {                                                                      \
        if (ci[k] < i) {                                                                                             \
          el[pl] = ci[k]; vl[pl++] = v[k];                                                                           \
          ela[pla] = ci[k]; vla[pla++] = v[k];                                                                       \
        } else if (ci[k] == i) {                                                                                     \
          D[i] = given_inverse_diagonal ? given_inverse_diagonal[i] : v[k];                                          \
          Da[i] = v[k];                                                                                              \
        } else if (ci[k] < n) {                                                                                      \
          eu[pu] = ci[k]; vu[pu++] = v[k];                                                                           \
          eua[pua] = ci[k]; vua[pua++] = v[k];                                                                       \
        } else {                                                                                                     \
          ela[pla] = ci[k]; vla[pla++] = v[k];                                                                       \
          eua[pua] = ci[k]; vua[pua++] = v[k];                                                                       \
        }                                                                                                            \
      }                                                                                                              \
      if (!given_inverse_diagonal) D[i] = one / D[i];                                                                \
      for (int k = rl[i]; k < rl[i + 1]; ++k) vl[k] *= D[i];                                                         \
      for (int k = ru[i]; k < ru[i + 1]; ++k) vu[k] *= D[i];                                                         \
    }                                                                                                                \
    T* R = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                  \
    T* Tt = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                 \
    T* Z = (T*)malloc(sizeof(T) * (size_t)(n + 1));                                                                  \
    int sweeps = outer_sweeps > num_iter ? outer_sweeps : num_iter;                                                  \
    if (direction == 0) sweeps *= 2;                                                                                 \
    if (init_zero_x)                                                                                                 \
      for (int i = 0; i < ncols; ++i) x[i] = zero;                                                                   \
    for (int sweep = 0; sweep < sweeps; ++sweep) {                                                                   \
      const int forward = direction == 1 || (direction == 0 && sweep % 2 == 0);                                      \
      for (int i = 0; i < n; ++i) R[i] = one * b[i];                                                                 \
      if (sweep > 0 || !init_zero_x) {                                                                               \
        if (compact) {                                                                                               \
          if (forward) SPMV(n, rua, eua, vua, x, R, -one, one);                                                      \
          else SPMV(n, rla, ela, vla, x, R, -one, one);                                                              \
          if (omega != one) {                                                                                        \
            const T omega2 = one / omega - one;                                                                      \
            for (int i = 0; i < n; ++i) Z[i] = one * Da[i] * x[i];                                                   \
            for (int i = 0; i < n; ++i) R[i] += omega2 * Z[i];                                                       \
          }                                                                                                          \
        } else {                                                                                                     \
          SPMV(n, rm, ci, v, x, R, -one, one);                                                                       \
        }                                                                                                            \
      }                                                                                                              \
      if (inner_sweeps == 0) {                                                                                       \
        for (int i = 0; i < n; ++i) Z[i] = one * D[i] * R[i];                                                        \
        if (gamma != one)                                                                                            \
          for (int i = 0; i < n; ++i) Z[i] = gamma * Z[i];                                                           \
      } else {                                                                                                       \
        for (int i = 0; i < n; ++i) Tt[i] = one * D[i] * R[i];                                                       \
        for (int i = 0; i < n; ++i) R[i] = one * Tt[i];                                                              \
        if (gamma != one)                                                                                            \
          for (int i = 0; i < n; ++i) R[i] = gamma * R[i];                                                           \
      }                                                                                                              \
      for (int ii = 0; ii < inner_sweeps; ++ii) {                                                                    \
        for (int i = 0; i < n; ++i) Z[i] = one * Tt[i];                                                              \
        if (forward) SPMV(n, rl, el, vl, R, Z, -omega, one);                                                         \
        else SPMV(n, ru, eu, vu, R, Z, -omega, one);                                                                 \
        if (gamma != one) {                                                                                          \
          const T gamma2 = one - gamma;                                                                              \
          for (int i = 0; i < n; ++i) Z[i] = gamma * Z[i];                                                           \
          for (int i = 0; i < n; ++i) Z[i] += gamma2 * R[i];                                                         \
        }                                                                                                            \
        if (ii + 1 < inner_sweeps)                                                                                   \
          for (int i = 0; i < n; ++i) R[i] = one * Z[i];                                                             \
      }                                                                                                              \
      if (compact)                                                                                                   \
        for (int i = 0; i < n; ++i) x[i] = omega * Z[i];                                                             \
      else                                                                                                           \
        for (int i = 0; i < n; ++i) x[i] += omega * Z[i];                                                            \
    }                                                                                                                \
    free(rl); free(ru); free(rla); free(rua); free(el); free(eu); free(ela); free(eua);                              \
    free(vl); free(vu); free(vla); free(vua); free(D); free(Da); free(R); free(Tt); free(Z);                         \
    return 0;                                                                                                        \
  }

DEF_GS2(okk_gs2_apply_f64, double, okk_spmv_serial_f64)
DEF_GS2(okk_gs2_apply_f32, float, okk_spmv_serial_f32)
