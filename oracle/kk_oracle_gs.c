/*
 * kk_oracle_gs.c -- CPU restatement of the reference's point (multicolour) Gauss-Seidel apply (SURVEY.md section 8f rank 4:
 * the preconditioner of its CG driver).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c).
 *
 * Follows PointGaussSeidel::PSGS::operator() (sparse/impl/KokkosSparse_gauss_seidel_impl.hpp:159-179) and the sweep driver
 * (DoPSGS, same file): one sweep visits the colour sets in order (forward: first to last; backward: last to first; symmetric:
 * forward then backward), and for every row ii of the set
 *     sum = y(ii);  for each entry: sum -= a(ii, col) * x(col);  x(ii) += omega * sum * inverse_diagonal(ii)
 * (entries in storage order, diagonal entry included).  The reference permutes the matrix so that a colour set is a
 * contiguous row range; here the sets are given as a row list (color_rows grouped by color_ptr) on the unpermuted matrix --
 * the arithmetic per row is the same and rows of one set are independent of each other.
 * Which rows share a colour is the colouring algorithm's business (KokkosGraph, not restated): tests take the colouring the
 * library produced, check that it is a proper distance-1 colouring, and hand it to this function.  Pinned by definition
 * (a symmetric sweep with one colour per row is textbook SSOR, tests/test_oracle_gs.py) and by the reference unit test's
 * acceptance (error norm below the initial one after two sweeps, sparse/unit_test/Test_Sparse_gauss_seidel.hpp:198-216).
 * Compiled with -ffp-contract=off.
 */
#include <stdint.h>

#define OKK_API __attribute__((visibility("default")))

#define DEF_GS(NAME, T)                                                                                            \
  static void NAME##_set(const int* rm, const int* ci, const T* v, const int* rows, int b, int e, const T* dinv,   \
                         const T* y, T* x, T omega) {                                                              \
    for (int q = b; q < e; ++q) {                                                                                  \
      const int ii = rows[q];                                                                                      \
      T sum = y[ii];                                                                                               \
      for (int j = rm[ii]; j < rm[ii + 1]; ++j) sum -= v[j] * x[ci[j]];                                            \
      x[ii] += omega * sum * dinv[ii];                                                                             \
    }                                                                                                              \
  }                                                                                                                \
  /* direction: 0 symmetric, 1 forward, 2 backward (apply_type of the reference's test, :104) */                   \
  OKK_API void NAME(int n, const int* rm, const int* ci, const T* v, int ncolors, const int* color_ptr,            \
                    const int* color_rows, const T* dinv, const T* y, T* x, int init_zero_x, T omega, int sweeps,  \
                    int direction) {                                                                               \
    if (init_zero_x)                                                                                               \
      for (int i = 0; i < n; ++i) x[i] = (T)0;                                                                     \
    for (int s = 0; s < sweeps; ++s) {                                                                             \
      if (direction == 0 || direction == 1)                                                                        \
        for (int c = 0; c < ncolors; ++c) NAME##_set(rm, ci, v, color_rows, color_ptr[c], color_ptr[c + 1], dinv, y, x, omega); \
      if (direction == 0 || direction == 2)                                                                        \
        for (int c = ncolors - 1; c >= 0; --c)                                                                     \
          NAME##_set(rm, ci, v, color_rows, color_ptr[c], color_ptr[c + 1], dinv, y, x, omega);                    \
    }                                                                                                              \
  }

DEF_GS(okk_gs_apply_f64, double)
DEF_GS(okk_gs_apply_f32, float)
