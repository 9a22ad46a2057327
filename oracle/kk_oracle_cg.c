/*
 * kk_oracle_cg.c -- CPU restatement of the reference's unpreconditioned CG driver (SURVEY.md section 8f rank 4:
 * "drivers that call spmv in a loop").
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c).
 *
 * Follows KokkosKernels::Experimental::Example::pcgsolve with use_sgs = false
 * (perf_test/sparse/KokkosSparse_pcg.hpp:248-466; driver perf_test/sparse/KokkosSparse_pcg.cpp:69-122, tolerance 1e-7):
 *   p = x; Ap = A p; r = b - Ap; p = r; old_rdot = r.r; norm_res = sqrt(old_rdot)          (:279-292)
 *   while (tolerance < norm_res && iteration < maximum_iteration)                           (:372)
 *     Ap = A p; pAp = p.Ap; alpha = old_rdot / pAp                                          (:376-391)
 *     x = alpha p + x; r = -alpha Ap + r; r_dot = r.r                                       (:394-398)
 *     beta = r_dot / old_rdot; p = r + beta p; norm_res = sqrt(old_rdot = r_dot); ++iter    (:400,434,448-452)
 * spmv("N", 1, A, p, 0, Ap) in the functor order (sparse/impl/KokkosSparse_spmv_impl.hpp:110-132), KokkosBlas::dot and
 * axpby as their Serial loops (one accumulator; a*x + b*y per element).  Not pinned bit for bit on reference code (the
 * driver needs real Kokkos); pinned by definition: on return b - A x has the norm it reports (tests/test_oracle_cg.py).
 * Compiled with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define OKK_API __attribute__((visibility("default")))

static void cg_spmv(int n, const int* rm, const int* ci, const double* v, const double* x, double* y) {
  for (int i = 0; i < n; ++i) {
    double sum = 0.0;
    for (int j = rm[i]; j < rm[i + 1]; ++j) sum += v[j] * x[ci[j]];
    y[i] = sum; /* alpha = 1, beta = 0 */
  }
}
static double cg_dot(int n, const double* a, const double* b) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

int okk_gs2_apply_f64(int n, int ncols, const int* rm, const int* ci, const double* v, const double* given_inverse_diagonal, int compact,
                      int inner_sweeps, int outer_sweeps, double gamma, double* x, const double* b, int init_zero_x, double omega, int num_iter,
                      int direction);
void okk_gs_apply_f64(int n, const int* rm, const int* ci, const double* v, int ncolors, const int* color_ptr, const int* color_rows,
                      const double* dinv, const double* y, double* x, int init_zero_x, double omega, int sweeps, int direction);

/* use_sgs = true (pcg.hpp:339-358,412-437): z = one symmetric point Gauss-Seidel sweep on r from z = 0, omega = 1, over the
 * given colour sets (kk_oracle_gs.c); alpha = r.z / p.Ap; beta = r.z' / r.z; p = z + beta p.  ncolors = 0: no preconditioner. */
static int cg_core(int n, const int* row_map, const int* col_idx, const double* values, const double* b, double* x, int maximum_iteration,
                   double tolerance, double* norm_res_out, int ncolors, const int* color_ptr, const int* color_rows, const double* dinv);

/* returns the iteration count; *norm_res_out = sqrt(r.r) of the recurrence residual */
OKK_API int okk_cg_f64(int n, const int* row_map, const int* col_idx, const double* values, const double* b, double* x,
                       int maximum_iteration, double tolerance, double* norm_res_out) {
  return cg_core(n, row_map, col_idx, values, b, x, maximum_iteration, tolerance, norm_res_out, 0, 0, 0, 0);
}
OKK_API int okk_pcg_gs2_f64(int n, const int* row_map, const int* col_idx, const double* values, const double* b, double* x,
                            int maximum_iteration, double tolerance, double* norm_res_out, int inner_sweeps, int compact) {
  const int opt[2] = {inner_sweeps, compact};
  return cg_core(n, row_map, col_idx, values, b, x, maximum_iteration, tolerance, norm_res_out, -1, opt, 0, 0);
}
OKK_API int okk_pcg_f64(int n, const int* row_map, const int* col_idx, const double* values, const double* b, double* x,
                        int maximum_iteration, double tolerance, double* norm_res_out, int ncolors, const int* color_ptr,
                        const int* color_rows, const double* dinv) {
  return cg_core(n, row_map, col_idx, values, b, x, maximum_iteration, tolerance, norm_res_out, ncolors, color_ptr, color_rows, dinv);
}
/* ncolors > 0: point (multicolour) symmetric Gauss-Seidel over the given colour sets; ncolors == -1: the two-stage symmetric
 * Gauss-Seidel (kk_oracle_gs2.c), color_ptr[0] = inner sweeps, color_ptr[1] = compact form -- what the reference's pcgsolve runs when
 * the caller's kernel handle holds a GS_TWOSTAGE handle (symmetric_gauss_seidel_apply dispatches on it) */
static void cg_precond(int n, const int* row_map, const int* col_idx, const double* values, int ncolors, const int* color_ptr,
                       const int* color_rows, const double* dinv, const double* r, double* z) {
  if (ncolors > 0) okk_gs_apply_f64(n, row_map, col_idx, values, ncolors, color_ptr, color_rows, dinv, r, z, 1, 1.0, 1, 0);
  else okk_gs2_apply_f64(n, n, row_map, col_idx, values, 0, color_ptr[1], color_ptr[0], 1, 1.0, z, r, 1, 1.0, 1, 0);
}
static int cg_core(int n, const int* row_map, const int* col_idx, const double* values, const double* b, double* x, int maximum_iteration,
                   double tolerance, double* norm_res_out, int ncolors, const int* color_ptr, const int* color_rows, const double* dinv) {
  const int use_sgs = ncolors != 0;
  double* z = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double precond_old_rdot = 1;
  double* p = (double*)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  double* r = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  double* Ap = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) p[i] = x[i];
  cg_spmv(n, row_map, col_idx, values, p, Ap);
  for (int i = 0; i < n; ++i) r[i] = 1.0 * b[i] + -1.0 * Ap[i];
  for (int i = 0; i < n; ++i) p[i] = r[i];
  double old_rdot = cg_dot(n, r, r);
  double norm_res = sqrt(old_rdot);
  if (use_sgs) {
    cg_precond(n, row_map, col_idx, values, ncolors, color_ptr, color_rows, dinv, r, z);
    precond_old_rdot = cg_dot(n, r, z);
    for (int i = 0; i < n; ++i) p[i] = z[i];
  }
  int iteration = 0;
  while (tolerance < norm_res && iteration < maximum_iteration) {
    cg_spmv(n, row_map, col_idx, values, p, Ap);
    const double pAp_dot = cg_dot(n, p, Ap);
    const double alpha = (use_sgs ? precond_old_rdot : old_rdot) / pAp_dot;
    for (int i = 0; i < n; ++i) x[i] = alpha * p[i] + 1.0 * x[i];
    for (int i = 0; i < n; ++i) r[i] = -alpha * Ap[i] + 1.0 * r[i];
    const double r_dot = cg_dot(n, r, r);
    double beta = r_dot / old_rdot;
    if (use_sgs) {
      cg_precond(n, row_map, col_idx, values, ncolors, color_ptr, color_rows, dinv, r, z);
      const double precond_r_dot = cg_dot(n, r, z);
      beta = precond_r_dot / precond_old_rdot;
      for (int i = 0; i < n; ++i) p[i] = 1.0 * z[i] + beta * p[i];
      precond_old_rdot = precond_r_dot;
    } else {
      for (int i = 0; i < n; ++i) p[i] = 1.0 * r[i] + beta * p[i];
    }
    norm_res = sqrt(old_rdot = r_dot);
    ++iteration;
  }
  free(p);
  free(r);
  free(Ap);
  free(z);
  *norm_res_out = norm_res;
  return iteration;
}
