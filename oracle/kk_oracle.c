/*
 * kk_oracle.c -- CPU restatement of the Kokkos Kernels sparse hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke test
 * in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (libb200sparse.so) never links,
 * loads or calls anything in oracle/.
 *
 * Every function restates, operation for operation, a loop of the reference
 * (paths relative to /root/reference).  The reference as a whole cannot be
 * built here (needs Kokkos >= 4.6.02, CMakeLists.txt:150-157; only Kokkos 3.3
 * is on disk); parity is pinned (a) by the reference's own known-answer tests
 * (issue 101, NaN/beta==0, merge-matrix diagonal tables, issue-402 fixture) in
 * tests/test_oracle_*.py and (b) bit for bit by the reference's own code for
 * this path -- spmv_impl.hpp (Serial loop, generic functor, transpose,
 * multivector), spgemm_impl_seq.hpp, spgemm_jacobi_seq_impl.hpp -- compiled
 * from the reference tree where it lies over small stand-ins for the Kokkos
 * names it mentions (oracle/_ref, oracle/kkref_*.cpp, oracle/Makefile).
 *
 * Build flags: -O2 -ffp-contract=off (no FMA contraction, no fast-math) so
 * the rounding sequence is exactly the one the C expressions spell out.  A
 * second build with contraction on (libkkoracle_fma.so) brackets a reference
 * build whose compiler fused multiply-adds.
 *
 * Index types: Ordinal = Offset = int32 (default_types.hpp:41-58).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OKK_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * O1: Serial rank-1 SpMV, modes N / C (C == N for real scalars).
 * sparse/impl/KokkosSparse_spmv_impl.hpp:233-305.
 * dobeta is picked by the spec layer from beta: 0 / 1 / -1 / 2
 * (sparse/impl/KokkosSparse_spmv_spec.hpp:146-157); the Serial branch treats
 * -1 like 2 (its else-branch, :296-298).  Loop indices are int (:260-263).
 * YT = y's value type (accumulator type), AT = matrix value type, XT = x's.
 * ---------------------------------------------------------------------- */
#define DEF_SPMV_SERIAL(NAME, AT, XT, YT)                                          \
  OKK_API void NAME(int nrow, const int* row_map, const int* col_idx,              \
                    const AT* values, const XT* x, YT* y, YT alpha, YT beta) {     \
    const YT zero = (YT)0;                                                         \
    int dobeta = (beta == zero) ? 0 : (beta == (YT)1 ? 1 : 2);                     \
    if (nrow <= 0) return;                                                         \
    if (alpha == zero) { /* :249-258 */                                            \
      if (dobeta == 0) {                                                           \
        for (int i = 0; i < nrow; ++i) y[i] = zero;                                \
      } else if (dobeta == 1) {                                                    \
      } else {                                                                     \
        for (int i = 0; i < nrow; ++i) y[i] *= beta;                               \
      }                                                                            \
      return;                                                                      \
    }                                                                              \
    for (int i = 0; i < nrow; ++i) { /* :260-300 */                                \
      const int jbeg = row_map[i];                                                 \
      const int jend = row_map[i + 1];                                             \
      int j = jbeg;                                                                \
      const int jdist = (jend - jbeg) / 4;                                         \
      YT tmp1 = 0, tmp2 = 0, tmp3 = 0, tmp4 = 0;                                   \
      for (int jj = 0; jj < jdist; ++jj) {                                         \
        const AT value1 = values[j];                                               \
        const AT value2 = values[j + 1];                                           \
        const AT value3 = values[j + 2];                                           \
        const AT value4 = values[j + 3];                                           \
        const XT x_val1 = x[col_idx[j]];                                           \
        const XT x_val2 = x[col_idx[j + 1]];                                       \
        const XT x_val3 = x[col_idx[j + 2]];                                       \
        const XT x_val4 = x[col_idx[j + 3]];                                       \
        tmp1 += value1 * x_val1;                                                   \
        tmp2 += value2 * x_val2;                                                   \
        tmp3 += value3 * x_val3;                                                   \
        tmp4 += value4 * x_val4;                                                   \
        j += 4;                                                                    \
      }                                                                            \
      for (; j < jend; ++j) tmp1 += values[j] * x[col_idx[j]];                     \
      if (dobeta == 0) {                                                           \
        y[i] = alpha * (tmp1 + tmp2 + tmp3 + tmp4);                                \
      } else if (dobeta == 1) {                                                    \
        y[i] += alpha * (tmp1 + tmp2 + tmp3 + tmp4);                               \
      } else {                                                                     \
        const YT y_val = y[i] * beta;                                              \
        y[i] = y_val + alpha * (tmp1 + tmp2 + tmp3 + tmp4);                        \
      }                                                                            \
    }                                                                              \
  }

DEF_SPMV_SERIAL(okk_spmv_serial_f64, double, double, double)
DEF_SPMV_SERIAL(okk_spmv_serial_f32, float, float, float)
/* mixed precision of test_github_issue_101: float matrix, double vectors */
DEF_SPMV_SERIAL(okk_spmv_serial_f32mat_f64vec, float, double, double)

/* ------------------------------------------------------------------------
 * O2: generic functor order (the OpenMP RangePolicy path).
 * sparse/impl/KokkosSparse_spmv_impl.hpp:110-132: one accumulator in storage
 * order, sum *= alpha, dobeta==0 ? y=sum : y = beta*y+sum.
 * The front end short-cuts alpha==0 / empty A first
 * (sparse/src/KokkosSparse_spmv.hpp:145-154): beta==0 ? fill 0 : y *= beta.
 * `threads` > 1 runs rows under OpenMP (static schedule; dynamic when
 * nnz > 10M as :323-333 does) -- rows are independent so the bits do not
 * depend on the thread count.
 * ---------------------------------------------------------------------- */
#define DEF_SPMV_FUNCTOR(NAME, AT, XT, YT)                                         \
  OKK_API void NAME(int nrow, int ncol, const int* row_map, const int* col_idx,    \
                    const AT* values, const XT* x, YT* y, YT alpha, YT beta,       \
                    int threads) {                                                 \
    const YT zero = (YT)0;                                                         \
    const int64_t nnz = nrow > 0 ? (int64_t)row_map[nrow] : 0;                     \
    (void)threads;                                                                 \
    if (alpha == zero || nrow == 0 || ncol == 0 || nnz == 0) {                     \
      if (beta == zero) { for (int i = 0; i < nrow; ++i) y[i] = zero; }            \
      else { for (int i = 0; i < nrow; ++i) y[i] = beta * y[i]; }                  \
      return;                                                                      \
    }                                                                              \
    const int dobeta0 = (beta == zero);                                            \
    const int dyn = nnz > 10000000;                                                \
    if (dyn) {                                                                     \
      _Pragma("omp parallel for schedule(dynamic, 64) num_threads(threads)")       \
      for (int iRow = 0; iRow < nrow; ++iRow) {                                    \
        YT sum = 0;                                                                \
        for (int k = row_map[iRow]; k < row_map[iRow + 1]; ++k)                    \
          sum += values[k] * x[col_idx[k]];                                        \
        sum *= alpha;                                                              \
        y[iRow] = dobeta0 ? sum : beta * y[iRow] + sum;                            \
      }                                                                            \
    } else {                                                                       \
      _Pragma("omp parallel for schedule(static) num_threads(threads)")            \
      for (int iRow = 0; iRow < nrow; ++iRow) {                                    \
        YT sum = 0;                                                                \
        for (int k = row_map[iRow]; k < row_map[iRow + 1]; ++k)                    \
          sum += values[k] * x[col_idx[k]];                                        \
        sum *= alpha;                                                              \
        y[iRow] = dobeta0 ? sum : beta * y[iRow] + sum;                            \
      }                                                                            \
    }                                                                              \
  }

DEF_SPMV_FUNCTOR(okk_spmv_functor_f64, double, double, double)
DEF_SPMV_FUNCTOR(okk_spmv_functor_f32, float, float, float)

/* ------------------------------------------------------------------------
 * O3: the unit tests' own oracle, Test::sequential_spmv.
 * sparse/unit_test/Test_Sparse_spmv.hpp:106-166: scale y by beta (exact 0
 * when beta==0), then y(row) += alpha*val*x(col) (N,C) or
 * y(col) += alpha*val*x(row) (T,H) per entry in storage order.
 * ylen = length of y (nrow for N/C, ncol for T/H).
 * ---------------------------------------------------------------------- */
#define DEF_SPMV_TEST(NAME, AT, XT, YT)                                            \
  OKK_API void NAME(char mode, int nrow, int ylen, const int* row_map,             \
                    const int* col_idx, const AT* values, const XT* x, YT* y,      \
                    YT alpha, YT beta) {                                           \
    for (int i = 0; i < ylen; ++i) {                                               \
      if (beta == (YT)0) y[i] = (YT)0; else y[i] *= beta;                          \
    }                                                                              \
    const int trans = (mode == 'T' || mode == 'H' || mode == 't' || mode == 'h');  \
    for (int row = 0; row < nrow; ++row) {                                         \
      for (int j = row_map[row]; j < row_map[row + 1]; ++j) {                      \
        const int col = col_idx[j];                                                \
        const AT val = values[j];                                                  \
        if (!trans) y[row] += alpha * val * x[col];                                \
        else y[col] += alpha * val * x[row];                                       \
      }                                                                            \
    }                                                                              \
  }

DEF_SPMV_TEST(okk_spmv_test_f64, double, double, double)
DEF_SPMV_TEST(okk_spmv_test_f32, float, float, float)

/* ------------------------------------------------------------------------
 * O5: Serial transpose path.
 * sparse/impl/KokkosSparse_spmv_impl.hpp:398-452: y is zero-filled (dobeta 0
 * or beta==0) or scaled by beta (dobeta != 1), then for each row i
 * x_val = alpha*x[i]; y[col] += value*x_val in storage order (the 4-way
 * unroll there keeps the order).  alpha==0 leaves the scaled y.
 * Front-end shortcut as in O2.
 * ---------------------------------------------------------------------- */
#define DEF_SPMV_TRANS(NAME, AT, XT, YT)                                           \
  OKK_API void NAME(int nrow, int ncol, const int* row_map, const int* col_idx,    \
                    const AT* values, const XT* x, YT* y, YT alpha, YT beta) {     \
    const YT zero = (YT)0;                                                         \
    const int64_t nnz = nrow > 0 ? (int64_t)row_map[nrow] : 0;                     \
    if (alpha == zero || nrow == 0 || ncol == 0 || nnz == 0) {                     \
      if (beta == zero) { for (int i = 0; i < ncol; ++i) y[i] = zero; }            \
      else { for (int i = 0; i < ncol; ++i) y[i] = beta * y[i]; }                  \
      return;                                                                      \
    }                                                                              \
    if (beta == zero) { for (int i = 0; i < ncol; ++i) y[i] = zero; }              \
    else if (beta != (YT)1) { for (int i = 0; i < ncol; ++i) y[i] = beta * y[i]; } \
    for (int i = 0; i < nrow; ++i) {                                               \
      const XT x_val = alpha * x[i];                                               \
      for (int j = row_map[i]; j < row_map[i + 1]; ++j)                            \
        y[col_idx[j]] += values[j] * x_val;                                        \
    }                                                                              \
  }

DEF_SPMV_TRANS(okk_spmv_transpose_f64, double, double, double)
DEF_SPMV_TRANS(okk_spmv_transpose_f32, float, float, float)

/* ------------------------------------------------------------------------
 * O4: multivector SpMV, CPU (non-team) order.
 * sparse/impl/KokkosSparse_spmv_impl.hpp:745-792 (strip_mine<UNROLL>) and
 * :816-846 (strip_mine_1): every column has its own accumulator filled in
 * storage order.  In strip_mine alpha is folded per term
 * (sum += alpha*val*x unless alpha==+-1, :773-780); strip_mine_1 applies
 * alpha after the sum (:826-830) and is reached only for n == 1 (case 1 of
 * the remainder switch; for n > 16 with n % 16 == 1 the strip of 17 absorbs
 * the odd column).  Strip layout (non-GPU branch :874-925): if n > 16 and
 * n % 16 == 1 a first strip of 17, then strips of 16, then one strip of the
 * remainder (1..15; remainder 1 -> strip_mine_1).
 * beta: 0 -> y=sum, 1 -> y=y+sum, -1 -> y=-y+sum, else beta*y+sum.
 * Layout: X(i,k) = X[i*ldx_r + k*ldx_c], likewise Y (covers LayoutLeft and
 * LayoutRight).  Front end: alpha==0 / empty -> fill 0 or scale
 * (KokkosSparse_spmv.hpp:145-154).
 * ---------------------------------------------------------------------- */
#define DEF_SPMV_MV(NAME, AT, XT, YT)                                              \
  static void NAME##_strip(int iRow, int kk, int U, int post_alpha,                \
                           const int* row_map, const int* col_idx,                 \
                           const AT* values, const XT* X, int64_t xr, int64_t xc,  \
                           YT* Y, int64_t yr, int64_t yc, YT alpha, YT beta) {     \
    YT sum[17];                                                                    \
    for (int k = 0; k < U; ++k) sum[k] = 0;                                        \
    const int doalpha = (alpha == (YT)1) ? 1 : (alpha == (YT)-1 ? -1 : 2);         \
    for (int e = row_map[iRow]; e < row_map[iRow + 1]; ++e) {                      \
      const AT val = values[e];                                                    \
      const int64_t ind = col_idx[e];                                              \
      for (int k = 0; k < U; ++k) {                                                \
        const XT xv = X[ind * xr + (int64_t)(kk + k) * xc];                        \
        if (post_alpha || doalpha == 1) sum[k] += val * xv;                        \
        else if (doalpha == -1) sum[k] -= val * xv;                                \
        else sum[k] += alpha * val * xv;                                           \
      }                                                                            \
    }                                                                              \
    if (post_alpha) { /* strip_mine_1 :826-830 */                                  \
      if (doalpha == -1) sum[0] = -sum[0];                                         \
      else if (doalpha != 1) sum[0] *= alpha;                                      \
    }                                                                              \
    for (int k = 0; k < U; ++k) {                                                  \
      YT* yp = &Y[(int64_t)iRow * yr + (int64_t)(kk + k) * yc];                    \
      if (beta == (YT)0) *yp = sum[k];                                             \
      else if (beta == (YT)1) *yp = *yp + sum[k];                                  \
      else if (beta == (YT)-1) *yp = -*yp + sum[k];                                \
      else *yp = beta * *yp + sum[k];                                              \
    }                                                                              \
  }                                                                                \
  OKK_API void NAME(int nrow, int ncol, int nvec, const int* row_map,              \
                    const int* col_idx, const AT* values, const XT* X,             \
                    int64_t xr, int64_t xc, YT* Y, int64_t yr, int64_t yc,         \
                    YT alpha, YT beta, int threads) {                              \
    const int64_t nnz = nrow > 0 ? (int64_t)row_map[nrow] : 0;                     \
    (void)threads;                                                                 \
    if (alpha == (YT)0 || nrow == 0 || ncol == 0 || nnz == 0) {                    \
      for (int i = 0; i < nrow; ++i)                                               \
        for (int k = 0; k < nvec; ++k) {                                           \
          YT* yp = &Y[(int64_t)i * yr + (int64_t)k * yc];                          \
          *yp = (beta == (YT)0) ? (YT)0 : beta * *yp;                              \
        }                                                                          \
      return;                                                                      \
    }                                                                              \
    _Pragma("omp parallel for schedule(static) num_threads(threads)")              \
    for (int iRow = 0; iRow < nrow; ++iRow) {                                      \
      int kk = 0;                                                                  \
      const int n = nvec;                                                          \
      if ((n > 16) && (n % 16 == 1)) {                                             \
        NAME##_strip(iRow, kk, 17, 0, row_map, col_idx, values, X, xr, xc, Y, yr,  \
                     yc, alpha, beta);                                             \
        kk += 17;                                                                  \
      }                                                                            \
      for (; kk + 16 <= n; kk += 16)                                               \
        NAME##_strip(iRow, kk, 16, 0, row_map, col_idx, values, X, xr, xc, Y, yr,  \
                     yc, alpha, beta);                                             \
      if (kk < n) {                                                                \
        const int rem = n - kk;                                                    \
        NAME##_strip(iRow, kk, rem, rem == 1, row_map, col_idx, values, X, xr, xc, \
                     Y, yr, yc, alpha, beta);                                      \
      }                                                                            \
    }                                                                              \
  }

DEF_SPMV_MV(okk_spmv_mv_f64, double, double, double)
DEF_SPMV_MV(okk_spmv_mv_f32, float, float, float)
DEF_SPMV_MV(okk_spmv_mv_f32mat_f64vec, float, double, double)

/* Multivector transpose, CPU RangePolicy functor run on one thread:
 * sparse/impl/KokkosSparse_spmv_impl.hpp:571-596 + :1130-1160: y scaled by
 * beta first when dobeta != 1 (scal; beta==0 => exact zero fill as the
 * front end documents), then per entry, per column
 * y(ind,k) += alpha*val*x(iRow,k) (doalpha != 1) or val*x(iRow,k). */
#define DEF_SPMV_MV_TRANS(NAME, AT, XT, YT)                                        \
  OKK_API void NAME(int nrow, int ncol, int nvec, const int* row_map,              \
                    const int* col_idx, const AT* values, const XT* X,             \
                    int64_t xr, int64_t xc, YT* Y, int64_t yr, int64_t yc,         \
                    YT alpha, YT beta) {                                           \
    const int64_t nnz = nrow > 0 ? (int64_t)row_map[nrow] : 0;                     \
    const int trivial = (alpha == (YT)0 || nrow == 0 || ncol == 0 || nnz == 0);    \
    if (trivial || beta != (YT)1) {                                                \
      for (int i = 0; i < ncol; ++i)                                               \
        for (int k = 0; k < nvec; ++k) {                                           \
          YT* yp = &Y[(int64_t)i * yr + (int64_t)k * yc];                          \
          *yp = (beta == (YT)0) ? (YT)0 : beta * *yp;                              \
        }                                                                          \
    }                                                                              \
    if (trivial) return;                                                           \
    for (int iRow = 0; iRow < nrow; ++iRow)                                        \
      for (int e = row_map[iRow]; e < row_map[iRow + 1]; ++e) {                    \
        const AT val = values[e];                                                  \
        const int64_t ind = col_idx[e];                                            \
        for (int k = 0; k < nvec; ++k) {                                           \
          const XT xv = X[(int64_t)iRow * xr + (int64_t)k * xc];                   \
          YT* yp = &Y[ind * yr + (int64_t)k * yc];                                 \
          if (alpha != (YT)1) *yp += (YT)(alpha * val * xv);                       \
          else *yp += (YT)(val * xv);                                              \
        }                                                                          \
      }                                                                            \
  }

DEF_SPMV_MV_TRANS(okk_spmv_mv_transpose_f64, double, double, double)
DEF_SPMV_MV_TRANS(okk_spmv_mv_transpose_f32, float, float, float)

/* ------------------------------------------------------------------------
 * O2b: spmv_raw_openmp_no_transpose, sparse/impl/KokkosSparse_spmv_impl_omp.hpp:20-78 (taken by spmv_beta_no_transpose only
 * for double on OpenMP when the graph carries row_block_offsets with omp_get_max_threads()+1 entries and x, y are 64-byte
 * aligned, spmv_impl.hpp:307-320).  Thread t owns rows [block_offsets[t], block_offsets[t+1]); per row
 *   sum = 0;  sum += (alpha * a_ij) * x_j  in storage order;  y_i = (beta == 0) ? sum : beta * y_i + sum
 * -- alpha is folded into every coefficient (unlike the functor O2, which scales the finished sum), so the two paths
 * agree bit for bit only for alpha == 1 (tests/test_oracle_spmv.py pins that, and the tolerance law otherwise).
 * ---------------------------------------------------------------------- */
OKK_API void okk_spmv_raw_openmp_f64(int nblocks, const int* block_offsets, const int* row_map, const int* col_idx,
                                     const double* vals, const double* x, double* y, double s_a, double s_b) {
  const double zero = 0;
#pragma omp parallel for schedule(static, 1) num_threads(nblocks > 0 ? nblocks : 1)
  for (int myID = 0; myID < nblocks; ++myID) {
    const int myStart = block_offsets[myID];
    const int myEnd = block_offsets[myID + 1];
    for (int row = myStart; row < myEnd; ++row) {
      const int rowStart = row_map[row];
      const int rowEnd = row_map[row + 1];
      double sum = 0.0;
      for (int i = rowStart; i < rowEnd; ++i) {
        const int x_entry = col_idx[i];
        const double alpha_MC = s_a * vals[i];
        sum += alpha_MC * x[x_entry];
      }
      if (zero == s_b) {
        y[row] = sum;
      } else {
        y[row] = s_b * y[row] + sum;
      }
    }
  }
}

/* ------------------------------------------------------------------------
 * Merge matrix (merge-path SpMV partitioning).
 * sparse/impl/KokkosSparse_merge_matrix.hpp:80-227.
 * M[i,j] = 1 iff a[i] > b[j]; diagonal d holds `size()` entries counted from
 * the bottom-left; out-of-range on the a side reads 1, on the b side 0.
 * b == NULL means b = iota(nb) (0,1,2,...), the SpMV case (:262-272).
 * ---------------------------------------------------------------------- */
OKK_API int64_t okk_mmd_size(int64_t na, int64_t nb, int64_t d) {
  /* :167-177 */
  if (d <= na && d <= nb) return d;
  else if (d > na && d > nb) return na + nb - d;
  else return na < nb ? na : nb;
}

static void mmd_diag_to_a_b(int64_t na, int64_t d, int64_t di, int64_t* ai, int64_t* bi) {
  /* :186-192 */
  *ai = d < na ? (d - 1) - di : na - 1 - di;
  *bi = d < na ? di : d + di - na;
}

OKK_API int okk_mmd_entry(const int64_t* a, int64_t na, const int64_t* b, int64_t nb,
                          int64_t d, int64_t di) {
  /* :148-159 */
  int64_t ai, bi;
  mmd_diag_to_a_b(na, d, di, &ai, &bi);
  if (ai >= na) return 1;
  else if (bi >= nb) return 0;
  else return a[ai] > (b ? b[bi] : bi);
}

/* diagonal_search (:199-227): lower bound over the diagonal for the first
 * entry that is not 1, then MergeMatrixDiagonal::position (:124-135). */
OKK_API void okk_diagonal_search(const int64_t* a, int64_t na, const int64_t* b, int64_t nb,
                                 int64_t d, int64_t* ai_out, int64_t* bi_out) {
  int64_t lo = 0, len = okk_mmd_size(na, nb, d);
  /* lower_bound_thread with Equal<bool>(x, true): first idx where entry != 1 */
  while (len > 0) {
    int64_t half = len / 2;
    if (okk_mmd_entry(a, na, b, nb, d, lo + half)) { lo += half + 1; len -= half + 1; }
    else len = half;
  }
  if (d == 0) { *ai_out = 0; *bi_out = 0; return; }
  int64_t ai, bi;
  mmd_diag_to_a_b(na, d, lo, &ai, &bi);
  *ai_out = ai + 1;
  *bi_out = bi;
}

/* ------------------------------------------------------------------------
 * O6: SpGEMM, the reference's SPGEMM_DEBUG host Gustavson
 * (default algorithm on Serial/OpenMP, spgemm_handle.hpp:563-588).
 * symbolic: sparse/impl/KokkosSparse_spgemm_impl_seq.hpp:23-97
 * numeric : :99-182 -- accumulator[b_col] += b_val * val iterating A's row
 *           then B's row in storage order, C columns in first-touch order,
 * followed by sort_crs_matrix (spgemm_numeric_spec.hpp:138-140).
 * k = number of columns of B/C.  Returns c_nnz.
 * ---------------------------------------------------------------------- */
OKK_API int64_t okk_spgemm_symbolic(int m, int k, const int* rmA, const int* entA,
                                    const int* rmB, const int* entB, int* rmC) {
  unsigned char* acc_flag = (unsigned char*)calloc((size_t)(k > 0 ? k : 1), 1);
  int* cols = (int*)malloc(sizeof(int) * (size_t)(k > 0 ? k : 1));
  int64_t result_index = 0;
  rmC[0] = 0;
  for (int i = 0; i < m; ++i) {
    int row_size = 0;
    for (int ja = rmA[i]; ja < rmA[i + 1]; ++ja) {
      const int col = entA[ja];
      for (int jb = rmB[col]; jb < rmB[col + 1]; ++jb) {
        const int b_col = entB[jb];
        if (!acc_flag[b_col]) { acc_flag[b_col] = 1; cols[row_size++] = b_col; }
      }
    }
    result_index += row_size;
    rmC[i + 1] = (int)result_index;
    for (int j = 0; j < row_size; ++j) acc_flag[cols[j]] = 0;
  }
  free(acc_flag); free(cols);
  return result_index;
}

#define DEF_SPGEMM_NUMERIC(NAME, ST)                                               \
  OKK_API void NAME(int m, int k, const int* rmA, const int* entA, const ST* valA, \
                    const int* rmB, const int* entB, const ST* valB,               \
                    const int* rmC, int* entC, ST* valC) {                         \
    ST* accumulator = (ST*)calloc((size_t)(k > 0 ? k : 1), sizeof(ST));            \
    unsigned char* acc_flag = (unsigned char*)calloc((size_t)(k > 0 ? k : 1), 1);  \
    for (int i = 0; i < m; ++i) {                                                  \
      const int c_row_begin = rmC[i];                                              \
      const int c_row_size = rmC[i + 1] - c_row_begin;                             \
      int counter = 0;                                                             \
      for (int ja = rmA[i]; ja < rmA[i + 1]; ++ja) {                               \
        const int col = entA[ja];                                                  \
        const ST val = valA[ja];                                                   \
        for (int jb = rmB[col]; jb < rmB[col + 1]; ++jb) {                         \
          const int b_col = entB[jb];                                              \
          const ST b_val = valB[jb];                                               \
          if (!acc_flag[b_col]) {                                                  \
            acc_flag[b_col] = 1;                                                   \
            entC[c_row_begin + counter++] = b_col;                                 \
          }                                                                        \
          accumulator[b_col] += b_val * val;                                       \
        }                                                                          \
      }                                                                            \
      for (int j = 0; j < c_row_size; ++j) {                                       \
        const int c = entC[c_row_begin + j];                                       \
        valC[c_row_begin + j] = accumulator[c];                                    \
        accumulator[c] = 0;                                                        \
        acc_flag[c] = 0;                                                           \
      }                                                                            \
    }                                                                              \
    free(accumulator); free(acc_flag);                                             \
  }

DEF_SPGEMM_NUMERIC(okk_spgemm_numeric_f64, double)
DEF_SPGEMM_NUMERIC(okk_spgemm_numeric_f32, float)

/* sort_crs_matrix (sparse/src/KokkosSparse_SortCrs.hpp:43-120): every row
 * sorted by column index, values permuted along.  Rows of a product have no
 * duplicate columns so the sort is a unique permutation. */
#define DEF_SORT_CRS(NAME, ST)                                                     \
  typedef struct { int c; ST v; } NAME##_pair;                                     \
  static int NAME##_cmp(const void* a, const void* b) {                            \
    const int ca = ((const NAME##_pair*)a)->c, cb = ((const NAME##_pair*)b)->c;    \
    return (ca > cb) - (ca < cb);                                                  \
  }                                                                                \
  OKK_API void NAME(int m, const int* rm, int* ent, ST* val) {                     \
    int maxlen = 0;                                                                \
    for (int i = 0; i < m; ++i) if (rm[i + 1] - rm[i] > maxlen) maxlen = rm[i + 1] - rm[i]; \
    NAME##_pair* buf = (NAME##_pair*)malloc(sizeof(NAME##_pair) * (size_t)(maxlen + 1)); \
    for (int i = 0; i < m; ++i) {                                                  \
      const int b = rm[i], n = rm[i + 1] - rm[i];                                  \
      int sorted = 1;                                                              \
      for (int j = 1; j < n; ++j) if (ent[b + j - 1] > ent[b + j]) { sorted = 0; break; } \
      if (sorted) continue;                                                        \
      for (int j = 0; j < n; ++j) { buf[j].c = ent[b + j]; buf[j].v = val ? val[b + j] : (ST)0; } \
      qsort(buf, (size_t)n, sizeof(NAME##_pair), NAME##_cmp);                      \
      for (int j = 0; j < n; ++j) { ent[b + j] = buf[j].c; if (val) val[b + j] = buf[j].v; } \
    }                                                                              \
    free(buf);                                                                     \
  }

DEF_SORT_CRS(okk_sort_crs_f64, double)
DEF_SORT_CRS(okk_sort_crs_f32, float)

/* Row block [r0, r1) of the same product, rows dealt to OpenMP threads (each with its own dense accumulator): per row the very
 * loops of okk_spgemm_symbolic / okk_spgemm_numeric (impl_seq.hpp:23-182) followed by the row sort (numeric_spec.hpp:138-140),
 * so every row holds the bits the serial functions give (rows of a product are independent).  Lets the full-size parity tests
 * (config 4: 2M rows, 2e9 products) check ALL rows in blocks without materialising C on the host.
 * rowlen[r1-r0] out; ent/val (capacity cap entries) receive the block's rows back to back; returns the block's nnz, or -1 when
 * cap is too small (nothing written beyond rowlen). */
#define DEF_SPGEMM_BLOCK(NAME, ST, PAIR, CMP)                                      \
  OKK_API int64_t NAME(int r0, int r1, int k, const int* rmA, const int* entA, const ST* valA, \
                       const int* rmB, const int* entB, const ST* valB, int64_t cap, \
                       int* rowlen, int* ent, ST* val, int threads) {              \
    const int nr = r1 - r0;                                                        \
    if (threads < 1) threads = 1;                                                  \
    _Pragma("omp parallel num_threads(threads)")                                   \
    {                                                                              \
      unsigned char* acc_flag = (unsigned char*)calloc((size_t)(k > 0 ? k : 1), 1); \
      int* cols = (int*)malloc(sizeof(int) * (size_t)(k > 0 ? k : 1));             \
      _Pragma("omp for schedule(dynamic, 64)")                                     \
      for (int q = 0; q < nr; ++q) {                                               \
        const int i = r0 + q;                                                      \
        int row_size = 0;                                                          \
        for (int ja = rmA[i]; ja < rmA[i + 1]; ++ja) {                             \
          const int col = entA[ja];                                                \
          for (int jb = rmB[col]; jb < rmB[col + 1]; ++jb) {                       \
            const int b_col = entB[jb];                                            \
            if (!acc_flag[b_col]) { acc_flag[b_col] = 1; cols[row_size++] = b_col; } \
          }                                                                        \
        }                                                                          \
        rowlen[q] = row_size;                                                      \
        for (int j = 0; j < row_size; ++j) acc_flag[cols[j]] = 0;                  \
      }                                                                            \
      free(acc_flag); free(cols);                                                  \
    }                                                                              \
    int64_t total = 0;                                                             \
    for (int q = 0; q < nr; ++q) total += rowlen[q];                               \
    if (total > cap) return -1;                                                    \
    int64_t* start = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nr + 1));         \
    start[0] = 0;                                                                  \
    for (int q = 0; q < nr; ++q) start[q + 1] = start[q] + rowlen[q];              \
    _Pragma("omp parallel num_threads(threads)")                                   \
    {                                                                              \
      ST* accumulator = (ST*)calloc((size_t)(k > 0 ? k : 1), sizeof(ST));          \
      unsigned char* acc_flag = (unsigned char*)calloc((size_t)(k > 0 ? k : 1), 1); \
      PAIR* buf = NULL;                                                            \
      int bufcap = 0;                                                              \
      _Pragma("omp for schedule(dynamic, 64)")                                     \
      for (int q = 0; q < nr; ++q) {                                               \
        const int i = r0 + q;                                                      \
        int* entC = ent + start[q];                                                \
        ST* valC = val + start[q];                                                 \
        const int c_row_size = rowlen[q];                                          \
        int counter = 0;                                                           \
        for (int ja = rmA[i]; ja < rmA[i + 1]; ++ja) {                             \
          const int col = entA[ja];                                                \
          const ST v = valA[ja];                                                   \
          for (int jb = rmB[col]; jb < rmB[col + 1]; ++jb) {                       \
            const int b_col = entB[jb];                                            \
            const ST b_val = valB[jb];                                             \
            if (!acc_flag[b_col]) { acc_flag[b_col] = 1; entC[counter++] = b_col; } \
            accumulator[b_col] += b_val * v;                                       \
          }                                                                        \
        }                                                                          \
        if (c_row_size > bufcap) {                                                 \
          bufcap = c_row_size * 2;                                                 \
          buf = (PAIR*)realloc(buf, sizeof(PAIR) * (size_t)bufcap);                \
        }                                                                          \
        for (int j = 0; j < c_row_size; ++j) {                                     \
          const int c = entC[j];                                                   \
          buf[j].c = c;                                                            \
          buf[j].v = accumulator[c];                                               \
          accumulator[c] = 0;                                                      \
          acc_flag[c] = 0;                                                         \
        }                                                                          \
        qsort(buf, (size_t)c_row_size, sizeof(PAIR), CMP);                         \
        for (int j = 0; j < c_row_size; ++j) { entC[j] = buf[j].c; valC[j] = buf[j].v; } \
      }                                                                            \
      free(accumulator); free(acc_flag); free(buf);                                \
    }                                                                              \
    free(start);                                                                   \
    return total;                                                                  \
  }

DEF_SPGEMM_BLOCK(okk_spgemm_block_f64, double, okk_sort_crs_f64_pair, okk_sort_crs_f64_cmp)
DEF_SPGEMM_BLOCK(okk_spgemm_block_f32, float, okk_sort_crs_f32_pair, okk_sort_crs_f32_cmp)

/* transpose_matrix (sparse/src/KokkosSparse_Utils.hpp:338) -- counting sort by
 * column, stable in row order; used by the issue-402 fixture (C = A*A^T). */
OKK_API void okk_transpose_f64(int nrow, int ncol, const int* rm, const int* ent,
                               const double* val, int* trm, int* tent, double* tval) {
  memset(trm, 0, sizeof(int) * (size_t)(ncol + 1));
  for (int j = 0; j < rm[nrow]; ++j) trm[ent[j] + 1]++;
  for (int c = 0; c < ncol; ++c) trm[c + 1] += trm[c];
  int* pos = (int*)malloc(sizeof(int) * (size_t)(ncol + 1));
  memcpy(pos, trm, sizeof(int) * (size_t)(ncol + 1));
  for (int i = 0; i < nrow; ++i)
    for (int j = rm[i]; j < rm[i + 1]; ++j) {
      const int p = pos[ent[j]]++;
      tent[p] = i;
      tval[p] = val[j];
    }
  free(pos);
}

/* is_same_matrix value law (sparse/unit_test/Test_Sparse_Utils.hpp:86-118,
 * common/src/KokkosKernels_SimpleUtils.hpp:262-330): entries are "relatively
 * identical" when both |a|,|b| <= eps or |a-b|/(|a|+|b|) <= eps.
 * Returns the number of violating entries. */
OKK_API int64_t okk_count_rel_mismatch_f64(int64_t n, const double* a, const double* b, double eps) {
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double aa = fabs(a[i]), bb = fabs(b[i]);
    if (aa <= eps && bb <= eps) continue;
    if (!(fabs(a[i] - b[i]) / (aa + bb) <= eps)) bad++;
  }
  return bad;
}

/* First-touch placement for the CPU timing legs (bench.py cpu_baseline / --impl reference): copies `bytes`
 * from src to the untouched allocation dst with an OpenMP static partition, so that the pages of dst are
 * spread over the NUMA nodes of the threads that will stream them -- what a Kokkos application gets from
 * initialising its views in parallel (reference protocol
 * perf_test/sparse/KokkosSparse_kk_spmv.cpp:121-167).  Not part of any result. */
OKK_API void okk_parallel_copy(void* dst, const void* src, int64_t bytes, int threads) {
  const int64_t chunk = 1 << 16;
  const int64_t nchunks = (bytes + chunk - 1) / chunk;
  (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t o = c * chunk;
    const int64_t n = bytes - o < chunk ? bytes - o : chunk;
    memcpy((char*)dst + o, (const char*)src + o, (size_t)n);
  }
}

OKK_API int okk_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
