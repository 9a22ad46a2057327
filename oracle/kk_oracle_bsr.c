/*
 * kk_oracle_bsr.c -- CPU restatement of the reference's BsrMatrix SpMV.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c): only tests/,
 * __graft_entry__.smoke() and tools/gpu_check may load it.
 *
 * BsrMatrix (sparse/src/KokkosSparse_BsrMatrix.hpp:355-370): block row map (mb+1), block column
 * indices (nnzb), values nnzb*bs*bs with every block stored row-major (LayoutRight), block K of the
 * matrix at values[K*bs*bs], element (i,j) at +i*bs+j (:242-253).
 *
 * Restated paths (paths relative to /root/reference):
 *   B1  BsrSpmvV42NonTrans      sparse/impl/KokkosSparse_spmv_bsrmatrix_impl_v42.hpp:45-91
 *       -- the native mode-N path of every GPU execution space (spec.hpp:261-266) and of
 *       SPMV_BSR_V42 on the host; one work item per entry of y.
 *   B2  BSR_GEMV_Functor        sparse/impl/KokkosSparse_spmv_bsrmatrix_impl.hpp:508-541 + :594-632
 *       -- Serial / OpenMP default for modes N and C (V41): y scaled first, then one Blocked serial
 *       gemv per block (blas/impl/KokkosBlas2_serial_gemv_internal.hpp:84-113, inner dot products
 *       KokkosBlas2_serial_gemv_inner_multiple_dot.hpp:87-118: t = sum_j a(i,j)*x(j); y(i) += alpha*t).
 *   B3  BSR_GEMV_Transpose_Functor  ...bsrmatrix_impl.hpp:737-775 + :843-882  (modes T and H, host)
 *   B4  multivector forms: B1 per column (same functor, irhs = k / y.extent(0)); V41
 *       BSR_GEMM_Functor (:1001-1043) restated as its conjugate-branch loop nest (the non-conjugate branch calls
 *       a blocked serial GEMM whose summation order over the block is implementation-defined tiling; parity with
 *       it is by the unit test's tolerance law, not bit-for-bit).
 * For real scalars conj() is the identity, so C == N and H == T.
 *
 * Pinned (tests/test_oracle_bsr.py): B1 must equal, bit for bit, (a) the reference's own functor compiled from
 * the reference tree (oracle/_ref/libkkref.so, kkref_bsr.cpp) and (b) the already pinned CrsMatrix functor
 * order O2 (kk_oracle.c) applied to bsr_to_crs(A) -- the comparison the reference's unit test makes
 * (sparse/unit_test/Test_Sparse_spmv_bsr.hpp:142-213) with its tolerance law; B2/B3 by that law.
 *
 * Index types: Ordinal = Offset = int32.  Compiled with -ffp-contract=off.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OKK_API __attribute__((visibility("default")))

/* B1.  X, Y addressed as base[row*rs + col*cs] (rank 1: nvec = 1). */
#define DEF_BSR_V42(NAME, T)                                                                        \
  OKK_API void NAME(int mb, int bs, int nvec, const int* row_map, const int* entries,               \
                    const T* values, const T* X, int64_t xr, int64_t xc, T* Y, int64_t yr,          \
                    int64_t yc, T alpha, T beta) {                                                  \
    const int64_t nrow = (int64_t)mb * bs;                                                          \
    for (int64_t k = 0; k < nrow * nvec; ++k) {                                                     \
      const int64_t irhs = k / nrow, row = k % nrow;                                                \
      T* yp = &Y[row * yr + irhs * yc];                                                             \
      if (beta == (T)0) *yp = (T)0;                                                                 \
      else if (beta != (T)1) *yp *= beta;                                                           \
      if (alpha != (T)0) {                                                                          \
        const int blockRow = (int)(row / bs), lclrow = (int)(row % bs);                             \
        T accum = (T)0;                                                                             \
        for (int j = row_map[blockRow]; j < row_map[blockRow + 1]; ++j) {                           \
          const T* b = values + (int64_t)j * bs * bs;                                               \
          const int64_t x_start = (int64_t)entries[j] * bs;                                         \
          for (int i = 0; i < bs; ++i) accum += b[lclrow * bs + i] * X[(x_start + i) * xr + irhs * xc]; \
        }                                                                                           \
        *yp += alpha * accum;                                                                       \
      }                                                                                             \
    }                                                                                               \
  }

DEF_BSR_V42(okk_bsr_spmv_v42_f64, double)
DEF_BSR_V42(okk_bsr_spmv_v42_f32, float)

/* B2 / B3 (+ their multivector forms).  ylen_b = block rows of y: mb for N/C, nb (block columns) for T/H. */
#define DEF_BSR_V41(NAME, T)                                                                        \
  OKK_API void NAME(char mode, int mb, int ylen_b, int bs, int nvec, const int* row_map,            \
                    const int* entries, const T* values, const T* X, int64_t xr, int64_t xc, T* Y,  \
                    int64_t yr, int64_t yc, T alpha, T beta) {                                      \
    const int trans = (mode == 'T' || mode == 'H' || mode == 't' || mode == 'h');                   \
    const int64_t ylen = (int64_t)ylen_b * bs;                                                      \
    for (int k = 0; k < nvec; ++k)                                                                  \
      for (int64_t i = 0; i < ylen; ++i) {                                                          \
        T* yp = &Y[i * yr + k * yc];                                                                \
        if (beta == (T)0) *yp = (T)0;                                                               \
        else if (beta != (T)1) *yp = beta * *yp;                                                    \
      }                                                                                             \
    if (trans && alpha == (T)0) return; /* :855 */                                                  \
    for (int iBlock = 0; iBlock < mb; ++iBlock)                                                     \
      for (int ic = row_map[iBlock]; ic < row_map[iBlock + 1]; ++ic) {                              \
        const T* A = values + (int64_t)ic * bs * bs;                                                \
        const int64_t cstart = (int64_t)entries[ic] * bs, rstart = (int64_t)iBlock * bs;            \
        for (int jr = 0; jr < nvec; ++jr) {                                                         \
          if (!trans) {                                                                             \
            for (int ii = 0; ii < bs; ++ii) {                                                       \
              T t = (T)0;                                                                           \
              for (int jj = 0; jj < bs; ++jj) t += A[ii * bs + jj] * X[(cstart + jj) * xr + jr * xc]; \
              Y[(rstart + ii) * yr + jr * yc] += alpha * t;                                         \
            }                                                                                       \
          } else {                                                                                  \
            for (int jj = 0; jj < bs; ++jj) {                                                       \
              T t = (T)0;                                                                           \
              for (int ii = 0; ii < bs; ++ii) t += A[ii * bs + jj] * X[(rstart + ii) * xr + jr * xc]; \
              t *= alpha;                                                                           \
              Y[(cstart + jj) * yr + jr * yc] += t;                                                 \
            }                                                                                       \
          }                                                                                         \
        }                                                                                           \
      }                                                                                             \
  }

DEF_BSR_V41(okk_bsr_spmv_v41_f64, double)
DEF_BSR_V41(okk_bsr_spmv_v41_f32, float)

/* bsr_to_crs (sparse/impl/KokkosSparse_bsr_to_crs_impl.hpp:31-117): the point matrix of a BsrMatrix; the
 * entries of every point row are sorted by column (std::sort by column, :101; blocks of a block row that repeat a
 * block column keep an unspecified relative order there -- here: storage order).  crs_row_map has mb*bs+1 entries,
 * crs_entries / crs_values nnzb*bs*bs. */
#define DEF_BSR_TO_CRS(NAME, T)                                                                     \
  OKK_API void NAME(int mb, int bs, const int* row_map, const int* entries, const T* values,        \
                    int* crs_row_map, int* crs_entries, T* crs_values) {                            \
    int64_t out = 0;                                                                                \
    crs_row_map[0] = 0;                                                                             \
    int cap = 0;                                                                                    \
    for (int b = 0; b < mb; ++b)                                                                    \
      if (row_map[b + 1] - row_map[b] > cap) cap = row_map[b + 1] - row_map[b];                     \
    int* order = (int*)malloc(sizeof(int) * (size_t)(cap > 0 ? cap : 1));                           \
    for (int b = 0; b < mb; ++b) {                                                                  \
      const int s = row_map[b], n = row_map[b + 1] - s;                                             \
      for (int q = 0; q < n; ++q) order[q] = s + q;                                                 \
      for (int q = 1; q < n; ++q) { /* stable insertion sort by block column */                     \
        const int v = order[q];                                                                     \
        int p = q - 1;                                                                              \
        while (p >= 0 && entries[order[p]] > entries[v]) { order[p + 1] = order[p]; --p; }          \
        order[p + 1] = v;                                                                           \
      }                                                                                             \
      for (int lr = 0; lr < bs; ++lr) {                                                             \
        for (int q = 0; q < n; ++q) {                                                               \
          const int j = order[q];                                                                   \
          for (int lc = 0; lc < bs; ++lc) {                                                         \
            crs_entries[out] = entries[j] * bs + lc;                                                \
            crs_values[out] = values[(int64_t)j * bs * bs + lr * bs + lc];                          \
            ++out;                                                                                  \
          }                                                                                         \
        }                                                                                           \
        crs_row_map[(int64_t)b * bs + lr + 1] = (int)out;                                           \
      }                                                                                             \
    }                                                                                               \
    free(order);                                                                                    \
  }

DEF_BSR_TO_CRS(okk_bsr_to_crs_f64, double)
DEF_BSR_TO_CRS(okk_bsr_to_crs_f32, float)
