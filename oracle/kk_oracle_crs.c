/*
 * kk_oracle_crs.c -- CPU restatement of the CrsMatrix utilities either side of
 * the hot path: sort_crs_matrix, sort_and_merge_matrix,
 * spadd (sorted and unsorted input).  (transpose_matrix lives in kk_oracle.c.)
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as kk_oracle.c):
 * only tests/, __graft_entry__.smoke() and tools/gpu_check may load it.
 *
 * Every function restates, operation for operation, the reference's Serial /
 * OpenMP (host) loop it cites (paths relative to /root/reference).  Pinned by
 * the reference's own golden cases: the five sort_and_merge matrices of
 * sparse/unit_test/Test_Sparse_SortCrs.hpp:195-290 and, for spadd, the dense
 * row check of sparse/unit_test/Test_Sparse_spadd.hpp (tests/test_oracle_crs.py).
 *
 * Index types: Ordinal = Offset = int32.  Compiled with -ffp-contract=off.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define OKK_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------
 * SerialRadixSort2 (common/src/KokkosKernels_Sorting.hpp:301-369): LSD radix
 * sort, 4 bits per pass, as many passes as the largest key needs; `perm`
 * follows `values`.  Stable.
 * ---------------------------------------------------------------------- */
#define DEF_RADIX2(NAME, PT)                                                       \
  static void NAME(uint32_t* values, uint32_t* valuesAux, PT* perm, PT* permAux,   \
                   int n) {                                                        \
    if (n <= 1) return;                                                            \
    uint32_t maxVal = 0;                                                           \
    for (int i = 0; i < n; i++)                                                    \
      if (maxVal < values[i]) maxVal = values[i];                                  \
    int passes = 0;                                                                \
    while (maxVal) { maxVal >>= 4; passes++; }                                     \
    int inAux = 0;                                                                 \
    uint32_t mask = 0xF;                                                           \
    int maskPos = 0;                                                               \
    for (int p = 0; p < passes; p++) {                                             \
      int count[16] = {0};                                                         \
      int offset[17];                                                              \
      const uint32_t* src = inAux ? valuesAux : values;                            \
      for (int i = 0; i < n; i++) count[(src[i] & mask) >> maskPos]++;             \
      offset[0] = 0;                                                               \
      for (int i = 0; i < 16; i++) offset[i + 1] = offset[i] + count[i];           \
      if (!inAux) {                                                                \
        for (int i = 0; i < n; i++) {                                              \
          const int bucket = (int)((values[i] & mask) >> maskPos);                 \
          valuesAux[offset[bucket + 1] - count[bucket]] = values[i];               \
          permAux[offset[bucket + 1] - count[bucket]] = perm[i];                   \
          count[bucket]--;                                                         \
        }                                                                          \
      } else {                                                                     \
        for (int i = 0; i < n; i++) {                                              \
          const int bucket = (int)((valuesAux[i] & mask) >> maskPos);              \
          values[offset[bucket + 1] - count[bucket]] = valuesAux[i];               \
          perm[offset[bucket + 1] - count[bucket]] = permAux[i];                   \
          count[bucket]--;                                                         \
        }                                                                          \
      }                                                                            \
      inAux = !inAux;                                                              \
      mask = mask << 4;                                                            \
      maskPos += 4;                                                                \
    }                                                                              \
    if (inAux)                                                                     \
      for (int i = 0; i < n; i++) { values[i] = valuesAux[i]; perm[i] = permAux[i]; } \
  }

DEF_RADIX2(radix2_f64, double)
DEF_RADIX2(radix2_f32, float)
DEF_RADIX2(radix2_i32, int)

/* sort_crs_matrix, host branch (sparse/src/KokkosSparse_SortCrs.hpp:69-74 ->
 * MatrixRadixSortFunctor, sparse/impl/KokkosSparse_sort_crs_impl.hpp:26-57).
 * val == NULL: sort_crs_graph (GraphRadixSortFunctor, :86-112). */
#define DEF_SORT_STABLE(NAME, PT, RADIX)                                           \
  OKK_API void NAME(int m, const int* rm, int* ent, PT* val) {                     \
    const int nnz = m > 0 ? rm[m] : 0;                                             \
    if (nnz <= 1) return; /* SortCrs.hpp:62-67 */                                  \
    uint32_t* entAux = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nnz);          \
    PT* valAux = (PT*)malloc(sizeof(PT) * (size_t)nnz);                            \
    PT* dummy = val ? NULL : (PT*)calloc((size_t)nnz, sizeof(PT));                 \
    PT* v = val ? val : dummy;                                                     \
    for (int i = 0; i < m; i++) {                                                  \
      const int rowStart = rm[i], rowNum = rm[i + 1] - rm[i];                      \
      RADIX((uint32_t*)ent + rowStart, entAux + rowStart, v + rowStart,            \
            valAux + rowStart, rowNum);                                            \
    }                                                                              \
    free(entAux); free(valAux); free(dummy);                                       \
  }

DEF_SORT_STABLE(okk_sort_crs_stable_f64, double, radix2_f64)
DEF_SORT_STABLE(okk_sort_crs_stable_f32, float, radix2_f32)
DEF_SORT_STABLE(okk_sort_crs_stable_i32, int, radix2_i32)

/* MergedRowmapFunctor + exclusive prefix sum (sort_crs_impl.hpp:130-161,
 * SortCrs.hpp:338-343,366): rows must be sorted.  Writes the merged row map
 * (m+1 entries) and returns the merged nnz. */
OKK_API int64_t okk_merged_rowmap(int m, const int* rm, const int* ent, int* rm_out) {
  int64_t total = 0;
  rm_out[0] = 0;
  for (int row = 0; row < m; row++) {
    const int rowBegin = rm[row], rowEnd = rm[row + 1];
    int uniqueEntries = 0;
    if (rowEnd != rowBegin) {
      uniqueEntries = 1;
      for (int j = rowBegin + 1; j < rowEnd; j++)
        if (ent[j - 1] != ent[j]) uniqueEntries++;
    }
    total += uniqueEntries;
    rm_out[row + 1] = (int)total;
  }
  return total;
}

/* MatrixMergedEntriesFunctor / GraphMergedEntriesFunctor (sort_crs_impl.hpp:163-248) */
#define DEF_MERGED_ENTRIES(NAME, ST)                                               \
  OKK_API void NAME(int m, const int* rm, const int* ent, const ST* val,           \
                    const int* rm_out, int* ent_out, ST* val_out) {                \
    for (int row = 0; row < m; row++) {                                            \
      const int rowBegin = rm[row], rowEnd = rm[row + 1];                          \
      if (rowEnd == rowBegin) continue;                                            \
      ST accumVal = val ? val[rowBegin] : (ST)0;                                   \
      int accumCol = ent[rowBegin];                                                \
      int insertPos = rm_out[row];                                                 \
      for (int j = rowBegin + 1; j < rowEnd; j++) {                                \
        if (accumCol == ent[j]) {                                                  \
          if (val) accumVal += val[j];                                             \
        } else {                                                                   \
          if (val) val_out[insertPos] = accumVal;                                  \
          ent_out[insertPos] = accumCol;                                           \
          insertPos++;                                                             \
          if (val) accumVal = val[j];                                              \
          accumCol = ent[j];                                                       \
        }                                                                          \
      }                                                                            \
      if (val) val_out[insertPos] = accumVal;                                      \
      ent_out[insertPos] = accumCol;                                               \
    }                                                                              \
  }

DEF_MERGED_ENTRIES(okk_merged_entries_f64, double)
DEF_MERGED_ENTRIES(okk_merged_entries_f32, float)

/* ------------------------------------------------------------------------
 * spadd, sorted input: SortedCountEntriesRange + prefix sum
 * (sparse/impl/KokkosSparse_spadd_symbolic_impl.hpp:33-77,463-467) and
 * SortedNumericSumFunctor (sparse/impl/KokkosSparse_spadd_numeric_impl.hpp:27-107).
 * ---------------------------------------------------------------------- */
OKK_API int64_t okk_spadd_sorted_symbolic(int m, const int* rmA, const int* entA, const int* rmB,
                                          const int* entB, int* rmC) {
  const int ORDINAL_MAX = INT_MAX;
  int64_t total = 0;
  rmC[0] = 0;
  for (int i = 0; i < m; i++) {
    int numEntries = 0;
    int ai = 0, bi = 0;
    const int Arowstart = rmA[i], Arowlen = rmA[i + 1] - Arowstart;
    const int Browstart = rmB[i], Browlen = rmB[i + 1] - Browstart;
    int Acol = (Arowlen == 0) ? ORDINAL_MAX : entA[Arowstart];
    int Bcol = (Browlen == 0) ? ORDINAL_MAX : entB[Browstart];
    /* the reference pre-loads entry 0 and then reads entry ai++ again (:62-63); the net effect is
       "skip every entry equal to Ccol", restated here with the same reads */
    while (Acol != ORDINAL_MAX || Bcol != ORDINAL_MAX) {
      const int Ccol = (Acol < Bcol) ? Acol : Bcol;
      numEntries++;
      while (Acol == Ccol) Acol = (ai == Arowlen) ? ORDINAL_MAX : entA[Arowstart + ai++];
      while (Bcol == Ccol) Bcol = (bi == Browlen) ? ORDINAL_MAX : entB[Browstart + bi++];
    }
    total += numEntries;
    rmC[i + 1] = (int)total;
  }
  return total;
}

#define DEF_SPADD_SORTED_NUMERIC(NAME, ST)                                         \
  OKK_API void NAME(int m, const int* rmA, const int* entA, const ST* valA,        \
                    ST alpha, const int* rmB, const int* entB, const ST* valB,     \
                    ST beta, const int* rmC, int* entC, ST* valC) {                \
    const int ORDINAL_MAX = INT_MAX;                                               \
    for (int i = 0; i < m; i++) {                                                  \
      int ai = 0, bi = 0;                                                          \
      const int Arowstart = rmA[i], Arowlen = rmA[i + 1] - Arowstart;              \
      const int Browstart = rmB[i], Browlen = rmB[i + 1] - Browstart;              \
      int Acol = (Arowlen == 0) ? ORDINAL_MAX : entA[Arowstart];                   \
      int Bcol = (Browlen == 0) ? ORDINAL_MAX : entB[Browstart];                   \
      int Coffset = rmC[i];                                                        \
      while (Acol != ORDINAL_MAX || Bcol != ORDINAL_MAX) {                         \
        const int Ccol = (Acol < Bcol) ? Acol : Bcol;                              \
        ST accum = (ST)0;                                                          \
        while (Acol == Ccol) {                                                     \
          accum += (ST)(alpha * valA[Arowstart + ai]);                             \
          ai++;                                                                    \
          Acol = (ai == Arowlen) ? ORDINAL_MAX : entA[Arowstart + ai];             \
        }                                                                          \
        while (Bcol == Ccol) {                                                     \
          accum += (ST)(beta * valB[Browstart + bi]);                              \
          bi++;                                                                    \
          Bcol = (bi == Browlen) ? ORDINAL_MAX : entB[Browstart + bi];             \
        }                                                                          \
        entC[Coffset] = Ccol;                                                      \
        valC[Coffset] = accum;                                                     \
        Coffset++;                                                                 \
      }                                                                            \
    }                                                                              \
  }

DEF_SPADD_SORTED_NUMERIC(okk_spadd_sorted_numeric_f64, double)
DEF_SPADD_SORTED_NUMERIC(okk_spadd_sorted_numeric_f32, float)

/* ------------------------------------------------------------------------
 * spadd, unsorted input (spadd_symbolic_impl.hpp:468-503): upper bound row
 * map, UnmergedSumFunctor (:232-276), sort_crs_matrix of (columns, A/B
 * permutation), MergeEntriesFunctor (:278-343) -> Apos / Bpos / row counts,
 * prefix sum.  Returns nnz(C).  apos has nnz(A) entries, bpos nnz(B).
 * ---------------------------------------------------------------------- */
OKK_API int64_t okk_spadd_unsorted_symbolic(int m, const int* rmA, const int* entA, const int* rmB,
                                            const int* entB, int* rmC, int* apos, int* bpos) {
  int* rmU = (int*)malloc(sizeof(int) * (size_t)(m + 1));
  rmU[0] = 0;
  for (int i = 0; i < m; i++) rmU[i + 1] = rmU[i] + (rmA[i + 1] - rmA[i]) + (rmB[i + 1] - rmB[i]);
  const int ub = rmU[m];
  int* entU = (int*)malloc(sizeof(int) * (size_t)(ub > 0 ? ub : 1));
  int* perm = (int*)malloc(sizeof(int) * (size_t)(ub > 0 ? ub : 1));
  for (int i = 0; i < m; i++) {
    int inserted = 0;
    const int crowstart = rmU[i];
    const int arowstart = rmA[i], arowlen = rmA[i + 1] - arowstart;
    const int browstart = rmB[i], browlen = rmB[i + 1] - browstart;
    for (int j = 0; j < arowlen; j++) {
      entU[crowstart + inserted] = entA[arowstart + j];
      perm[crowstart + inserted] = j;
      inserted++;
    }
    for (int j = 0; j < browlen; j++) {
      entU[crowstart + inserted] = entB[browstart + j];
      perm[crowstart + inserted] = j + arowlen;
      inserted++;
    }
  }
  okk_sort_crs_stable_i32(m, rmU, entU, perm);
  int64_t total = 0;
  rmC[0] = 0;
  for (int i = 0; i < m; i++) {
    const int CrowStart = rmU[i], CrowEnd = rmU[i + 1];
    int count = 0;
    if (CrowEnd != CrowStart) {
      const int ArowStart = rmA[i], ArowNum = rmA[i + 1] - ArowStart;
      const int BrowStart = rmB[i];
      int CFit = 0;
      for (int Cit = CrowStart; Cit < CrowEnd; Cit++) {
        if ((Cit > CrowStart) && (entU[Cit] != entU[Cit - 1])) CFit++;
        const int permVal = perm[Cit];
        if (permVal < ArowNum) apos[ArowStart + permVal] = CFit;
        else bpos[BrowStart + (permVal - ArowNum)] = CFit;
      }
      count = CFit + 1;
    }
    total += count;
    rmC[i + 1] = (int)total;
  }
  free(rmU); free(entU); free(perm);
  return total;
}

/* UnsortedNumericSumFunctor (spadd_numeric_impl.hpp:109-171) */
#define DEF_SPADD_UNSORTED_NUMERIC(NAME, ST)                                       \
  OKK_API void NAME(int m, const int* rmA, const int* entA, const ST* valA,        \
                    ST alpha, const int* rmB, const int* entB, const ST* valB,     \
                    ST beta, const int* rmC, int* entC, ST* valC,                  \
                    const int* apos, const int* bpos) {                            \
    for (int i = 0; i < m; i++) {                                                  \
      const int CrowStart = rmC[i], CrowEnd = rmC[i + 1];                          \
      for (int j = CrowStart; j < CrowEnd; j++) valC[j] = (ST)0;                   \
      for (int j = rmA[i]; j < rmA[i + 1]; j++) {                                  \
        valC[CrowStart + apos[j]] += alpha * valA[j];                              \
        entC[CrowStart + apos[j]] = entA[j];                                       \
      }                                                                            \
      for (int j = rmB[i]; j < rmB[i + 1]; j++) {                                  \
        valC[CrowStart + bpos[j]] += beta * valB[j];                               \
        entC[CrowStart + bpos[j]] = entB[j];                                       \
      }                                                                            \
    }                                                                              \
  }

DEF_SPADD_UNSORTED_NUMERIC(okk_spadd_unsorted_numeric_f64, double)
DEF_SPADD_UNSORTED_NUMERIC(okk_spadd_unsorted_numeric_f32, float)

/* ------------------------------------------------------------------------
 * spgemm_jacobi_seq (sparse/impl/KokkosSparse_spgemm_jacobi_seq_impl.hpp:26-131):
 * C = (I - omega*diag(dinv)*A)*B on the row map of spgemm_symbolic(A, B): per
 * row, B's row i is inserted first (weight 1), then every product with
 * val = a_ij * (-omega*dinv_i), b_val = b * val; columns in first-touch order
 * (the caller sorts, like the unit test does for the SPGEMM_SERIAL result,
 * sparse/unit_test/Test_Sparse_spgemm_jacobi.hpp:216-219).
 * ---------------------------------------------------------------------- */
#define DEF_SPGEMM_JACOBI(NAME, ST)                                                \
  OKK_API void NAME(int m, int k, const int* rmA, const int* entA, const ST* valA, \
                    const int* rmB, const int* entB, const ST* valB,               \
                    const int* rmC, int* entC, ST* valC, ST omega,                 \
                    const ST* dinv) {                                              \
    ST* accumulator = (ST*)calloc((size_t)(k > 0 ? k : 1), sizeof(ST));            \
    unsigned char* acc_flag = (unsigned char*)calloc((size_t)(k > 0 ? k : 1), 1);  \
    for (int i = 0; i < m; ++i) {                                                  \
      const int c_row_begin = rmC[i];                                              \
      const int c_row_size = rmC[i + 1] - c_row_begin;                             \
      int counter = 0;                                                             \
      const ST mult = -omega * dinv[i];                                            \
      for (int z = rmB[i]; z < rmB[i + 1]; ++z) {                                  \
        const int b_col = entB[z];                                                 \
        const ST b_val = valB[z];                                                  \
        if (!acc_flag[b_col]) {                                                    \
          acc_flag[b_col] = 1;                                                     \
          entC[c_row_begin + counter++] = b_col;                                   \
        }                                                                          \
        accumulator[b_col] += b_val;                                               \
      }                                                                            \
      for (int ja = rmA[i]; ja < rmA[i + 1]; ++ja) {                               \
        const int col = entA[ja];                                                  \
        const ST val = valA[ja] * mult;                                            \
        for (int jb = rmB[col]; jb < rmB[col + 1]; ++jb) {                         \
          const int b_col = entB[jb];                                              \
          const ST b_val = valB[jb] * val;                                         \
          if (!acc_flag[b_col]) {                                                  \
            acc_flag[b_col] = 1;                                                   \
            entC[c_row_begin + counter++] = b_col;                                 \
          }                                                                        \
          accumulator[b_col] += b_val;                                             \
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

DEF_SPGEMM_JACOBI(okk_spgemm_jacobi_f64, double)
DEF_SPGEMM_JACOBI(okk_spgemm_jacobi_f32, float)
