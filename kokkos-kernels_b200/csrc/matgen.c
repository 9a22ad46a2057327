/*
 * matgen.c -- synthetic CrsMatrix inputs for tests and bench (host, C + OpenMP).
 *
 * Input definitions follow the reference's own generators (paths relative to
 * /root/reference); nothing here is on the SpMV/SpGEMM product path.
 *
 *  - b200gen_kk_*      : KokkosSparse::Impl::kk_sparseMatrix_generate,
 *                        sparse/src/KokkosSparse_IOUtils.hpp:29-81 (libc
 *                        srand(13721)/rand() structure; unsorted rows without
 *                        duplicates; band around the diagonal with wrap).
 *  - b200gen_lap27_*   : Test::generate_structured_matrix3D("FE", ...),
 *                        test_common/KokkosKernels_Test_Structured_Matrix.hpp:
 *                        1906-1977 (interior), 1979-2050 (faces), 3364-3449:
 *                        27-point trilinear-FE Laplacian with Neumann
 *                        boundaries, optionally `ndof` unknowns per node.
 *  - b200gen_uniform_* : exactly `deg` distinct uniform-random columns per row
 *                        (BASELINE.json config 4, SpGEMM A*A).
 *  - b200gen_rmat_*    : Graph500 R-MAT (config 3).
 *  - b200gen_fill_*    : counter-based uniform values (splitmix64), replacing
 *                        Kokkos::Random_XorShift64_Pool streams which are not
 *                        reproducible without Kokkos.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GEN_API __attribute__((visibility("default")))

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
static inline double u01(uint64_t seed, uint64_t i) {
  return (double)(splitmix64(seed * 0xD1342543DE82EF95ull + i) >> 11) * (1.0 / 9007199254740992.0);
}

GEN_API void b200gen_fill_f64(int64_t n, double* v, double lo, double hi, uint64_t seed) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) v[i] = lo + (hi - lo) * u01(seed, (uint64_t)i);
}
GEN_API void b200gen_fill_f32(int64_t n, float* v, float lo, float hi, uint64_t seed) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) v[i] = (float)(lo + (hi - lo) * u01(seed, (uint64_t)i));
}

/* ---- kk_sparseMatrix_generate clone (IOUtils.hpp:29-81) ----------------- */
/* Pass 1: row pointer; returns nnz.  rowptr has nrows+1 entries. */
GEN_API int64_t b200gen_kk_rowptr(int nrows, int ncols, int64_t nnz_target, int row_size_variance,
                                  int* rowptr) {
  int elements_per_row = nrows ? (int)(nnz_target / nrows) : 0;
  srand(13721);
  rowptr[0] = 0;
  for (int row = 0; row < nrows; row++) {
    int varianz = (1.0 * rand() / RAND_MAX - 0.5) * row_size_variance;
    int numRowEntries = elements_per_row + varianz;
    if (numRowEntries < 0) numRowEntries = 0;
    if (numRowEntries > 0.66 * ncols) numRowEntries = 0.66 * ncols;
    rowptr[row + 1] = rowptr[row] + numRowEntries;
  }
  return rowptr[nrows];
}
/* Pass 2: must directly follow pass 1 with the same arguments -- it re-seeds
 * and replays the row loop so that the rand() stream continues exactly as in
 * the reference's single function. */
GEN_API void b200gen_kk_colidx(int nrows, int ncols, int64_t nnz_target, int row_size_variance,
                               int bandwidth, const int* rowptr, int* colind) {
  srand(13721);
  for (int row = 0; row < nrows; row++) (void)rand();
  (void)nnz_target; (void)row_size_variance;
  for (int row = 0; row < nrows; row++) {
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      while (1) {
        int pos = (1.0 * rand() / RAND_MAX - 0.5) * bandwidth + row;
        while (pos < 0) pos += ncols;
        while (pos >= ncols) pos -= ncols;
        int dup = 0;
        for (int j = rowptr[row]; j < k; j++)
          if (colind[j] == pos) { dup = 1; break; }
        if (!dup) { colind[k] = pos; break; }
      }
    }
  }
}

/* ---- 27-point FE Laplacian, ndof unknowns per node ---------------------- */
/* Element-assembly closed form of the reference tables: for node offset
 * (dx,dy,dz) in {-1,0,1}^3 with z nonzero components the coupling is
 * w(z) * prod_{axes with offset 0} (2 if the node is interior on that axis
 * else 1), w = {4, 0, -1, -1}: interior centre 32, faces 0, edges -2, corners
 * -1 (Structured_Matrix.hpp:1949-1976); x==0 Neumann face centre 16 etc.
 * (:2030-2049).  Columns ascend within a row, as the reference writes them.
 * With ndof > 1 every node coupling becomes an ndof x ndof block
 * B[p][q] = (p==q ? 1 : 0.25); noise > 0 adds noise*u01 per entry so that no
 * stored value is an exact zero (parity tests want every entry to matter). */
static inline int lap27_row_nodes(int nx, int ny, int nz, int ix, int iy, int iz) {
  int cx = (ix > 0) + 1 + (ix < nx - 1);
  int cy = (iy > 0) + 1 + (iy < ny - 1);
  int cz = (iz > 0) + 1 + (iz < nz - 1);
  return cx * cy * cz;
}

GEN_API int64_t b200gen_lap27_nnz(int nx, int ny, int nz, int ndof) {
  int64_t cx = 3LL * nx - 2, cy = 3LL * ny - 2, cz = 3LL * nz - 2; /* sum over i of neighbours on the axis */
  if (nx == 1) cx = 1; if (ny == 1) cy = 1; if (nz == 1) cz = 1;
  return cx * cy * cz * (int64_t)ndof * ndof;
}

/* Rows [row_begin,row_end) of the matrix (row = node*ndof + dof) are written
 * with offsets rebased so that rowptr_out[0] = 0: a row-block shard for the
 * multi-GPU case.  Column indices stay global.  Returns shard nnz.  If
 * colidx == NULL only rowptr_out is produced (sizing pass). */
GEN_API int64_t b200gen_lap27_rows(int nx, int ny, int nz, int ndof, int64_t row_begin,
                                   int64_t row_end, int* rowptr_out, int* colidx, double* vals,
                                   double noise, uint64_t seed) {
  const int64_t nrows = row_end - row_begin;
  /* pass 1: row lengths */
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nrows; ++r) {
    int64_t node = (row_begin + r) / ndof;
    int ix = (int)(node % nx), iy = (int)((node / nx) % ny), iz = (int)(node / ((int64_t)nx * ny));
    rowptr_out[r + 1] = lap27_row_nodes(nx, ny, nz, ix, iy, iz) * ndof;
  }
  rowptr_out[0] = 0;
  int64_t acc = 0;
  for (int64_t r = 0; r < nrows; ++r) { acc += rowptr_out[r + 1]; rowptr_out[r + 1] = (int)acc; }
  if (!colidx) return acc;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < nrows; ++r) {
    const int64_t grow = row_begin + r;
    const int64_t node = grow / ndof;
    const int p = (int)(grow % ndof);
    int ix = (int)(node % nx), iy = (int)((node / nx) % ny), iz = (int)(node / ((int64_t)nx * ny));
    int64_t o = rowptr_out[r];
    for (int dz = -1; dz <= 1; ++dz) {
      if (iz + dz < 0 || iz + dz >= nz) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        if (iy + dy < 0 || iy + dy >= ny) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          if (ix + dx < 0 || ix + dx >= nx) continue;
          const int z = (dx != 0) + (dy != 0) + (dz != 0);
          double w = (z == 0) ? 4.0 : (z == 1 ? 0.0 : -1.0);
          if (dx == 0 && ix > 0 && ix < nx - 1) w *= 2.0;
          if (dy == 0 && iy > 0 && iy < ny - 1) w *= 2.0;
          if (dz == 0 && iz > 0 && iz < nz - 1) w *= 2.0;
          const int64_t nb = node + dx + (int64_t)dy * nx + (int64_t)dz * nx * ny;
          for (int q = 0; q < ndof; ++q) {
            colidx[o] = (int)(nb * ndof + q);
            if (vals) {
              double v = w * (p == q ? 1.0 : 0.25);
              if (noise != 0.0) v += noise * u01(seed, (uint64_t)(grow * 64 + (o - rowptr_out[r])));
              vals[o] = v;
            }
            ++o;
          }
        }
      }
    }
  }
  return acc;
}

/* ---- fixed-degree uniform random rows (sorted, distinct) ---------------- */
static int cmp_int(const void* a, const void* b) {
  int x = *(const int*)a, y = *(const int*)b;
  return (x > y) - (x < y);
}
GEN_API void b200gen_uniform(int nrows, int ncols, int deg, uint64_t seed, int* rowptr, int* colidx) {
  for (int r = 0; r <= nrows; ++r) rowptr[r] = (int)((int64_t)r * deg);
#pragma omp parallel for schedule(static)
  for (int r = 0; r < nrows; ++r) {
    int* c = colidx + (int64_t)r * deg;
    uint64_t ctr = 0;
    for (int k = 0; k < deg; ++k) {
      while (1) {
        int pos = (int)(u01(seed, ((uint64_t)r << 20) + ctr++) * ncols);
        if (pos >= ncols) pos = ncols - 1;
        int dup = 0;
        for (int j = 0; j < k; ++j) if (c[j] == pos) { dup = 1; break; }
        if (!dup) { c[k] = pos; break; }
      }
    }
    qsort(c, (size_t)deg, sizeof(int), cmp_int);
  }
}

/* ---- R-MAT (Graph500 a,b,c,d), duplicates merged, rows sorted ----------- */
/* Two calls: b200gen_rmat_build returns an opaque sorted unique key array and
 * its length; b200gen_rmat_emit writes CSR and frees it. */
typedef struct { uint64_t* keys; int64_t n; int scale; } rmat_t;

static uint64_t* radix_sort_u64(uint64_t* a, uint64_t* tmp, int64_t n, int bits) {
  /* LSD radix sort, 11 bits per pass; returns whichever buffer holds the result */
  for (int shift = 0; shift < bits; shift += 11) {
    int64_t cnt[2049];
    memset(cnt, 0, sizeof(cnt));
    for (int64_t i = 0; i < n; ++i) cnt[((a[i] >> shift) & 2047) + 1]++;
    for (int i = 0; i < 2048; ++i) cnt[i + 1] += cnt[i];
    for (int64_t i = 0; i < n; ++i) tmp[cnt[(a[i] >> shift) & 2047]++] = a[i];
    uint64_t* t = a; a = tmp; tmp = t;
  }
  return a;
}

GEN_API void* b200gen_rmat_build(int scale, int edge_factor, double a, double b, double c,
                                 uint64_t seed, int64_t* nnz_out) {
  const int64_t nedges = (int64_t)edge_factor << scale;
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nedges);
  uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)nedges);
  const double ab = a + b, abc = a + b + c;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nedges; ++e) {
    uint64_t row = 0, col = 0;
    for (int l = 0; l < scale; ++l) {
      double u = u01(seed, (uint64_t)e * 64 + (uint64_t)l);
      int rb = u >= ab, cb = (u >= a && u < ab) || (u >= abc);
      row = (row << 1) | (uint64_t)rb;
      col = (col << 1) | (uint64_t)cb;
    }
    keys[e] = (row << 32) | col;
  }
  uint64_t* sorted = radix_sort_u64(keys, tmp, nedges, 32 + scale);
  uint64_t* other = (sorted == keys) ? tmp : keys;
  int64_t n = 0;
  for (int64_t i = 0; i < nedges; ++i)
    if (i == 0 || sorted[i] != sorted[i - 1]) other[n++] = sorted[i];
  free(sorted);
  rmat_t* r = (rmat_t*)malloc(sizeof(rmat_t));
  r->keys = other; r->n = n; r->scale = scale;
  *nnz_out = n;
  return r;
}

GEN_API void b200gen_rmat_emit(void* h, int* rowptr, int* colidx) {
  rmat_t* r = (rmat_t*)h;
  const int64_t nrows = (int64_t)1 << r->scale;
  memset(rowptr, 0, sizeof(int) * (size_t)(nrows + 1));
  for (int64_t i = 0; i < r->n; ++i) {
    rowptr[(r->keys[i] >> 32) + 1]++;
    colidx[i] = (int)(r->keys[i] & 0xFFFFFFFFull);
  }
  for (int64_t i = 0; i < nrows; ++i) rowptr[i + 1] += rowptr[i];
  free(r->keys);
  free(r);
}
