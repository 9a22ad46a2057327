// oracle/_ref/libkkref.so, BsrMatrix part: the reference's own mode-N BsrMatrix SpMV functor
// (sparse/impl/KokkosSparse_spmv_bsrmatrix_impl_v42.hpp:35-121, the native path of every GPU execution space),
// compiled from the reference tree in place (path injected by oracle/Makefile as KKREF_BSR_V42) over the stand-ins
// in oracle/kokkos_mock/bsr.  Its work items are run 0 .. y.size()-1 in order, as a serial RangePolicy would.
// No reference source is copied into this repository.  TEST INFRASTRUCTURE ONLY: validates
// oracle/kk_oracle_bsr.c's restatement (tests/test_oracle_bsr.py).
#include <cstddef>
#include <cstdint>
#include KKREF_BSR_V42

namespace {

template <class T>
struct Strided2D {  // rank-2 view: element (i, j) at base[i*rs + j*cs]
  T* base;
  int64_t n0, n1, rs, cs;
  using non_const_value_type = typename std::remove_const<T>::type;
  size_t extent(int d) const { return (size_t)(d == 0 ? n0 : n1); }
  size_t size() const { return (size_t)(n0 * n1); }
  T& operator()(int64_t i, int64_t j) const { return base[i * rs + j * cs]; }
};

template <class T>
struct Block {  // one bs x bs block, row-major (BsrMatrix::block_layout_type = LayoutRight, BsrMatrix.hpp:364)
  const T* p;
  int bs;
  const T& operator()(int i, int j) const { return p[i * bs + j]; }
};

struct IndexView {
  const int* p;
  int operator()(int64_t i) const { return p[i]; }
};

template <class T>
struct BsrMock {
  using non_const_ordinal_type = int;
  using non_const_size_type    = int;
  using const_block_type       = Block<T>;
  struct {
    IndexView row_map, entries;
  } graph;
  const T* values;
  int bs;
  int blockDim() const { return bs; }
  const_block_type unmanaged_block_const(int64_t j) const { return Block<T>{values + j * bs * bs, bs}; }
};

template <class T>
void run(int mb, int bs, int nvec, const int* row_map, const int* entries, const T* values, const T* X, int64_t xr,
         int64_t xc, T* Y, int64_t yr, int64_t yc, T alpha, T beta) {
  BsrMock<T> a;
  a.graph.row_map = IndexView{row_map};
  a.graph.entries = IndexView{entries};
  a.values        = values;
  a.bs            = bs;
  Strided2D<const T> x{X, 0, nvec, xr, xc};
  Strided2D<T> y{Y, (int64_t)mb * bs, nvec, yr, yc};
  KokkosSparse::Impl::BsrSpmvV42NonTrans<T, BsrMock<T>, Strided2D<const T>, T, Strided2D<T>> op(alpha, a, x, beta, y);
  for (size_t k = 0; k < y.size(); ++k) op(k);
}

}  // namespace

extern "C" {
__attribute__((visibility("default"))) void kkref_bsr_spmv_v42_f64(int mb, int bs, int nvec, const int* row_map,
                                                                    const int* entries, const double* values,
                                                                    const double* X, int64_t xr, int64_t xc, double* Y,
                                                                    int64_t yr, int64_t yc, double alpha, double beta) {
  run<double>(mb, bs, nvec, row_map, entries, values, X, xr, xc, Y, yr, yc, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_bsr_spmv_v42_f32(int mb, int bs, int nvec, const int* row_map,
                                                                    const int* entries, const float* values,
                                                                    const float* X, int64_t xr, int64_t xc, float* Y,
                                                                    int64_t yr, int64_t yc, float alpha, float beta) {
  run<float>(mb, bs, nvec, row_map, entries, values, X, xr, xc, Y, yr, yc, alpha, beta);
}
}
