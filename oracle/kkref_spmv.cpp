// oracle/_ref/libkkref.so, SpMV part: the reference's own host SpMV -- the Serial hand-unrolled loop (the parity oracle
// north_star names; sparse/impl/KokkosSparse_spmv_impl.hpp:233-305) and the generic functor every other host back end runs
// (SPMV_Functor::operator()(row), :110-132, through spmv_beta_no_transpose's RangePolicy launch, :323-333) -- compiled from
// the reference tree in place (path injected by oracle/Makefile as KKREF_SPMV_IMPL) over the stand-ins in
// oracle/kokkos_mock/spmv.  No reference source is copied into this repository.  TEST INFRASTRUCTURE ONLY: validates O1 / O2
// of oracle/kk_oracle.c bit for bit (tests/test_oracle_spmv.py).
#include <cstddef>
#include <cstdint>
#include KKREF_SPMV_IMPL

namespace {
template <class T>
struct Vec {
  static constexpr int rank = 1;
  using value_type           = T;
  using const_value_type     = const T;
  using non_const_value_type = typename std::remove_const<T>::type;
  T* p;
  size_t n = 0;
  T* data() const { return p; }
  T& operator()(int64_t i) const { return p[i]; }
  size_t extent(int) const { return n; }
};
template <class T>
struct Mat {  // rank-2 strided view: element (i, k) at p[i*rs + k*cs]
  static constexpr int rank = 2;
  using value_type           = T;
  using const_value_type     = const T;
  using non_const_value_type = typename std::remove_const<T>::type;
  T* p;
  size_t n0, n1;
  int64_t rs, cs;
  T* data() const { return p; }
  T& operator()(int64_t i, int64_t k) const { return p[i * rs + k * cs]; }
  size_t extent(int d) const { return d == 0 ? n0 : n1; }
};
struct Graph {
  Vec<const int> row_map, entries;
};
template <class S>
struct Crs {
  using non_const_ordinal_type = int;
  using non_const_size_type    = int;
  using non_const_value_type   = S;
  using value_type             = const S;
  using const_ordinal_type     = const int;
  Graph graph;
  Vec<const S> values;
  int nrows;
  int64_t nnz_;
  int numRows() const { return nrows; }
  int64_t nnz() const { return nnz_; }
  KokkosSparse::SparseRowViewConst<Crs> rowConst(int i) const {
    const int b = graph.row_map(i);
    return KokkosSparse::SparseRowViewConst<Crs>{values.p + b, graph.entries.p + b, graph.row_map(i + 1) - b};
  }
};
struct Handle {  // SPMVHandleImpl's scheduling switches (sparse/src/KokkosSparse_spmv_handle.hpp:306-307): both false by default
  bool force_dynamic_schedule = false, force_static_schedule = false;
};
struct NotSerial {  // an execution space that is not Kokkos::Serial: takes the generic (functor) path
  int concurrency() const { return 2; }
};

template <class Exec, class S>
void run(int nrows, const int* rm, const int* ci, const S* v, const S* x, S* y, S alpha, S beta) {
  Crs<S> A{{Vec<const int>{rm}, Vec<const int>{ci}}, Vec<const S>{v}, nrows, nrows > 0 ? (int64_t)rm[nrows] : 0};
  Vec<const S> X{x};
  Vec<S> Y{y};
  Handle h;
  Exec exec;
  // dobeta as the reference's dispatch sets it (sparse/impl/KokkosSparse_spmv_impl.hpp:517-536): 0, 1, -1, or 2
  using namespace KokkosSparse::Impl;
  if (beta == S(0)) spmv_beta_no_transpose<Exec, Handle, Crs<S>, Vec<const S>, Vec<S>, 0, false>(exec, &h, alpha, A, X, beta, Y);
  else if (beta == S(1)) spmv_beta_no_transpose<Exec, Handle, Crs<S>, Vec<const S>, Vec<S>, 1, false>(exec, &h, alpha, A, X, beta, Y);
  else if (beta == S(-1)) spmv_beta_no_transpose<Exec, Handle, Crs<S>, Vec<const S>, Vec<S>, -1, false>(exec, &h, alpha, A, X, beta, Y);
  else spmv_beta_no_transpose<Exec, Handle, Crs<S>, Vec<const S>, Vec<S>, 2, false>(exec, &h, alpha, A, X, beta, Y);
}
// modes T / H (real scalars): spmv_beta_transpose, :383-460 -- y scaled first, then the order-preserving 4-way unrolled scatter
template <class S>
void run_transpose(int nrows, int ncols, const int* rm, const int* ci, const S* v, const S* x, S* y, S alpha, S beta) {
  Crs<S> A{{Vec<const int>{rm}, Vec<const int>{ci}}, Vec<const S>{v}, nrows, nrows > 0 ? (int64_t)rm[nrows] : 0};
  Vec<const S> X{x, (size_t)nrows};
  Vec<S> Y{y, (size_t)ncols};
  Kokkos::Serial exec;
  using namespace KokkosSparse::Impl;
  if (beta == S(0)) spmv_beta_transpose<Kokkos::Serial, Crs<S>, Vec<const S>, Vec<S>, 0, false>(exec, alpha, A, X, beta, Y);
  else if (beta == S(1)) spmv_beta_transpose<Kokkos::Serial, Crs<S>, Vec<const S>, Vec<S>, 1, false>(exec, alpha, A, X, beta, Y);
  else if (beta == S(-1)) spmv_beta_transpose<Kokkos::Serial, Crs<S>, Vec<const S>, Vec<S>, -1, false>(exec, alpha, A, X, beta, Y);
  else spmv_beta_transpose<Kokkos::Serial, Crs<S>, Vec<const S>, Vec<S>, 2, false>(exec, alpha, A, X, beta, Y);
}
// multivector: spmv_alpha_mv<doalpha> picked from alpha as the unification layer does (sparse/impl/KokkosSparse_spmv_spec.hpp:
// SPMV_MV<...>::spmv_mv: alpha == 0 -> 0, 1 -> 1, -1 -> -1, else 2), then the reference's own dispatch on beta and mode
template <class S>
void run_mv(char mode, int nrows, int ncols, int nvec, const int* rm, const int* ci, const S* v, const S* X, int64_t xr, int64_t xc, S* Y,
            int64_t yr, int64_t yc, S alpha, S beta) {
  Crs<S> A{{Vec<const int>{rm}, Vec<const int>{ci}}, Vec<const S>{v}, nrows, nrows > 0 ? (int64_t)rm[nrows] : 0};
  const bool trans = (mode == 'T' || mode == 'H');
  Mat<const S> Xv{X, (size_t)(trans ? nrows : ncols), (size_t)nvec, xr, xc};
  Mat<S> Yv{Y, (size_t)(trans ? ncols : nrows), (size_t)nvec, yr, yc};
  NotSerial exec;
  const char m[2] = {mode, 0};
  using namespace KokkosSparse::Impl;
  if (alpha == S(0)) spmv_alpha_mv<NotSerial, Crs<S>, Mat<const S>, Mat<S>, 0>(exec, m, alpha, A, Xv, beta, Yv);
  else if (alpha == S(1)) spmv_alpha_mv<NotSerial, Crs<S>, Mat<const S>, Mat<S>, 1>(exec, m, alpha, A, Xv, beta, Yv);
  else if (alpha == S(-1)) spmv_alpha_mv<NotSerial, Crs<S>, Mat<const S>, Mat<S>, -1>(exec, m, alpha, A, Xv, beta, Yv);
  else spmv_alpha_mv<NotSerial, Crs<S>, Mat<const S>, Mat<S>, 2>(exec, m, alpha, A, Xv, beta, Yv);
}
}  // namespace

extern "C" {
__attribute__((visibility("default"))) void kkref_spmv_mv_f64(char mode, int nrows, int ncols, int nvec, const int* rm, const int* ci,
                                                               const double* v, const double* X, int64_t xr, int64_t xc, double* Y,
                                                               int64_t yr, int64_t yc, double alpha, double beta) {
  run_mv<double>(mode, nrows, ncols, nvec, rm, ci, v, X, xr, xc, Y, yr, yc, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_spmv_mv_f32(char mode, int nrows, int ncols, int nvec, const int* rm, const int* ci,
                                                               const float* v, const float* X, int64_t xr, int64_t xc, float* Y, int64_t yr,
                                                               int64_t yc, float alpha, float beta) {
  run_mv<float>(mode, nrows, ncols, nvec, rm, ci, v, X, xr, xc, Y, yr, yc, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_spmv_transpose_f64(int nrows, int ncols, const int* rm, const int* ci, const double* v,
                                                                      const double* x, double* y, double alpha, double beta) {
  run_transpose<double>(nrows, ncols, rm, ci, v, x, y, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_spmv_transpose_f32(int nrows, int ncols, const int* rm, const int* ci, const float* v,
                                                                      const float* x, float* y, float alpha, float beta) {
  run_transpose<float>(nrows, ncols, rm, ci, v, x, y, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_spmv_serial_f64(int nrows, const int* rm, const int* ci, const double* v,
                                                                   const double* x, double* y, double alpha, double beta) {
  run<Kokkos::Serial, double>(nrows, rm, ci, v, x, y, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_spmv_serial_f32(int nrows, const int* rm, const int* ci, const float* v, const float* x,
                                                                   float* y, float alpha, float beta) {
  run<Kokkos::Serial, float>(nrows, rm, ci, v, x, y, alpha, beta);
}
__attribute__((visibility("default"))) void kkref_spmv_functor_f64(int nrows, const int* rm, const int* ci, const double* v,
                                                                    const double* x, double* y, double alpha, double beta) {
  run<NotSerial, double>(nrows, rm, ci, v, x, y, alpha, beta);
}
// the same functor path on `threads` OpenMP threads (bench.py's CPU legs): rows are independent, so the result is the serial one
__attribute__((visibility("default"))) void kkref_spmv_functor_omp_f64(int threads, int nrows, const int* rm, const int* ci, const double* v,
                                                                        const double* x, double* y, double alpha, double beta) {
  Kokkos::kkmock_threads() = threads > 1 ? threads : 1;
  run<NotSerial, double>(nrows, rm, ci, v, x, y, alpha, beta);
  Kokkos::kkmock_threads() = 1;
}
__attribute__((visibility("default"))) void kkref_spmv_functor_f32(int nrows, const int* rm, const int* ci, const float* v, const float* x,
                                                                    float* y, float alpha, float beta) {
  run<NotSerial, float>(nrows, rm, ci, v, x, y, alpha, beta);
}
}
