// oracle/_ref/libkkref.so, spadd part: the reference's own numeric functors of KokkosSparse::spadd_numeric
// (SortedNumericSumFunctor / UnsortedNumericSumFunctor, sparse/impl/KokkosSparse_spadd_numeric_impl.hpp:27-171), compiled from
// the reference tree in place (path injected by oracle/Makefile as KKREF_SPADD_NUMERIC) over the stand-ins in
// oracle/kokkos_mock/spadd, and run row by row as a serial RangePolicy would.  No reference source is copied into this repository.
// TEST INFRASTRUCTURE ONLY: validates oracle/kk_oracle_crs.c's restatement bit for bit (tests/test_oracle_crs.py).
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include KKREF_SPADD_NUMERIC

namespace {
template <class T>
struct V1 {  // rank-1 view over caller memory
  using value_type           = T;
  using non_const_value_type = typename std::remove_const<T>::type;
  T* p;
  T& operator()(int64_t i) const { return p[i]; }
};
}  // namespace

extern "C" {
__attribute__((visibility("default"))) void kkref_spadd_sorted_numeric_f64(int m, const int* rpA, const int* ciA, const double* vA,
                                                                            double alpha, const int* rpB, const int* ciB,
                                                                            const double* vB, double beta, const int* rpC, int* ciC,
                                                                            double* vC) {
  using F = KokkosSparse::Impl::SortedNumericSumFunctor<int, int, V1<const int>, V1<const int>, V1<const int>, V1<const int>,
                                                        V1<const int>, V1<int>, V1<const double>, V1<const double>, V1<double>, double,
                                                        double>;
  F f(V1<const int>{rpA}, V1<const int>{rpB}, V1<const int>{rpC}, V1<const int>{ciA}, V1<const int>{ciB}, V1<int>{ciC}, V1<const double>{vA},
      V1<const double>{vB}, V1<double>{vC}, alpha, beta);
  for (int i = 0; i < m; ++i) f(i);
}
// a_pos / b_pos: where each entry of A / B goes inside its row of C (spadd_symbolic's by-products for unsorted input)
__attribute__((visibility("default"))) void kkref_spadd_unsorted_numeric_f64(int m, const int* rpA, const int* ciA, const double* vA,
                                                                              double alpha, const int* rpB, const int* ciB,
                                                                              const double* vB, double beta, const int* rpC, int* ciC,
                                                                              double* vC, const int* a_pos, const int* b_pos) {
  using F = KokkosSparse::Impl::UnsortedNumericSumFunctor<int, int, V1<const int>, V1<const int>, V1<const int>, V1<const int>,
                                                          V1<const int>, V1<int>, V1<const double>, V1<const double>, V1<double>, double,
                                                          double>;
  F f(V1<const int>{rpA}, V1<const int>{rpB}, V1<const int>{rpC}, V1<const int>{ciA}, V1<const int>{ciB}, V1<int>{ciC}, V1<const double>{vA},
      V1<const double>{vB}, V1<double>{vC}, alpha, beta, V1<int>{const_cast<int*>(a_pos)}, V1<int>{const_cast<int*>(b_pos)});
  for (int i = 0; i < m; ++i) f(i);
}
}
