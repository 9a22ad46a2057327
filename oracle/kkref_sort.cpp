// oracle/_ref/libkkref.so, sorting part: the reference's own SerialRadixSort2 (common/src/KokkosKernels_Sorting.hpp:300-369),
// the per-row sort of sort_crs_matrix on host execution spaces (sparse/impl/KokkosSparse_sort_crs_impl.hpp), compiled from the
// reference tree in place (path injected by oracle/Makefile as KKREF_SORTING) over the stand-ins in oracle/kokkos_mock/sort.
// No reference source is copied into this repository.  TEST INFRASTRUCTURE ONLY: validates the restatement in
// oracle/kk_oracle_crs.c bit for bit (tests/test_oracle_crs.py).
#include <cstdint>
#include KKREF_SORTING

extern "C" {
// keys sorted ascending (stable), perm follows; aux arrays of n entries each
__attribute__((visibility("default"))) void kkref_radix_sort2_u32_i32(uint32_t* keys, uint32_t* keys_aux, int32_t* perm,
                                                                     int32_t* perm_aux, int n) {
  KokkosKernels::SerialRadixSort2<int, uint32_t, int32_t>(keys, keys_aux, perm, perm_aux, n);
}
__attribute__((visibility("default"))) void kkref_radix_sort2_u32_f64(uint32_t* keys, uint32_t* keys_aux, double* perm,
                                                                     double* perm_aux, int n) {
  KokkosKernels::SerialRadixSort2<int, uint32_t, double>(keys, keys_aux, perm, perm_aux, n);
}
}
