// spmm_items.h -- per-matrix state of the item-based rank-2 SpMV (spmm.cu), owned by the SpMV plan (spmv.cu).
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace b200sp {

// A work item is a run of <= LMAX consecutive entries of one row: (row, first entry, length, partial slot or -1).  Rows of up
// to LMAX entries are one item (it writes Y directly); longer rows are cut into pieces whose sums go to `partial` and are
// added up in piece order by a second kernel.  Items are sorted by length, longest first, so that the row groups of a warp
// run loops of equal length and the long items start first.
struct MMItems {
  int4* items = nullptr;
  int n_items = 0;
  int4* multi = nullptr;  // rows of several pieces: (row, first partial slot, pieces, 0)
  int n_multi = 0;
  int n_multi_long = 0;   // of them, at the END of `multi`: rows of more than 64 pieces (added up by a CTA each)
  int n_partial = 0;
  void* partial = nullptr;
  size_t partial_bytes = 0;
  int lmax = 0;
  // cache key
  const int* key_row_ptr = nullptr;
  int key_m = -1, key_lmax = -1;
  int64_t key_nnz = -1;
};

}  // namespace b200sp
