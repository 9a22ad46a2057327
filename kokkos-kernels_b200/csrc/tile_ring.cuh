// tile_ring.cuh -- the shared-memory ring of row-aligned CSR tiles that a producer warp fills with
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), as used by spmv_tile_kernel (spmv.cu).
// The producer loop here is the one of spmv_tile_kernel, factored out for the rank-2 kernel
// (spmm.cu); the tile descriptors are the plan's (build_tiles_kernel, spmv.cu).
#pragma once
#include "common.cuh"

namespace b200sp {

template <typename S, int CAP, int STAGES>
struct TileRing {
  static constexpr int RCAP = CAP / 2;  // staged row_ptr entries per tile
  alignas(128) S vals[STAGES][CAP];
  alignas(128) int cols[STAGES][CAP];
  alignas(128) int rows[STAGES][RCAP];
  int4 desc[STAGES];
  alignas(8) uint64_t full[STAGES];
  alignas(8) uint64_t empty[STAGES];
};

template <typename S, int CAP, int STAGES>
__device__ __forceinline__ void tile_ring_init(TileRing<S, CAP, STAGES>& sm, int n_consumer_warps) {
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sm.full[s], 1);                    // producer's arrive.expect_tx (+ TMA byte count)
      mbar_init(&sm.empty[s], n_consumer_warps);    // one arrive per consumer warp
    }
    fence_mbar_init();
  }
  __syncthreads();
}

// Runs in ONE warp (all 32 lanes).  Tile `it` of this CTA is tiles[blockIdx.x + it*gridDim.x]:
// descriptor {r0, r1, s, e}; stage layout: entries [s & ~3, e) at vals/cols[0..), row_ptr[r0 & ~3 ..] at
// rows[0..) (at most RCAP entries).  Sources are aligned down to 16 bytes; the last < 4 entries of an
// array are copied with plain loads so that no bulk copy reads past the allocation.
template <typename S, int CAP, int STAGES>
__device__ __forceinline__ void tile_ring_produce(TileRing<S, CAP, STAGES>& sm, int lane, int m, int64_t nnz, int n_tiles,
                                                  const int4* __restrict__ tiles, const int* __restrict__ row_ptr,
                                                  const int* __restrict__ col_idx, const S* __restrict__ vals) {
  constexpr int RCAP = TileRing<S, CAP, STAGES>::RCAP;
  const uint64_t pol = l2_policy_evict_first();
  const int64_t nnz_al = nnz & ~(int64_t)3;  // bulk copies stay below this entry
  const int rp_al_end = (m + 1) & ~3;        // ... and below this row_ptr entry
  int4 mine = make_int4(0, 0, 0, 0);
  for (int it = 0;; ++it) {
    const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
    if (tile >= n_tiles) break;
    if ((it & 31) == 0) {
      const int64_t t = blockIdx.x + (int64_t)(it + lane) * gridDim.x;
      if (t < n_tiles) mine = tiles[t];
    }
    int4 d;
    d.x = __shfl_sync(0xffffffffu, mine.x, it & 31);
    d.y = __shfl_sync(0xffffffffu, mine.y, it & 31);
    d.z = __shfl_sync(0xffffffffu, mine.z, it & 31);
    d.w = __shfl_sync(0xffffffffu, mine.w, it & 31);
    const int stage = it % STAGES;
    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
    mbar_wait(&sm.empty[stage], ph ^ 1u);

    const int r0 = d.x, r1 = d.y, s = d.z, e = d.w;
    S* sv = sm.vals[stage];
    int* sc = sm.cols[stage];
    int* sr = sm.rows[stage];
    const int s_al = s & ~3;
    const int e_up = (e + 3) & ~3;
    const int64_t bulk_end64 = (int64_t)e_up < nnz_al ? (int64_t)e_up : nnz_al;
    const int bulk_end = (int)bulk_end64;
    const int nb = (r1 > r0 && bulk_end > s_al) ? bulk_end - s_al : 0;
    if (r1 > r0 && (int64_t)e > nnz_al) {
      const int t0 = (int)((int64_t)s_al > nnz_al ? (int64_t)s_al : nnz_al);
      for (int i = t0 + lane; i < e; i += 32) {
        sv[i - s_al] = vals[i];
        sc[i - s_al] = col_idx[i];
      }
    }
    const int r0_al = r0 & ~3;
    int nrp = 0;
    if (r1 > r0) {
      const int want_end = min(r1 + 1, r0_al + RCAP);  // exclusive
      const int want_up = (want_end + 3) & ~3;
      const int rbulk_end = min(min(want_up, r0_al + RCAP), rp_al_end);
      nrp = rbulk_end > r0_al ? rbulk_end - r0_al : 0;
      if (want_end > rp_al_end) {
        const int t0 = max(r0_al, rp_al_end);
        for (int i = t0 + lane; i < want_end; i += 32) sr[i - r0_al] = row_ptr[i];
      }
    }
    __syncwarp();
    if (lane == 0) {
      sm.desc[stage] = d;
      mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)(nb * (sizeof(S) + 4) + nrp * 4));
      if (nb > 0) {
        bulk_g2s(sv, vals + s_al, (uint32_t)(nb * sizeof(S)), &sm.full[stage], pol);
        bulk_g2s(sc, col_idx + s_al, (uint32_t)(nb * 4), &sm.full[stage], pol);
      }
      if (nrp > 0) bulk_g2s(sr, row_ptr + r0_al, (uint32_t)(nrp * 4), &sm.full[stage], pol);
    }
    __syncwarp();
  }
}

// what plan_analyse_mm (spmv.cu) hands to the rank-2 tile kernel (spmm.cu)
struct MMTileView {
  const int4* tiles;  // {r0, r1, s, e} per tile
  int n_tiles;
  int LMAX;           // rows longer than this are left to the segment kernel
  const int4* segs;   // {row, e0, e1, flags} (flags bit 0: row has several segments, bit 1: first segment)
  const int* n_seg;   // device counter
  int seg_cap;
};

}  // namespace b200sp
