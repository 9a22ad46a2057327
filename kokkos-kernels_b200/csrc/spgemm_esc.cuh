// spgemm_esc.cuh -- register-resident expand / sort / compress kernels of the SpGEMM (included by spgemm.cu).
//
// One CTA of T threads builds one row of C = A*B whose product count f = sum_j nnz(B_{a_ij}) is at most T*I:
// product p (its ORDINAL in the reference's accumulation order: A's row in storage order, each B row in storage
// order, sparse/impl/KokkosSparse_spgemm_impl_seq.hpp:147-164) belongs to thread p % T, item p / T, so the T*I
// gathers of a row are issued back to back from registers (I independent col/val loads per thread in flight)
// and consecutive lanes read consecutive entries of one B row.  Nothing is gathered twice.
//
//  symbolic  esc_sym_kernel : the columns go through a shared-memory hash set (atomicCAS), nnz(C_i) = inserts.
//  numeric   esc_num_kernel : counting sort by a MONOTONE bucket map of the column
//      bucket(c) = c - cmin                       (row span <= NB: injective)
//                = ((c - cmin) * floor(NB*2^32 / span)) >> 32   otherwise
//    histogram (shared atomicAdd, the returned count is the product's arrival rank in its bucket) -> exclusive
//    scan -> scatter of the keys -> every product ranks itself among the few members of its bucket by
//    (column, ordinal) -> the row is sorted, equal columns adjacent in ORDINAL order.  Rows without duplicate
//    columns (f == nnz(C_i), known from symbolic) leave at once with coalesced stores; otherwise the first
//    product of each run adds its run up in ordinal order -- the oracle's order, so the values equal the
//    reference's SPGEMM_DEBUG result bit for bit when it is compiled without FMA contraction -- and an
//    exclusive scan of the run heads gives the output positions.  No floating-point atomics, no second walk,
//    deterministic.  Rows come out sorted by column: the reference's separate sort_crs_matrix pass
//    (sparse/impl/KokkosSparse_spgemm_numeric_spec.hpp:138-140) is not needed.
//
// A row qualifies when max(f, 2 * nnz(A_i)) <= T*I (the staged A row shares memory with the sorted output).
#pragma once
#include "common.cuh"
#include <limits.h>

namespace b200sp {

template <typename S>
__device__ __forceinline__ S mul_rn(S a, S b);
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }

constexpr int esc_log2(int v) { return v <= 1 ? 0 : 1 + esc_log2(v >> 1); }

// exclusive scan in place of a[0..n), n <= T*K, thread t owns a[t*K .. t*K+K); every thread gets the total.
// wsum: 34 ints of shared memory.  Ends with a barrier (a[] and the total are visible to all).
template <int T, int K>
__device__ __forceinline__ int esc_block_scan(int* a, int n, int* wsum) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = T / 32;
  int v[K];
  int s = 0;
  const int base = tid * K;
  if (K % 4 == 0 && base + K <= n) {
#pragma unroll
    for (int c = 0; c < K / 4; ++c) {
      const int4 q = reinterpret_cast<const int4*>(a + base)[c];
      v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
    }
#pragma unroll
    for (int i = 0; i < K; ++i) s += v[i];
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      v[i] = (base + i < n) ? a[base + i] : 0;
      s += v[i];
    }
  }
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  int woff = 0, total;
  if (NW > 1) {
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      const int x = lane < NW ? wsum[lane] : 0;
      int xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) xi += t;
      }
      if (lane < NW) wsum[lane] = xi - x;
      if (lane == 31) wsum[32] = xi;
    }
    __syncthreads();
    woff = wsum[w];
    total = wsum[32];
  } else {
    total = __shfl_sync(0xffffffffu, inc, 31);
  }
  int run = woff + inc - s;
  if (K % 4 == 0 && base + K <= n) {
#pragma unroll
    for (int c = 0; c < K / 4; ++c) {
      int4 q;
      q.x = run; run += v[4 * c];
      q.y = run; run += v[4 * c + 1];
      q.z = run; run += v[4 * c + 2];
      q.w = run; run += v[4 * c + 3];
      reinterpret_cast<int4*>(a + base)[c] = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (base + i < n) a[base + i] = run;
      run += v[i];
    }
  }
  __syncthreads();
  return total;
}

// exclusive scan of one int per thread over the CTA (registers only + wsum); total to all threads.
template <int T>
__device__ __forceinline__ int esc_block_scan1(int v, int* wsum, int& total) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int NW = T / 32;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  int woff = 0;
  if (NW > 1) {
    __syncthreads();  // wsum may still be read from an earlier scan
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      const int x = lane < NW ? wsum[lane] : 0;
      int xi = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, xi, o);
        if (lane >= o) xi += t;
      }
      if (lane < NW) wsum[lane] = xi - x;
      if (lane == 31) wsum[32] = xi;
    }
    __syncthreads();
    woff = wsum[w];
    total = wsum[32];
  } else {
    total = __shfl_sync(0xffffffffu, inc, 31);
  }
  return woff + inc - v;
}

// ---- staging of A's row and the product -> (entry of A, offset in its B row) map -------------------
// bs[j] = start of the B row of A's j-th entry, pre[j] = ordinal of its first product (pre[nA] = f).
// Returns the common length of the B rows (>= 1) when all of them have the same length, else 0.
template <int T, int I>
__device__ __forceinline__ int esc_stage_row(int a0, int nA, int f, const int* __restrict__ ciA,
                                             const int* __restrict__ rpB, int* bs, int* pre, int* wsum, int* sflag) {
  const int tid = threadIdx.x;
  if (tid == 0) *sflag = 0;
  __syncthreads();
  int first_len = -1;
  for (int j = tid; j < nA; j += T) {
    const int c = ldg(ciA + a0 + j);
    const int b0 = ldg(rpB + c);
    const int len = ldg(rpB + c + 1) - b0;
    bs[j] = b0;
    pre[j] = len;
  }
  __syncthreads();
  first_len = pre[0];
  for (int j = tid; j < nA; j += T)
    if (pre[j] != first_len) *sflag = 1;  // benign race: every writer stores 1
  __syncthreads();
  const bool uniform = (*sflag == 0) && first_len > 0;
  // nA <= T*I/2: each thread scans I/2 (>= 1) consecutive lengths
  constexpr int K = (I / 2 > 0) ? I / 2 : 1;
  if (tid == 0) pre[nA] = f;  // outside the scanned range; published by the scan's closing barrier
  esc_block_scan<T, K>(pre, nA, wsum);
  return uniform ? first_len : 0;
}

// ---------------------------------------------------------------------------------------------------
// SYMBOLIC
// ---------------------------------------------------------------------------------------------------
template <int T, int I>
struct EscSymLayout {
  static constexpr int CAP = T * I;
  static constexpr int SLOTS = 2 * CAP;  // a power of two (T and I are)
  static constexpr int LOG2SLOTS = esc_log2(SLOTS);
  static constexpr int NA = CAP / 2;
  // keys[SLOTS] | bs[NA] | pre[NA + 4] | wsum[36]
  static constexpr size_t BYTES = sizeof(int) * (size_t)(SLOTS + NA + NA + 4 + 36);
};

template <int T, int I, int MINB>
__global__ void __launch_bounds__(T, MINB)
    esc_sym_kernel(int nrows_bin, const int* __restrict__ rows, const int* __restrict__ rpA, const int* __restrict__ ciA,
                   const int* __restrict__ rpB, const int* __restrict__ ciB, const int* __restrict__ flops,
                   int* __restrict__ row_nnz) {
  using L = EscSymLayout<T, I>;
  constexpr int SLOTS = L::SLOTS;
  constexpr int EMPTYK = -1;
  extern __shared__ __align__(16) int esc_sm[];
  int* keys = esc_sm;
  int* bs = keys + SLOTS;
  int* pre = bs + L::NA;
  int* wsum = pre + L::NA + 4;
  const int tid = threadIdx.x;
  const int i = rows[blockIdx.x];
  const int f = flops[i];
  if (f == 0) {
    if (tid == 0) row_nnz[i] = 0;
    return;
  }
  const int a0 = rpA[i], nA = rpA[i + 1] - a0;
  {
    int4* k4 = reinterpret_cast<int4*>(keys);
    for (int s = tid; s < SLOTS / 4; s += T) k4[s] = make_int4(EMPTYK, EMPTYK, EMPTYK, EMPTYK);
  }
  const int L0 = esc_stage_row<T, I>(a0, nA, f, ciA, rpB, bs, pre, wsum, wsum + 35);
  int col[I];
  if (L0 > 0) {
    int j = tid / L0, t = tid - j * L0;
    const int dj = T / L0, dt = T - dj * L0;
#pragma unroll
    for (int k = 0; k < I; ++k) {
      col[k] = (k * T + tid < f) ? ld_stream(ciB + bs[j] + t) : EMPTYK;
      j += dj;
      t += dt;
      if (t >= L0) {
        t -= L0;
        ++j;
      }
    }
  } else {
    int lo = 0;
#pragma unroll
    for (int k = 0; k < I; ++k) {
      const int p = k * T + tid;
      col[k] = EMPTYK;
      if (p < f) {
        int hi = nA;  // largest j with pre[j] <= p (pre[nA] = f > p)
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (pre[mid] <= p) lo = mid; else hi = mid;
        }
        col[k] = ld_stream(ciB + bs[lo] + (p - pre[lo]));
      }
    }
  }
  int mine = 0;
#pragma unroll
  for (int k = 0; k < I; ++k) {
    const int c = col[k];
    if (c != EMPTYK) {
      unsigned h = ((unsigned)c * 0x9E3779B1u) >> (32 - L::LOG2SLOTS);
      while (true) {
        const int kcur = ((volatile int*)keys)[h];
        if (kcur == c) break;
        if (kcur == EMPTYK) {
          const int old = atomicCAS(&keys[h], EMPTYK, c);
          if (old == EMPTYK) {
            ++mine;
            break;
          }
          if (old == c) break;
        }
        h = (h + 1) & (unsigned)(SLOTS - 1);
      }
    }
  }
  int total;
  esc_block_scan1<T>(mine, wsum, total);
  if (tid == 0) row_nnz[i] = total;
}

// ---------------------------------------------------------------------------------------------------
// NUMERIC
// ---------------------------------------------------------------------------------------------------
template <typename S, int T, int I, int LOG2NB>
struct EscNumLayout {
  static constexpr int CAP = T * I;
  static constexpr int NB = 1 << LOG2NB;
  static constexpr int NA = CAP / 2;
  // off[NB + 4] | skey[CAP] | wsum[36] | sord[CAP] (u16) | union { bs[NA], pre[NA + 4], va[NA] ; skey2[CAP], sval[CAP] }
  static constexpr size_t OFF_BYTES = sizeof(int) * (size_t)(NB + 4);
  static constexpr size_t KEY_BYTES = sizeof(int) * (size_t)CAP;
  static constexpr size_t WS_BYTES = sizeof(int) * 36;
  static constexpr size_t ORD_BYTES = ((sizeof(unsigned short) * (size_t)CAP) + 15) & ~(size_t)15;
  static constexpr size_t STAGE_BYTES = sizeof(int) * (size_t)(NA + NA + 4) + sizeof(S) * (size_t)NA;
  static constexpr size_t SORT_BYTES = (sizeof(int) + sizeof(S)) * (size_t)CAP;
  static constexpr size_t UNION_BYTES = ((STAGE_BYTES > SORT_BYTES ? STAGE_BYTES : SORT_BYTES) + 15) & ~(size_t)15;
  static constexpr size_t BYTES = OFF_BYTES + KEY_BYTES + WS_BYTES + ORD_BYTES + UNION_BYTES;
};

template <typename S, int T, int I, int LOG2NB, int MINB>
__global__ void __launch_bounds__(T, MINB)
    esc_num_kernel(int nrows_bin, const int* __restrict__ rows, const int* __restrict__ rpA, const int* __restrict__ ciA,
                   const S* __restrict__ vA, const int* __restrict__ rpB, const int* __restrict__ ciB,
                   const S* __restrict__ vB, const int* __restrict__ rpC, int* __restrict__ ciC, S* __restrict__ vC,
                   const int* __restrict__ cmin_arr, const int* __restrict__ cmax_arr, const int* __restrict__ flops) {
  using L = EscNumLayout<S, T, I, LOG2NB>;
  constexpr int CAP = L::CAP;
  constexpr int NB = L::NB;
  constexpr int NOKEY = INT_MAX;
  extern __shared__ __align__(16) unsigned char esc_raw[];
  int* off = reinterpret_cast<int*>(esc_raw);
  int* skey = reinterpret_cast<int*>(esc_raw + L::OFF_BYTES);
  int* wsum = reinterpret_cast<int*>(esc_raw + L::OFF_BYTES + L::KEY_BYTES);
  unsigned short* sord = reinterpret_cast<unsigned short*>(esc_raw + L::OFF_BYTES + L::KEY_BYTES + L::WS_BYTES);
  unsigned char* un = esc_raw + L::OFF_BYTES + L::KEY_BYTES + L::WS_BYTES + L::ORD_BYTES;
  // staging view (S first: 8-byte alignment)
  S* va = reinterpret_cast<S*>(un);
  int* bs = reinterpret_cast<int*>(un + sizeof(S) * (size_t)L::NA);
  int* pre = bs + L::NA;
  // sorted view
  S* sval = reinterpret_cast<S*>(un);
  int* skey2 = reinterpret_cast<int*>(un + sizeof(S) * (size_t)CAP);

  const int tid = threadIdx.x;
  const int i = rows[blockIdx.x];
  const int cbase = rpC[i];
  const int nz = rpC[i + 1] - cbase;
  if (nz == 0) return;
  const int f = flops[i];
  const int a0 = rpA[i], nA = rpA[i + 1] - a0;
  const int cmin = cmin_arr[i];
  const long long span = (long long)cmax_arr[i] - cmin + 1;
  const bool dense = span <= NB;
  const unsigned long long mult = dense ? 0ull : (((unsigned long long)NB << 32) / (unsigned long long)span);
  {
    int4* o4 = reinterpret_cast<int4*>(off);
    for (int s = tid; s < (NB + 4) / 4; s += T) o4[s] = make_int4(0, 0, 0, 0);
  }
  for (int j = tid; j < nA; j += T) va[j] = ldg(vA + a0 + j);
  const int L0 = esc_stage_row<T, I>(a0, nA, f, ciA, rpB, bs, pre, wsum, wsum + 35);

  // ---- expand: I products per thread, in registers
  int col[I];
  S val[I];
  if (L0 > 0) {
    int j = tid / L0, t = tid - j * L0;
    const int dj = T / L0, dt = T - dj * L0;
#pragma unroll
    for (int k = 0; k < I; ++k) {
      col[k] = NOKEY;
      val[k] = S(0);
      if (k * T + tid < f) {
        const int jb = bs[j] + t;
        col[k] = ld_stream(ciB + jb);
        val[k] = mul_rn(ld_stream(vB + jb), va[j]);  // b_val * a_val (impl_seq.hpp:163)
      }
      j += dj;
      t += dt;
      if (t >= L0) {
        t -= L0;
        ++j;
      }
    }
  } else {
    int lo = 0;
#pragma unroll
    for (int k = 0; k < I; ++k) {
      const int p = k * T + tid;
      col[k] = NOKEY;
      val[k] = S(0);
      if (p < f) {
        int hi = nA;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (pre[mid] <= p) lo = mid; else hi = mid;
        }
        const int jb = bs[lo] + (p - pre[lo]);
        col[k] = ld_stream(ciB + jb);
        val[k] = mul_rn(ld_stream(vB + jb), va[lo]);
      }
    }
  }
  // ---- histogram over the monotone buckets; the old count is the arrival rank
  int br[I];
#pragma unroll
  for (int k = 0; k < I; ++k) {
    br[k] = 0;
    if (col[k] != NOKEY) {
      const unsigned d = (unsigned)(col[k] - cmin);
      const int b = dense ? (int)d : (int)(((unsigned long long)d * mult) >> 32);
      const int r = atomicAdd(&off[b], 1);
      br[k] = b | (r << LOG2NB);
    }
  }
  __syncthreads();  // staging (bs / pre / va) is dead from here on
  esc_block_scan<T, NB / T>(off, NB, wsum);
  if (tid == 0) off[NB] = f;
  // ---- scatter keys + ordinals to bucket order
#pragma unroll
  for (int k = 0; k < I; ++k)
    if (col[k] != NOKEY) {
      const int b = br[k] & (NB - 1);
      const int pos0 = off[b] + (br[k] >> LOG2NB);
      skey[pos0] = col[k];
      sord[pos0] = (unsigned short)(k * T + tid);
    }
  __syncthreads();
  // ---- rank inside the bucket by (column, ordinal) -> sorted position
#pragma unroll
  for (int k = 0; k < I; ++k)
    if (col[k] != NOKEY) {
      const int b = br[k] & (NB - 1);
      const int lo = off[b], hi = off[b + 1];
      int less = 0;
      if (hi - lo > 1) {
        const int c = col[k];
        const int p = k * T + tid;
        for (int m = lo; m < hi; ++m) {
          const int km = skey[m];
          less += (km < c || (km == c && (int)sord[m] < p)) ? 1 : 0;
        }
      }
      skey2[lo + less] = col[k];
      sval[lo + less] = val[k];
    }
  __syncthreads();
  if (nz == f) {  // no duplicate column: the sorted products are the row
    for (int q = tid; q < f; q += T) {
      ciC[cbase + q] = skey2[q];
      vC[cbase + q] = sval[q];
    }
    return;
  }
  // ---- compress: run heads, output positions by an exclusive scan of the head flags
  int heads = 0;
  const int q0 = tid * I;
#pragma unroll
  for (int k = 0; k < I; ++k) {
    const int q = q0 + k;
    if (q < f && (q == 0 || skey2[q] != skey2[q - 1])) ++heads;
  }
  int total;
  int outp = esc_block_scan1<T>(heads, wsum, total);
  // positions of the heads, parked in skey (free since the ranking) so that the stores below are in output order
#pragma unroll
  for (int k = 0; k < I; ++k) {
    const int q = q0 + k;
    if (q < f) {
      const bool head = (q == 0 || skey2[q] != skey2[q - 1]);
      skey[q] = head ? outp : -1;
      if (head) ++outp;
    }
  }
  __syncthreads();
  for (int q = tid; q < f; q += T) {
    const int o = skey[q];
    if (o >= 0 && o < nz) {
      const int c = skey2[q];
      S v = sval[q];
      for (int q2 = q + 1; q2 < f && skey2[q2] == c; ++q2) v += sval[q2];  // ordinal order = the oracle's order
      ciC[cbase + o] = c;
      vC[cbase + o] = v;
    }
  }
}

}  // namespace b200sp
