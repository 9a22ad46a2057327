// gmres.cu -- restarted GMRES around the library's SpMV (callers of spmv in a loop).
//
// Follows KokkosSparse::Experimental::gmres = GmresWrap::gmres, sparse/impl/KokkosSparse_gmres_impl.hpp:58-327 (public entry
// sparse/src/KokkosSparse_gmres.hpp:60-160, options and results sparse/src/KokkosSparse_gmres_handle.hpp:76-110,175), for
// real scalars on a CrsMatrix or a BsrMatrix, with the optional right preconditioner in the form the reference ships and tests: MatrixPrec,
// an spmv with a given matrix (sparse/src/KokkosSparse_MatrixPrec.hpp:79-83).
//
// The algorithm is the reference's, step for step (Arnoldi with CGS2 or MGS, Givens rotations and the triangular solve on
// the host, true residual at convergence of the shortcut residual and at every restart).  What changes is the traffic
// between host and device: the reference reads a scalar or a column of H back after every BLAS call (j+3 synchronisations
// per MGS step); here one Arnoldi step is a single stream-ordered sequence -- SpMV, the dots and updates of the
// orthogonalisation (the coefficients stay in device memory between them), the norm -- followed by ONE copy of the new
// column to the host.  Dots are two-stage with a fixed summation order (up to 8 basis vectors per pass over w), so a
// solve is bit-reproducible run to run.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

// (the b200sp_spmv_* / b200sp_bsr_spmv_* prototypes come with b200sparse.h through common.cuh)

namespace b200sp {
namespace {

constexpr int kGmThreads = 256;
constexpr int kGmDots = 8;  // basis vectors per pass of the multi-dot kernel

// y = A x for the two matrix types the reference's gmres accepts (sparse/impl/KokkosSparse_gmres_spec.hpp:76-82): a CrsMatrix
// (bs == 0, plan = b200sp_spmv_plan) or a BsrMatrix (bs >= 1, plan = b200sp_bsr_plan; rows / nnz count blocks)
template <typename S>
struct LinOp {
  void* plan;
  int rows;
  int64_t nnz;
  int bs;
  const int* rp;
  const int* ci;
  const S* v;
};
inline int apply_op(const LinOp<double>& a, cudaStream_t st, const double* x, double* y) {
  if (a.bs == 0) return b200sp_spmv_f64_i32((b200sp_spmv_plan*)a.plan, st, 'N', a.rows, a.rows, a.nnz, 1.0, a.rp, a.ci, a.v, x, 0.0, y);
  return b200sp_bsr_spmv_f64_i32((b200sp_bsr_plan*)a.plan, st, 'N', a.rows, a.rows, a.nnz, a.bs, 1.0, a.rp, a.ci, a.v, x, 0.0, y);
}
inline int apply_op(const LinOp<float>& a, cudaStream_t st, const float* x, float* y) {
  if (a.bs == 0) return b200sp_spmv_f32_i32((b200sp_spmv_plan*)a.plan, st, 'N', a.rows, a.rows, a.nnz, 1.0f, a.rp, a.ci, a.v, x, 0.0f, y);
  return b200sp_bsr_spmv_f32_i32((b200sp_bsr_plan*)a.plan, st, 'N', a.rows, a.rows, a.nnz, a.bs, 1.0f, a.rp, a.ci, a.v, x, 0.0f, y);
}

// out[i] = V(:, i) . w for i < cnt (cnt <= kGmDots); V column-major with leading dimension ldv.  Two-stage: per-block
// partials in `slots` (grid x kGmDots), summed in block order by the block that takes the last ticket.
template <typename S>
__global__ void __launch_bounds__(kGmThreads) gm_multidot_kernel(int n, int cnt, const S* __restrict__ V, int64_t ldv,
                                                                 const S* __restrict__ w, S* __restrict__ out, S* __restrict__ slots,
                                                                 unsigned* __restrict__ ticket) {
  S part[kGmDots];
#pragma unroll
  for (int i = 0; i < kGmDots; ++i) part[i] = S(0);
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const S wq = w[q];
#pragma unroll
    for (int i = 0; i < kGmDots; ++i)
      if (i < cnt) part[i] += V[q + i * ldv] * wq;
  }
  __shared__ S warp_part[kGmThreads / 32][kGmDots];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kGmDots; ++i) {
    S t = part[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) warp_part[warp][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < kGmDots) {
    S t = S(0);
    for (int wq = 0; wq < kGmThreads / 32; ++wq) t += warp_part[wq][threadIdx.x];
    slots[(size_t)blockIdx.x * kGmDots + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < cnt) {
    S t = S(0);
    for (unsigned b = 0; b < gridDim.x; ++b) t += ((volatile S*)slots)[(size_t)b * kGmDots + threadIdx.x];
    out[threadIdx.x] = t;
  }
  if (threadIdx.x == 0) *ticket = 0;
}

// w += sum_i (-h[i]) * V(:, i), columns in order (gemv "N" with alpha = -1 as a sequence of axpys)
template <typename S>
__global__ void __launch_bounds__(kGmThreads) gm_update_kernel(int n, int cnt, const S* __restrict__ V, int64_t ldv, const S* __restrict__ h,
                                                               S* __restrict__ w) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    S acc = w[q];
    for (int i = 0; i < cnt; ++i) acc += (S(-1) * h[i]) * V[q + i * ldv];
    w[q] = acc;
  }
}

// y = a*x + b*y  (b == 0: y = a*x without reading y)
template <typename S>
__global__ void __launch_bounds__(kGmThreads) gm_axpby_kernel(int n, S a, const S* __restrict__ x, S b, S* __restrict__ y) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    y[q] = (b == S(0)) ? a * x[q] : a * x[q] + b * y[q];
}

// out = base + sum_i c[i] * V(:, i)   (base may be null: 0)
template <typename S>
__global__ void __launch_bounds__(kGmThreads) gm_lincomb_kernel(int n, int cnt, const S* __restrict__ V, int64_t ldv, const S* __restrict__ c,
                                                                const S* __restrict__ base, S* __restrict__ out) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    S acc = base ? base[q] : S(0);
    for (int i = 0; i < cnt; ++i) acc += c[i] * V[q + i * ldv];
    out[q] = acc;
  }
}

template <typename S>
struct Gm {
  cudaStream_t st;
  int n, grid;
  S* slots;
  unsigned* ticket;
  S* hbuf;  // device scratch for dot results / coefficients

  int dots(int cnt, const S* V, int64_t ldv, const S* w, S* out) {  // out: device, cnt entries
    for (int i0 = 0; i0 < cnt; i0 += kGmDots) {
      const int c = std::min(kGmDots, cnt - i0);
      gm_multidot_kernel<S><<<grid, kGmThreads, 0, st>>>(n, c, V + (int64_t)i0 * ldv, ldv, w, out + i0, slots, ticket);
      B200SP_LAUNCH_CHECK();
    }
    return B200SP_OK;
  }
  int update(int cnt, const S* V, int64_t ldv, const S* h, S* w) {
    gm_update_kernel<S><<<grid, kGmThreads, 0, st>>>(n, cnt, V, ldv, h, w);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
  }
  int axpby(S a, const S* x, S b, S* y) {
    gm_axpby_kernel<S><<<grid, kGmThreads, 0, st>>>(n, a, x, b, y);
    B200SP_LAUNCH_CHECK();
    return B200SP_OK;
  }
  // host value of sqrt(w.w) (synchronises)
  int nrm2(const S* w, S* out_host) {
    int rc = dots(1, w, n, w, hbuf);
    if (rc) return rc;
    S d;
    B200SP_CUDA_TRY(cudaMemcpyAsync(&d, hbuf, sizeof(S), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA_TRY(cudaStreamSynchronize(st));
    *out_host = std::sqrt(d);
    return B200SP_OK;
  }
};

#define GM_TRY(expr)                  \
  do {                                \
    const int _rc = (expr);           \
    if (_rc != B200SP_OK) return _rc; \
  } while (0)

template <typename S>
int gmres_impl(const LinOp<S>& A, const LinOp<S>* M, cudaStream_t st, const S* B, S* X, int m, S tol, int max_restart, int ortho,
               int* num_iters_out, S* end_rel_res_out, int* conv_flag_out) {
  B200SP_REQUIRE(A.plan != nullptr, "gmres: null plan for A");
  B200SP_REQUIRE(A.rows >= 0 && A.nnz >= 0 && A.bs >= 0 && m >= 1 && max_restart >= 0, "gmres: bad size (rows=%d m=%d max_restart=%d)", A.rows, m,
                 max_restart);
  B200SP_REQUIRE((int64_t)A.rows * std::max(A.bs, 1) <= INT32_MAX, "gmres: point dimension exceeds int32");
  const int n = A.rows * std::max(A.bs, 1);
  const int* rp = A.rp;
  B200SP_REQUIRE(ortho == 0 || ortho == 1, "Invalid argument for 'ortho'.  Please use 'CGS2' or 'MGS'.");  // gmres_impl.hpp:173
  B200SP_REQUIRE(num_iters_out && end_rel_res_out && conv_flag_out, "gmres: null result pointer");
  B200SP_REQUIRE(n == 0 || (rp && B && X), "gmres: null array");
  const bool prec = M != nullptr;
  B200SP_REQUIRE(!prec || (M->plan != nullptr && M->rows * std::max(M->bs, 1) == n), "gmres: the preconditioner matrix needs its own plan and A's size");
  *num_iters_out = 0;
  *end_rel_res_out = S(0);
  *conv_flag_out = 0;
  if (n == 0) return B200SP_OK;

  DevTmp tmp(st);
  S *Xiter = nullptr, *Res = nullptr, *Wj = nullptr, *Wj2 = nullptr, *V = nullptr, *hdev = nullptr, *slots = nullptr;
  unsigned* ticket = nullptr;
  const int grid = std::max(1, std::min((n + kGmThreads - 1) / kGmThreads, sm_count() * 4));
  B200SP_CUDA_TRY(tmp.alloc(&Xiter, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&Res, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&Wj, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&Wj2, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&V, (size_t)n * (size_t)(m + 1)));
  B200SP_CUDA_TRY(tmp.alloc(&hdev, (size_t)(2 * (m + 1) + 2)));  // [0, m]: column of H, [m+1, 2m+1]: second CGS pass, [2m+2]: norm
  B200SP_CUDA_TRY(tmp.alloc(&slots, (size_t)grid * kGmDots));
  B200SP_CUDA_TRY(tmp.alloc(&ticket, 1));
  B200SP_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  Gm<S> g{st, n, grid, slots, ticket, hdev + 2 * (m + 1) + 1};
  const int64_t ld = n;
  const int ldh = m + 1;
  std::vector<S> H((size_t)ldh * m, S(0)), GVec((size_t)m + 1, S(0)), Ls((size_t)m, S(0)), Cos((size_t)m, S(0)), Sin((size_t)m, S(0)),
      col((size_t)(2 * (m + 1) + 2), S(0));

  bool converged = false;
  int cycle = 0, numIters = 0;
  S nrmB, trueRes, relRes, shortRelRes;
  GM_TRY(g.nrm2(B, &nrmB));
  B200SP_CUDA_TRY(cudaMemcpyAsync(Res, B, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  GM_TRY(apply_op(A, st, X, Wj));  // wj = A x
  GM_TRY(g.axpby(S(-1), Wj, S(1), Res));                   // res = b - A x
  GM_TRY(g.nrm2(Res, &trueRes));
  if (nrmB != S(0)) {
    relRes = trueRes / nrmB;
  } else if (trueRes == S(0)) {
    relRes = trueRes;
  } else {  // B is zero, but X has a wrong initial guess (:124-127)
    B200SP_CUDA_TRY(cudaMemsetAsync(X, 0, sizeof(S) * (size_t)n, st));
    relRes = S(0);
  }
  shortRelRes = relRes;
  if (relRes < tol) converged = true;
  B200SP_CUDA_TRY(cudaMemcpyAsync(Xiter, X, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st));

  while (!converged && cycle <= max_restart && shortRelRes >= S(1e-14)) {
    GVec[0] = trueRes;
    S* Vj = V;
    GM_TRY(g.axpby(S(1) / trueRes, Res, S(0), Vj));  // V0 = res / |res|
    for (int j = 0; j < m; j++) {
      if (prec) {  // right preconditioner: wj = A (M vj)
        GM_TRY(apply_op(*M, st, Vj, Wj2));
        GM_TRY(apply_op(A, st, Wj2, Wj));
      } else {
        GM_TRY(apply_op(A, st, Vj, Wj));
      }
      S* Hj = H.data() + (size_t)j * ldh;
      if (ortho == 1) {  // MGS: the coefficient of each step feeds the next update on the device
        for (int i = 0; i <= j; i++) {
          GM_TRY(g.dots(1, V + (int64_t)i * ld, ld, Wj, hdev + i));
          GM_TRY(g.update(1, V + (int64_t)i * ld, ld, hdev + i, Wj));
        }
      } else {  // CGS2
        GM_TRY(g.dots(j + 1, V, ld, Wj, hdev));
        GM_TRY(g.update(j + 1, V, ld, hdev, Wj));
        GM_TRY(g.dots(j + 1, V, ld, Wj, hdev + (m + 1)));
        GM_TRY(g.update(j + 1, V, ld, hdev + (m + 1), Wj));
      }
      GM_TRY(g.dots(1, Wj, ld, Wj, hdev + 2 * (m + 1)));
      // the one read-back of this Arnoldi step
      B200SP_CUDA_TRY(cudaMemcpyAsync(col.data(), hdev, sizeof(S) * (size_t)(2 * (m + 1) + 1), cudaMemcpyDeviceToHost, st));
      B200SP_CUDA_TRY(cudaStreamSynchronize(st));
      for (int i = 0; i <= j; i++) Hj[i] = (ortho == 1) ? col[i] : col[i] + col[m + 1 + i];
      const S tmpNrm = std::sqrt(col[2 * (m + 1)]);
      Hj[j + 1] = tmpNrm;
      if (tmpNrm > S(1e-14)) {
        Vj = V + (int64_t)(j + 1) * ld;
        GM_TRY(g.axpby(S(1) / Hj[j + 1], Wj, S(0), Vj));
      }
      // Givens rotations (Demmel et al., Alg. 3) and the shortcut residual -- host, as in the reference (:183-205)
      for (int i = 0; i < j; i++) {
        const S tempVal = Cos[i] * Hj[i] + Sin[i] * Hj[i + 1];
        Hj[i + 1] = -Sin[i] * Hj[i] + Cos[i] * Hj[i + 1];
        Hj[i] = tempVal;
      }
      const S f = Hj[j], gg = Hj[j + 1];
      const S f2 = f * f, g2 = gg * gg;
      S fg2 = f2 + g2;
      const S D1 = S(1) / std::sqrt(f2 * fg2);
      Cos[j] = f2 * D1;
      fg2 = fg2 * D1;
      Hj[j] = f * fg2;
      Sin[j] = f * D1 * gg;
      Hj[j + 1] = S(0);
      GVec[j + 1] = GVec[j] * (-Sin[j]);
      GVec[j] = GVec[j] * Cos[j];
      shortRelRes = std::fabs(GVec[j + 1]) / nrmB;
      if (tmpNrm <= S(1e-14) && shortRelRes >= tol) {
        set_error("GMRES has experienced lucky breakdown, but the residual has not converged. Solver terminated without convergence.");
        return B200SP_ERR_STATE;
      }
      if (shortRelRes != shortRelRes) {
        set_error("gmres: Relative residual is nan. Terminating solver.");
        return B200SP_ERR_STATE;
      }
      if (shortRelRes < tol || j == m - 1) {
        for (int i = 0; i < m; ++i) Ls[i] = GVec[i];
        for (int i = j; i >= 0; --i) {  // upper-triangular solve on the rotated H
          S s = Ls[i];
          for (int q = i + 1; q <= j; ++q) s -= H[i + (size_t)q * ldh] * Ls[q];
          Ls[i] = s / H[i + (size_t)i * ldh];
        }
        B200SP_CUDA_TRY(cudaMemcpyAsync(hdev, Ls.data(), sizeof(S) * (size_t)(j + 1), cudaMemcpyHostToDevice, st));
        if (prec) {  // Xiter = X + M (V y)
          gm_lincomb_kernel<S><<<grid, kGmThreads, 0, st>>>(n, j + 1, V, ld, hdev, (const S*)nullptr, Wj);
          B200SP_LAUNCH_CHECK();
          GM_TRY(apply_op(*M, st, Wj, Wj2));
          B200SP_CUDA_TRY(cudaMemcpyAsync(Xiter, X, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st));
          GM_TRY(g.axpby(S(1), Wj2, S(1), Xiter));
        } else {  // Xiter = X + V y
          gm_lincomb_kernel<S><<<grid, kGmThreads, 0, st>>>(n, j + 1, V, ld, hdev, X, Xiter);
          B200SP_LAUNCH_CHECK();
        }
        GM_TRY(apply_op(A, st, Xiter, Wj));
        B200SP_CUDA_TRY(cudaMemcpyAsync(Res, B, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        GM_TRY(g.axpby(S(-1), Wj, S(1), Res));
        GM_TRY(g.nrm2(Res, &trueRes));
        relRes = trueRes / nrmB;
        numIters = j + 1;
        if (relRes < tol) {
          converged = true;
          B200SP_CUDA_TRY(cudaMemcpyAsync(X, Xiter, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st));
          break;
        } else if (shortRelRes < S(1e-30)) {
          break;
        }
      }
    }
    cycle++;
    B200SP_CUDA_TRY(cudaMemcpyAsync(X, Xiter, sizeof(S) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  }
  B200SP_CUDA_TRY(cudaStreamSynchronize(st));
  *end_rel_res_out = relRes;
  *conv_flag_out = converged ? 0 : (shortRelRes < tol ? 2 : 1);  // Conv, LOA, NoConv (gmres_handle.hpp:84-89)
  *num_iters_out = cycle > 0 ? (cycle - 1) * m + numIters : 0;
  return B200SP_OK;
}

}  // namespace
}  // namespace b200sp

using namespace b200sp;

extern "C" {

#define B200SP_GMRES_ENTRY(NAME, S, PLAN, BSARG, BSVAL)                                                                                \
  int NAME(PLAN* plan_A, void* stream, int rows, int64_t nnz BSARG, const int* row_ptr, const int* col_idx, const S* vals, PLAN* plan_M,  \
           int64_t nnz_M, const int* row_ptr_M, const int* col_idx_M, const S* vals_M, const S* b, S* x, int m, S tol, int max_restart,   \
           int ortho, int* num_iters, S* end_rel_res, int* conv_flag) {                                                                \
    const LinOp<S> A{(void*)plan_A, rows, nnz, BSVAL, row_ptr, col_idx, vals};                                                          \
    const LinOp<S> M{(void*)plan_M, rows, nnz_M, BSVAL, row_ptr_M, col_idx_M, vals_M};                                                  \
    return gmres_impl<S>(A, row_ptr_M ? &M : nullptr, (cudaStream_t)stream, b, x, m, tol, max_restart, ortho, num_iters, end_rel_res,   \
                         conv_flag);                                                                                                    \
  }
#define B200SP_COMMA_BS , int bs
B200SP_GMRES_ENTRY(b200sp_gmres_f64_i32, double, b200sp_spmv_plan, , 0)
B200SP_GMRES_ENTRY(b200sp_gmres_f32_i32, float, b200sp_spmv_plan, , 0)
B200SP_GMRES_ENTRY(b200sp_gmres_bsr_f64_i32, double, b200sp_bsr_plan, B200SP_COMMA_BS, bs)
B200SP_GMRES_ENTRY(b200sp_gmres_bsr_f32_i32, float, b200sp_bsr_plan, B200SP_COMMA_BS, bs)

}  // extern "C"
