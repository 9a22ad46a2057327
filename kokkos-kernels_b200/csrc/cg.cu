// cg.cu -- unpreconditioned conjugate gradients on the device: the reference's CG driver (a caller of spmv in a loop,
// SURVEY.md section 8f rank 4) with no host round trip per iteration.
//
// Follows KokkosKernels::Experimental::Example::pcgsolve with use_sgs = false
// (perf_test/sparse/KokkosSparse_pcg.hpp:248-466; driver perf_test/sparse/KokkosSparse_pcg.cpp:69-122):
//   p = x; Ap = A p; r = b - Ap; p = r; old_rdot = r.r
//   while (tolerance < sqrt(old_rdot) && iteration < maximum_iteration)
//     Ap = A p; alpha = old_rdot / p.Ap; x += alpha p; r -= alpha Ap; beta = r.r / old_rdot; p = r + beta p
// The reference reads three scalars back per iteration (KokkosBlas::dot returns to the host, :385,398).  Here alpha, beta,
// the residual and the iteration counter live in device memory; one iteration is the library's SpMV plus three small
// kernels on the same stream, every kernel first looks at a `done` word, and the host polls that word (pinned, async copy)
// every `check_every` iterations -- so the loop is launch-bound on the host side only for tiny matrices.
// Dots are two-stage with a fixed reduction order (per-thread strided partials, shuffle tree, per-block slots summed by the
// last block in slot order): bit-reproducible run to run for a given grid.
#include <algorithm>
#include <cmath>

#include "common.cuh"

struct b200sp_spmv_plan;

namespace b200sp {
namespace {

struct CgState {       // device-resident scalars of the recurrence
  double precond_old;  // r.z of the previous iteration (preconditioned runs)
  double old_rdot;     // r.r of the previous iteration
  double pAp;          // p.Ap
  double beta;         // r.r / old_rdot
  double norm_res;     // sqrt(r.r)
  int iteration;       // completed iterations
  int done;            // 1 once norm_res <= tolerance or the limit is reached
  unsigned ticket;     // last-block election of the two-stage reductions
  unsigned pad;
};

constexpr int kCgThreads = 256;

// sum of v over the block in a fixed order; result valid in thread 0
__device__ __forceinline__ double cg_block_sum(double v) {
  __shared__ double warp_part[kCgThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < kCgThreads / 32; ++w) t += warp_part[w];
  __syncthreads();
  return t;
}

// Publishes this block's partial and elects the last block; returns true (in thread 0 of that block) with the total in
// *total, summed over the slots in slot order.
__device__ __forceinline__ bool cg_grid_sum(double block_total, double* __restrict__ slots, CgState* __restrict__ st, double* total) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    slots[blockIdx.x] = block_total;
    __threadfence();
    const unsigned t = atomicAdd(&st->ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return false;
  __threadfence();
  double s = 0.0;
  for (unsigned b = 0; b < gridDim.x; ++b) s += ((volatile double*)slots)[b];
  st->ticket = 0;
  *total = s;
  return true;
}

// r = b - Ap; p = r; old_rdot = r.r   (setup, after Ap = A x)
__global__ void __launch_bounds__(kCgThreads) cg_init_kernel(int n, const double* __restrict__ b, const double* __restrict__ Ap,
                                                             double* __restrict__ r, double* __restrict__ p, double* __restrict__ slots,
                                                             CgState* __restrict__ st, double tolerance, int maximum_iteration) {
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = 1.0 * b[i] + -1.0 * Ap[i];
    r[i] = ri;
    p[i] = ri;
    part += ri * ri;
  }
  double total;
  if (cg_grid_sum(cg_block_sum(part), slots, st, &total)) {
    st->old_rdot = total;
    st->norm_res = sqrt(total);
    st->iteration = 0;
    st->beta = 0.0;
    st->done = !(tolerance < st->norm_res && 0 < maximum_iteration);
  }
}

// pAp = p.Ap
__global__ void __launch_bounds__(kCgThreads) cg_dot_kernel(int n, const double* __restrict__ p, const double* __restrict__ Ap,
                                                            double* __restrict__ slots, CgState* __restrict__ st) {
  if (((volatile CgState*)st)->done) return;
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) part += p[i] * Ap[i];
  double total;
  if (cg_grid_sum(cg_block_sum(part), slots, st, &total)) st->pAp = total;
}

// x += alpha p; r -= alpha Ap; r_dot = r.r; then the scalar part of the iteration
__global__ void __launch_bounds__(kCgThreads) cg_update_kernel(int n, const double* __restrict__ p, const double* __restrict__ Ap,
                                                               double* __restrict__ x, double* __restrict__ r, double* __restrict__ slots,
                                                               CgState* __restrict__ st, int pcg) {
  if (((volatile CgState*)st)->done) return;
  const double alpha = (pcg ? st->precond_old : st->old_rdot) / st->pAp;  // pcg.hpp:387-391
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] = alpha * p[i] + 1.0 * x[i];
    const double ri = -alpha * Ap[i] + 1.0 * r[i];
    r[i] = ri;
    part += ri * ri;
  }
  double total;
  if (cg_grid_sum(cg_block_sum(part), slots, st, &total)) {
    if (!pcg) st->beta = total / st->old_rdot;  // preconditioned: beta comes from r.z (cg_rz_kernel)
    st->old_rdot = total;
    st->norm_res = sqrt(total);
    st->iteration += 1;
    // `done` is raised after the p update of this iteration (cg_flag_kernel), as the reference's loop does
  }
}

// p = r + beta p   (preconditioned: p = z + beta p; the caller passes z for r)
__global__ void __launch_bounds__(kCgThreads) cg_p_kernel(int n, const double* __restrict__ r, double* __restrict__ p, CgState* __restrict__ st) {
  if (((volatile CgState*)st)->done) return;
  const double beta = st->beta;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = 1.0 * r[i] + beta * p[i];
}

// preconditioned runs: rz = r.z; setup (init != 0): precond_old = rz; in the loop: beta = rz / precond_old, precond_old = rz
// (pcg.hpp:361,428-429,449)
__global__ void __launch_bounds__(kCgThreads) cg_rz_kernel(int n, const double* __restrict__ r, const double* __restrict__ z,
                                                           double* __restrict__ slots, CgState* __restrict__ st, int init) {
  if (((volatile CgState*)st)->done && !init) return;
  double part = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) part += r[i] * z[i];
  double total;
  if (cg_grid_sum(cg_block_sum(part), slots, st, &total)) {
    if (!init) st->beta = total / st->precond_old;
    st->precond_old = total;
  }
}

// The loop condition of the reference (:372), evaluated once per iteration AFTER the p update.  A kernel of its own: the blocks of
// cg_p_kernel run independently, so none of them may raise a flag the others still have to read as 0.
__global__ void cg_flag_kernel(CgState* __restrict__ st, double tolerance, int maximum_iteration) {
  if (st->done) return;
  if (!(tolerance < st->norm_res && st->iteration < maximum_iteration)) st->done = 1;
}

}  // namespace
}  // namespace b200sp

using namespace b200sp;

struct b200sp_gs_plan;
extern "C" int b200sp_spmv_f64_i32(b200sp_spmv_plan* plan, void* stream, char mode, int m, int n, int64_t nnz, double alpha,
                                   const int* row_ptr, const int* col_idx, const double* vals, const double* x, double beta, double* y);
extern "C" int b200sp_gs_apply_f64_i32(b200sp_gs_plan* p, void* stream, int n, const int* row_ptr, const int* col_idx, const double* vals,
                                       double* x, const double* y, int init_zero_x, double omega, int sweeps, int direction);

struct b200sp_gs2_plan;
extern "C" int b200sp_gs2_apply_f64_i32(b200sp_gs2_plan* p, void* stream, int n, int ncols, const int* row_ptr, const int* col_idx,
                                        const double* vals, double* x, int64_t ldx, const double* b, int64_t ldb, int nrhs, int init_zero_x,
                                        double omega, int num_iter, int direction);

// z = M^-1 r by one symmetric sweep from z = 0 with omega = 1: the point (multicolour) method or the two-stage one, whichever
// plan was given -- what symmetric_gauss_seidel_apply does for the GS handle held by pcgsolve's kernel handle
static int cg_precond(b200sp_gs_plan* gs, b200sp_gs2_plan* gs2, void* stream, int n, const int* row_ptr, const int* col_idx, const double* vals,
                      double* z, const double* r) {
  if (gs) return b200sp_gs_apply_f64_i32(gs, stream, n, row_ptr, col_idx, vals, z, r, 1, 1.0, 1, 0);
  return b200sp_gs2_apply_f64_i32(gs2, stream, n, n, row_ptr, col_idx, vals, z, n, r, n, 1, 1, 1.0, 1, 0);
}

// gs != nullptr: symmetric Gauss-Seidel preconditioner, z = SGS(r) with zero initial guess, omega = 1, one sweep
// (pcgsolve's use_sgs = true, perf_test/sparse/KokkosSparse_pcg.hpp:339-358,412-427)
static int cg_solve(b200sp_spmv_plan* plan, b200sp_gs_plan* gs, b200sp_gs2_plan* gs2, void* stream, int n, int64_t nnz, const int* row_ptr, const int* col_idx,
                    const double* vals, const double* b, double* x, int maximum_iteration, double tolerance, int check_every, int* iterations,
                    double* norm_res) {
  B200SP_REQUIRE(plan != nullptr, "cg_solve: null plan (create one with b200sp_spmv_plan_create)");
  B200SP_REQUIRE(n >= 0 && nnz >= 0 && maximum_iteration >= 0, "cg_solve: negative size");
  B200SP_REQUIRE(iterations != nullptr && norm_res != nullptr, "cg_solve: null result pointer");
  B200SP_REQUIRE(n == 0 || (row_ptr && b && x), "cg_solve: null array");
  cudaStream_t st = (cudaStream_t)stream;
  if (check_every <= 0) check_every = 8;
  *iterations = 0;
  *norm_res = 0.0;
  if (n == 0) return B200SP_OK;
  const int pcg = gs != nullptr || gs2 != nullptr;
  DevTmp tmp(st);
  double *p = nullptr, *r = nullptr, *Ap = nullptr, *z = nullptr, *slots = nullptr;
  CgState* state = nullptr;
  const int grid = std::max(1, std::min((n + kCgThreads - 1) / kCgThreads, sm_count() * 4));
  B200SP_CUDA_TRY(tmp.alloc(&p, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&r, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&Ap, (size_t)n));
  if (pcg) B200SP_CUDA_TRY(tmp.alloc(&z, (size_t)n));
  B200SP_CUDA_TRY(tmp.alloc(&slots, (size_t)grid));
  B200SP_CUDA_TRY(tmp.alloc(&state, 1));
  B200SP_CUDA_TRY(cudaMemsetAsync(state, 0, sizeof(CgState), st));
  CgState* host_state = nullptr;
  B200SP_CUDA_TRY(cudaMallocHost((void**)&host_state, sizeof(CgState)));
  struct HostFree {
    CgState* q;
    ~HostFree() { cudaFreeHost(q); }
  } host_free{host_state};

  // Ap = A x (p = x in the reference; x itself is read here), r = b - Ap, p = r
  int rc = b200sp_spmv_f64_i32(plan, stream, 'N', n, n, nnz, 1.0, row_ptr, col_idx, vals, x, 0.0, Ap);
  if (rc != B200SP_OK) return rc;
  cg_init_kernel<<<grid, kCgThreads, 0, st>>>(n, b, Ap, r, p, slots, state, tolerance, maximum_iteration);
  B200SP_LAUNCH_CHECK();
  if (pcg) {  // z = M^-1 r; precond_old_rdot = r.z; p = z
    rc = cg_precond(gs, gs2, stream, n, row_ptr, col_idx, vals, z, r);
    if (rc != B200SP_OK) return rc;
    cg_rz_kernel<<<grid, kCgThreads, 0, st>>>(n, r, z, slots, state, 1);
    B200SP_LAUNCH_CHECK();
    B200SP_CUDA_TRY(cudaMemcpyAsync(p, z, sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, st));
  }

  int issued = 0;
  for (;;) {
    B200SP_CUDA_TRY(cudaMemcpyAsync(host_state, state, sizeof(CgState), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA_TRY(cudaStreamSynchronize(st));
    if (host_state->done || issued >= maximum_iteration) break;
    const int batch = std::min(check_every, maximum_iteration - issued);
    for (int k = 0; k < batch; ++k) {
      // iterations issued past convergence are no-ops except for their SpMV / Gauss-Seidel sweeps (they cannot see the flag):
      // at most check_every - 1 of them
      rc = b200sp_spmv_f64_i32(plan, stream, 'N', n, n, nnz, 1.0, row_ptr, col_idx, vals, p, 0.0, Ap);
      if (rc != B200SP_OK) return rc;
      cg_dot_kernel<<<grid, kCgThreads, 0, st>>>(n, p, Ap, slots, state);
      B200SP_LAUNCH_CHECK();
      cg_update_kernel<<<grid, kCgThreads, 0, st>>>(n, p, Ap, x, r, slots, state, pcg);
      B200SP_LAUNCH_CHECK();
      if (pcg) {
        rc = cg_precond(gs, gs2, stream, n, row_ptr, col_idx, vals, z, r);
        if (rc != B200SP_OK) return rc;
        cg_rz_kernel<<<grid, kCgThreads, 0, st>>>(n, r, z, slots, state, 0);
        B200SP_LAUNCH_CHECK();
      }
      cg_p_kernel<<<grid, kCgThreads, 0, st>>>(n, pcg ? z : r, p, state);
      B200SP_LAUNCH_CHECK();
      cg_flag_kernel<<<1, 1, 0, st>>>(state, tolerance, maximum_iteration);
      B200SP_LAUNCH_CHECK();
    }
    issued += batch;
  }
  *iterations = host_state->iteration;
  *norm_res = host_state->norm_res;
  return B200SP_OK;
}

extern "C" int b200sp_cg_solve_f64_i32(b200sp_spmv_plan* plan, void* stream, int n, int64_t nnz, const int* row_ptr, const int* col_idx,
                                       const double* vals, const double* b, double* x, int maximum_iteration, double tolerance,
                                       int check_every, int* iterations, double* norm_res) {
  return cg_solve(plan, nullptr, nullptr, stream, n, nnz, row_ptr, col_idx, vals, b, x, maximum_iteration, tolerance, check_every, iterations, norm_res);
}

extern "C" int b200sp_pcg_solve_f64_i32(b200sp_spmv_plan* plan, b200sp_gs_plan* gs_plan, void* stream, int n, int64_t nnz,
                                        const int* row_ptr, const int* col_idx, const double* vals, const double* b, double* x,
                                        int maximum_iteration, double tolerance, int check_every, int* iterations, double* norm_res) {
  B200SP_REQUIRE(gs_plan != nullptr, "pcg_solve: null Gauss-Seidel plan (symbolic and numeric must have run on this matrix)");
  return cg_solve(plan, gs_plan, nullptr, stream, n, nnz, row_ptr, col_idx, vals, b, x, maximum_iteration, tolerance, check_every, iterations, norm_res);
}

extern "C" int b200sp_pcg_solve_gs2_f64_i32(b200sp_spmv_plan* plan, b200sp_gs2_plan* gs2_plan, void* stream, int n, int64_t nnz,
                                            const int* row_ptr, const int* col_idx, const double* vals, const double* b, double* x,
                                            int maximum_iteration, double tolerance, int check_every, int* iterations, double* norm_res) {
  B200SP_REQUIRE(gs2_plan != nullptr, "pcg_solve: null two-stage Gauss-Seidel plan (symbolic and numeric must have run on this matrix)");
  return cg_solve(plan, nullptr, gs2_plan, stream, n, nnz, row_ptr, col_idx, vals, b, x, maximum_iteration, tolerance, check_every, iterations,
                  norm_res);
}
