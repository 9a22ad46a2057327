// Full specialisations of GAUSS_SEIDEL_SYMBOLIC / GAUSS_SEIDEL_NUMERIC / GAUSS_SEIDEL_APPLY (CRS format) for Kokkos::Cuda that
// forward to libb200sparse.  Generic declarations: sparse/impl/KokkosSparse_gauss_seidel_spec.hpp:105-151; the native bodies
// being replaced (PointGaussSeidel through the handle's algorithm): :153-262.
//
// Needs one member on GaussSeidelHandle (sparse/src/KokkosSparse_gauss_seidel_handle.hpp, next to the colour views):
//   b200sp_gs_plan* b200_gs_plan = nullptr;   // released in the destructor with b200sp_gs_plan_destroy
// and the usual called-flags (set_call_symbolic / set_call_numeric, :150-160).  Only the point algorithms (GS_DEFAULT,
// GS_PERMUTED, GS_TEAM) are taken; the cluster and two-stage handles keep the native path: their handle types differ, the
// front end reaches these structs only with a PointGaussSeidelHandle when get_algorithm_type() says so -- the body checks.
// numeric with a given inverse diagonal keeps the native path too (declared, throws).  x / y with several columns are swept
// column by column (the columns are independent systems).
#ifndef KOKKOSSPARSE_GAUSS_SEIDEL_B200_TPL_SPEC_DECL_HPP_
#define KOKKOSSPARSE_GAUSS_SEIDEL_B200_TPL_SPEC_DECL_HPP_
#ifdef KOKKOSKERNELS_ENABLE_TPL_B200SPARSE

#include "KokkosSparse_b200_utils.hpp"

namespace KokkosSparse {
namespace Impl {

inline int b200_call_gs_numeric(b200sp_gs_plan* p, void* s, int n, const int* rp, const int* ci, const double* v) {
  return b200sp_gs_numeric_f64_i32(p, s, n, rp, ci, v);
}
inline int b200_call_gs_numeric(b200sp_gs_plan* p, void* s, int n, const int* rp, const int* ci, const float* v) {
  return b200sp_gs_numeric_f32_i32(p, s, n, rp, ci, v);
}
inline int b200_call_gs_apply(b200sp_gs_plan* p, void* s, int n, const int* rp, const int* ci, const double* v, double* x, const double* y,
                              int zero, double omega, int sweeps, int dir) {
  return b200sp_gs_apply_f64_i32(p, s, n, rp, ci, v, x, y, zero, omega, sweeps, dir);
}
inline int b200_call_gs_apply(b200sp_gs_plan* p, void* s, int n, const int* rp, const int* ci, const float* v, float* x, const float* y,
                              int zero, float omega, int sweeps, int dir) {
  return b200sp_gs_apply_f32_i32(p, s, n, rp, ci, v, x, y, zero, omega, sweeps, dir);
}

template <class GsHandle>
inline b200sp_gs_plan* b200_gs_plan_of(GsHandle* gsh) {
  if (!gsh->b200_gs_plan) KOKKOSSPARSE_IMPL_B200_SAFE_CALL(b200sp_gs_plan_create(&gsh->b200_gs_plan));
  return gsh->b200_gs_plan;
}

#define KOKKOSSPARSE_B200_GS_DECL(SCALAR, MEMSPACE, ETI_AVAIL)                                                          \
  template <>                                                                                                          \
  struct GAUSS_SEIDEL_SYMBOLIC<Kokkos::Cuda, KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE), \
                               KOKKOSSPARSE_B200_IV(const int, MEMSPACE), true, ETI_AVAIL> {                           \
    using KernelHandle = KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE);                                                       \
    using c_int_view_t = KOKKOSSPARSE_B200_IV(const int, MEMSPACE);                                                    \
    static void gauss_seidel_symbolic(const Kokkos::Cuda& exec, KernelHandle* handle, typename KernelHandle::const_nnz_lno_t num_rows, \
                                      typename KernelHandle::const_nnz_lno_t num_cols, c_int_view_t row_map, c_int_view_t entries, \
                                      bool is_graph_symmetric) {                                                       \
      if (num_rows != num_cols) throw std::runtime_error("KokkosSparse::gauss_seidel_symbolic[TPL_B200]: square matrices only"); \
      auto* gsh = handle->get_point_gs_handle();                                                                       \
      KOKKOSSPARSE_IMPL_B200_SAFE_CALL(b200sp_gs_symbolic_i32(b200_gs_plan_of(gsh), (void*)exec.cuda_stream(), num_rows, row_map.data(), \
                                                              entries.data(), is_graph_symmetric ? 1 : 0));            \
      gsh->set_call_symbolic(true);                                                                                    \
      gsh->set_call_numeric(false);                                                                                    \
    }                                                                                                                  \
  };                                                                                                                   \
  template <>                                                                                                          \
  struct GAUSS_SEIDEL_NUMERIC<Kokkos::Cuda, KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KokkosSparse::SparseMatrixFormat::CRS, \
                              KOKKOSSPARSE_B200_IV(const int, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE),   \
                              KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE), true, ETI_AVAIL> {                         \
    using KernelHandle    = KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE);                                                    \
    using c_int_view_t    = KOKKOSSPARSE_B200_IV(const int, MEMSPACE);                                                 \
    using c_scalar_view_t = KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE);                                              \
    static void gauss_seidel_numeric(const Kokkos::Cuda& exec, KernelHandle* handle, typename KernelHandle::const_nnz_lno_t num_rows, \
                                     typename KernelHandle::const_nnz_lno_t, c_int_view_t row_map, c_int_view_t entries,   \
                                     c_scalar_view_t values, bool) {                                                   \
      auto* gsh = handle->get_point_gs_handle();                                                                       \
      if (!gsh->is_symbolic_called()) throw std::runtime_error("KokkosSparse::gauss_seidel_numeric: call gauss_seidel_symbolic first"); \
      KOKKOSSPARSE_IMPL_B200_SAFE_CALL(b200_call_gs_numeric(b200_gs_plan_of(gsh), (void*)exec.cuda_stream(), num_rows, row_map.data(), \
                                                            entries.data(), values.data()));                           \
      gsh->set_call_numeric(true);                                                                                     \
    }                                                                                                                  \
    static void gauss_seidel_numeric(const Kokkos::Cuda&, KernelHandle*, typename KernelHandle::const_nnz_lno_t,           \
                                     typename KernelHandle::const_nnz_lno_t, c_int_view_t, c_int_view_t, c_scalar_view_t,   \
                                     c_scalar_view_t /*given_inverse_diagonal*/, bool) {                               \
      throw std::runtime_error("KokkosSparse::gauss_seidel_numeric[TPL_B200]: a given inverse diagonal is not supported");  \
    }                                                                                                                  \
  };                                                                                                                   \
  template <>                                                                                                          \
  struct GAUSS_SEIDEL_APPLY<Kokkos::Cuda, KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE), KokkosSparse::SparseMatrixFormat::CRS,  \
                            KOKKOSSPARSE_B200_IV(const int, MEMSPACE), KOKKOSSPARSE_B200_IV(const int, MEMSPACE),     \
                            KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE), KOKKOSSPARSE_B200_MV(SCALAR, MEMSPACE),     \
                            KOKKOSSPARSE_B200_MV(const SCALAR, MEMSPACE), true, ETI_AVAIL> {                           \
    using KernelHandle    = KOKKOSSPARSE_B200_KH(SCALAR, MEMSPACE);                                                    \
    using c_int_view_t    = KOKKOSSPARSE_B200_IV(const int, MEMSPACE);                                                 \
    using c_scalar_view_t = KOKKOSSPARSE_B200_IV(const SCALAR, MEMSPACE);                                              \
    using x_view_t        = KOKKOSSPARSE_B200_MV(SCALAR, MEMSPACE);                                                    \
    using y_view_t        = KOKKOSSPARSE_B200_MV(const SCALAR, MEMSPACE);                                              \
    static void gauss_seidel_apply(const Kokkos::Cuda& exec, KernelHandle* handle, typename KernelHandle::const_nnz_lno_t num_rows, \
                                   typename KernelHandle::const_nnz_lno_t, c_int_view_t row_map, c_int_view_t entries,     \
                                   c_scalar_view_t values, x_view_t x_lhs_output_vec, y_view_t y_rhs_input_vec,        \
                                   bool init_zero_x_vector, bool /*update_y_vector*/, typename KernelHandle::nnz_scalar_t omega, \
                                   int numIter, bool apply_forward, bool apply_backward) {                            \
      auto* gsh = handle->get_point_gs_handle();                                                                       \
      if (!gsh->is_numeric_called()) throw std::runtime_error("KokkosSparse::gauss_seidel_apply: call gauss_seidel_numeric first"); \
      if (!apply_forward && !apply_backward) return;                                                                   \
      const int dir = (apply_forward && apply_backward) ? 0 : (apply_forward ? 1 : 2);                                 \
      for (size_t col = 0; col < x_lhs_output_vec.extent(1); ++col) /* LayoutLeft: a column is contiguous */           \
        KOKKOSSPARSE_IMPL_B200_SAFE_CALL(b200_call_gs_apply(                                                           \
            b200_gs_plan_of(gsh), (void*)exec.cuda_stream(), num_rows, row_map.data(), entries.data(), values.data(),  \
            x_lhs_output_vec.data() + col * x_lhs_output_vec.stride(1), y_rhs_input_vec.data() + col * y_rhs_input_vec.stride(1), \
            init_zero_x_vector ? 1 : 0, omega, numIter, dir));                                                         \
    }                                                                                                                  \
  };

#define KOKKOSSPARSE_B200_GS_DECL_S(SCALAR, ETI_AVAIL)              \
  KOKKOSSPARSE_B200_GS_DECL(SCALAR, Kokkos::CudaSpace, ETI_AVAIL)   \
  KOKKOSSPARSE_B200_GS_DECL(SCALAR, Kokkos::CudaUVMSpace, ETI_AVAIL)

KOKKOSSPARSE_B200_GS_DECL_S(double, true)
KOKKOSSPARSE_B200_GS_DECL_S(float, true)
KOKKOSSPARSE_B200_GS_DECL_S(double, false)
KOKKOSSPARSE_B200_GS_DECL_S(float, false)

}  // namespace Impl
}  // namespace KokkosSparse
#endif
#endif
