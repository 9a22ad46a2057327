// crs_io.cpp -- host-side readers / writers of the matrix files the reference's drivers take
// (SURVEY.md section 8f rank 2): MatrixMarket (.mtx / .mm) and the raw binary CRS dump (.bin).
// No CUDA here; the arrays land in host memory owned by the library (b200sp_host_free) and the
// caller moves them to the device.
//
// Follows, step for step,
//   read_mtx                 sparse/src/KokkosSparse_IOUtils.hpp:784-996
//   read_graph_bin           sparse/src/KokkosSparse_IOUtils.hpp:680-695
//   write_matrix_mtx         sparse/src/KokkosSparse_IOUtils.hpp:631-653
//   write_graph_bin          sparse/src/KokkosSparse_IOUtils.hpp:487-500
//   read_kokkos_crst_matrix  sparse/src/KokkosSparse_IOUtils.hpp:1237-1290 (dispatch on the extension;
//                            .bin carries no column count: ncols = max column + 1)
// Harwell-Boeing (.hb / .rsa) is not implemented (B200SP_ERR_INVALID_ARGUMENT).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/b200sparse.h"

namespace b200sp {
void set_error(const char* fmt, ...);
}
using b200sp::set_error;

namespace {

enum MtxFormat { UNDEFINED_FORMAT, COORDINATE, ARRAY };
enum MtxField { UNDEFINED_FIELD, REAL, COMPLEX, INTEGER, PATTERN };
enum MtxSym { UNDEFINED_SYMMETRY, GENERAL, SYMMETRIC, SKEW_SYMMETRIC, HERMITIAN };

template <typename S>
struct Edge {
  int src, dst;
  S ew;
  bool operator<(const Edge& a) const { return (src < a.src) || (src == a.src && dst < a.dst); }
};

bool endswith(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

template <typename T>
T* host_alloc(size_t n) {
  return static_cast<T*>(malloc(sizeof(T) * std::max<size_t>(n, 1)));
}

template <typename S>
int read_mtx_impl(const char* path, bool symmetrize, bool remove_diagonal, bool transpose, int* nrows, int* ncols,
                  int64_t* ne, int** xadj, int** adj, S** ew) {
  std::ifstream mmf(path, std::ifstream::in);
  if (!mmf.is_open()) {
    set_error("File cannot be opened: %s", path);
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  std::string fline;
  getline(mmf, fline);
  if (fline.size() < 2 || fline[0] != '%' || fline[1] != '%') {
    set_error("Invalid MM file. Line-1");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  bool is_matrix = false;
  MtxFormat fmt = UNDEFINED_FORMAT;
  MtxField field = UNDEFINED_FIELD;
  MtxSym sym = UNDEFINED_SYMMETRY;
  if (fline.find("matrix") != std::string::npos) {
    is_matrix = true;
  } else if (fline.find("vector") != std::string::npos) {
    set_error("MatrixMarket \"vector\" is not supported by read_mtx()");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  if (fline.find("coordinate") != std::string::npos) fmt = COORDINATE;
  else if (fline.find("array") != std::string::npos) fmt = ARRAY;
  if (fline.find("real") != std::string::npos || fline.find("double") != std::string::npos) {
    field = REAL;
  } else if (fline.find("complex") != std::string::npos) {
    set_error("scalar_t in read_mtx() incompatible with complex-typed MatrixMarket file.");
    return B200SP_ERR_INVALID_ARGUMENT;
  } else if (fline.find("integer") != std::string::npos) {
    field = INTEGER;
  } else if (fline.find("pattern") != std::string::npos) {
    field = PATTERN;
  }
  if (fline.find("general") != std::string::npos) sym = GENERAL;
  else if (fline.find("skew-symmetric") != std::string::npos) sym = SKEW_SYMMETRIC;
  else if (fline.find("symmetric") != std::string::npos) sym = SYMMETRIC;  // after skew-symmetric: substring
  else if (fline.find("hermitian") != std::string::npos || fline.find("Hermitian") != std::string::npos) sym = HERMITIAN;
  if (fmt == ARRAY) {
    if (sym == UNDEFINED_SYMMETRY) sym = GENERAL;
    if (sym != GENERAL) {
      set_error("array format MatrixMarket file must have general symmetry (optional to include \"general\")");
      return B200SP_ERR_INVALID_ARGUMENT;
    }
  }
  if (!is_matrix) {
    set_error("MatrixMarket file header is missing the object type.");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  if (fmt == UNDEFINED_FORMAT) {
    set_error("MatrixMarket file header is missing the format.");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  if (field == UNDEFINED_FIELD) {
    set_error("MatrixMarket file header is missing the field type.");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  if (sym == UNDEFINED_SYMMETRY) {
    set_error("MatrixMarket file header is missing the symmetry type.");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  while (true) {
    if (!getline(mmf, fline)) {
      set_error("MatrixMarket file ends before the size line");
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    if (fline.empty() || fline[0] != '%') break;
  }
  std::stringstream ss(fline);
  long long nr = 0, nc = 0, nnz = 0;
  ss >> nr >> nc;
  if (fmt == COORDINATE) ss >> nnz;
  else nnz = nr * nc;
  if (nr < 0 || nc < 0 || nnz < 0 || nr > INT32_MAX || nc > INT32_MAX) {
    set_error("MatrixMarket size line out of range: %lld %lld %lld", nr, nc, nnz);
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  long long numEdges = nnz;
  symmetrize = symmetrize || sym != GENERAL;
  if (symmetrize && nr != nc) {
    set_error("A non-square matrix cannot be symmetrized.");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  if (fmt == ARRAY) {
    if (symmetrize) {
      set_error("array format MatrixMarket file cannot be symmetrized.");
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    if (field == PATTERN) {
      set_error("array format MatrixMarket file can't have \"pattern\" field type.");
      return B200SP_ERR_INVALID_ARGUMENT;
    }
  }
  if (symmetrize) numEdges = 2 * nnz;
  if (numEdges > INT32_MAX) {
    set_error("read_mtx: %lld entries exceed int32 offsets", numEdges);
    return B200SP_ERR_OVERFLOW;
  }
  std::vector<Edge<S>> edges((size_t)numEdges);
  size_t nE = 0;
  for (long long i = 0; i < nnz; ++i) {
    if (!getline(mmf, fline)) {
      set_error("MatrixMarket file ends after %lld of %lld entries", i, nnz);
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    std::stringstream ss2(fline);
    long long s, d;
    S w;
    if (fmt == ARRAY) {
      s = i % nr + 1;  // column-major listing
      d = i / nr + 1;
    } else {
      ss2 >> s >> d;
    }
    if (field == PATTERN) {
      w = S(1);
    } else {
      w = S(0);
      ss2 >> w;  // readScalar<scalar_t> (:566-571): parsed in the scalar type itself
    }
    if (s < 1 || d < 1 || s > nr || d > nc) {
      set_error("MatrixMarket entry %lld (%lld, %lld) outside the %lld x %lld matrix", i, s, d, nr, nc);
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    Edge<S> tmp;
    if (!transpose) {
      tmp.src = (int)(s - 1);
      tmp.dst = (int)(d - 1);
    } else {
      tmp.src = (int)(d - 1);
      tmp.dst = (int)(s - 1);
    }
    tmp.ew = w;
    if (tmp.src == tmp.dst) {
      if (!remove_diagonal) edges[nE++] = tmp;
      continue;
    }
    edges[nE++] = tmp;
    if (symmetrize) {
      Edge<S> tmp2;
      tmp2.src = tmp.dst;
      tmp2.dst = tmp.src;
      tmp2.ew = (sym == SKEW_SYMMETRIC) ? -tmp.ew : tmp.ew;  // symmetryFlip (:606-629), real scalars
      edges[nE++] = tmp2;
    }
  }
  mmf.close();
  std::sort(edges.begin(), edges.begin() + nE);
  if (transpose) std::swap(nr, nc);
  int* xa = host_alloc<int>((size_t)nr + 1);
  int* ad = host_alloc<int>(nE);
  S* wv = host_alloc<S>(nE);
  if (!xa || !ad || !wv) {
    free(xa); free(ad); free(wv);
    set_error("read_mtx: out of host memory");
    return B200SP_ERR_ALLOC;
  }
  size_t eind = 0;
  int actual = 0;
  for (int i = 0; i < (int)nr; ++i) {
    xa[i] = actual;
    bool is_first = true;
    while (eind < nE && edges[eind].src == i) {
      if (is_first || !symmetrize || eind == 0 || (eind > 0 && edges[eind - 1].dst != edges[eind].dst)) {
        ad[actual] = edges[eind].dst;
        wv[actual] = edges[eind].ew;
        ++actual;
      }
      is_first = false;
      ++eind;
    }
  }
  xa[nr] = actual;
  *nrows = (int)nr;
  *ncols = (int)nc;
  *ne = actual;
  *xadj = xa;
  *adj = ad;
  *ew = wv;
  return B200SP_OK;
}

template <typename S>
int read_bin_impl(const char* path, int* nrows, int* ncols, int64_t* ne, int** xadj, int** adj, S** ew) {
  std::ifstream f(path, std::ios::in | std::ios::binary);
  if (!f.is_open()) {
    set_error("File cannot be opened: %s", path);
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  int nv = 0, nnz = 0;  // lno_t and size_type are both int32 here (default_types.hpp:41-58)
  f.read((char*)&nv, sizeof(int));
  f.read((char*)&nnz, sizeof(int));
  if (!f || nv < 0 || nnz < 0) {
    set_error("read_graph_bin: bad header in %s", path);
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  int* xa = host_alloc<int>((size_t)nv + 1);
  int* ad = host_alloc<int>((size_t)nnz);
  S* wv = host_alloc<S>((size_t)nnz);
  if (!xa || !ad || !wv) {
    free(xa); free(ad); free(wv);
    set_error("read_graph_bin: out of host memory");
    return B200SP_ERR_ALLOC;
  }
  f.read((char*)xa, sizeof(int) * ((size_t)nv + 1));
  f.read((char*)ad, sizeof(int) * (size_t)nnz);
  f.read((char*)wv, sizeof(S) * (size_t)nnz);
  if (!f) {
    free(xa); free(ad); free(wv);
    set_error("read_graph_bin: %s is shorter than its header says", path);
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  int nc = 0;
  for (int i = 0; i < nnz; ++i) nc = std::max(nc, ad[i]);
  *nrows = nv;
  *ncols = nnz > 0 ? nc + 1 : 0;  // kk_view_reduce_max + 1 (:1279-1282)
  *ne = nnz;
  *xadj = xa;
  *adj = ad;
  *ew = wv;
  return B200SP_OK;
}

template <typename S>
int read_crs_impl(const char* path, int* m, int* n, int64_t* nnz, int** rp, int** ci, S** v) {
  if (!path || !m || !n || !nnz || !rp || !ci || !v) {
    set_error("read_kokkos_crst_matrix: null argument");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  const std::string s(path);
  if (endswith(s, ".mtx") || endswith(s, ".mm")) return read_mtx_impl<S>(path, false, false, false, m, n, nnz, rp, ci, v);
  if (endswith(s, ".rsa") || endswith(s, ".hb")) {
    set_error("read_kokkos_crst_matrix: Harwell-Boeing files are not supported by this library");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  if (endswith(s, ".bin")) return read_bin_impl<S>(path, m, n, nnz, rp, ci, v);
  set_error("read_matrix: File extension on %s does not correspond to an known format", path);
  return B200SP_ERR_INVALID_ARGUMENT;
}

template <typename S>
int write_crs_impl(const char* path, int m, int n, int64_t nnz, const int* rp, const int* ci, const S* v) {
  if (!path || m < 0 || n < 0 || nnz < 0 || (m > 0 && !rp) || (nnz > 0 && (!ci || !v))) {
    set_error("write_kokkos_crst_matrix: bad argument");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  const std::string s(path);
  if (endswith(s, ".mtx") || endswith(s, ".mm")) {
    std::ofstream f(path);
    if (!f.is_open()) {
      set_error("File cannot be opened: %s", path);
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    f << "%%MatrixMarket matrix coordinate real general\n";
    f << m << " " << n << " " << nnz << '\n';
    f << std::setprecision(17) << std::scientific;
    for (int i = 0; i < m; ++i)
      for (int j = rp[i]; j < rp[i + 1]; ++j) f << i + 1 << " " << ci[j] + 1 << " " << v[j] << '\n';
    return f.good() ? B200SP_OK : B200SP_ERR_INVALID_ARGUMENT;
  }
  if (endswith(s, ".bin")) {
    if (m != n) {  // :767-770
      set_error("write_kokkos_crst_matrix only supports square matrices");
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    if (nnz > INT32_MAX) {
      set_error("write_graph_bin: nnz exceeds int32 offsets");
      return B200SP_ERR_OVERFLOW;
    }
    std::ofstream f(path, std::ios::out | std::ios::binary);
    if (!f.is_open()) {
      set_error("File cannot be opened: %s", path);
      return B200SP_ERR_INVALID_ARGUMENT;
    }
    const int nv = m, ne = (int)nnz;
    const int zero = 0;
    f.write((const char*)&nv, sizeof(int));
    f.write((const char*)&ne, sizeof(int));
    if (m > 0) f.write((const char*)rp, sizeof(int) * ((size_t)m + 1));
    else f.write((const char*)&zero, sizeof(int));
    f.write((const char*)ci, sizeof(int) * (size_t)ne);
    f.write((const char*)v, sizeof(S) * (size_t)ne);
    return f.good() ? B200SP_OK : B200SP_ERR_INVALID_ARGUMENT;
  }
  set_error("write_kokkos_crst_matrix: File extension on %s does not correspond to a known format", path);
  return B200SP_ERR_INVALID_ARGUMENT;
}

}  // namespace

extern "C" {

int b200sp_read_crs_f64(const char* path, int* m, int* n, int64_t* nnz, int** row_ptr, int** col_idx, double** vals) {
  return read_crs_impl<double>(path, m, n, nnz, row_ptr, col_idx, vals);
}
int b200sp_read_crs_f32(const char* path, int* m, int* n, int64_t* nnz, int** row_ptr, int** col_idx, float** vals) {
  return read_crs_impl<float>(path, m, n, nnz, row_ptr, col_idx, vals);
}
int b200sp_read_mtx_f64(const char* path, int symmetrize, int remove_diagonal, int transpose, int* m, int* n, int64_t* nnz,
                        int** row_ptr, int** col_idx, double** vals) {
  if (!path || !m || !n || !nnz || !row_ptr || !col_idx || !vals) {
    set_error("read_mtx: null argument");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  return read_mtx_impl<double>(path, symmetrize != 0, remove_diagonal != 0, transpose != 0, m, n, nnz, row_ptr, col_idx, vals);
}
int b200sp_read_mtx_f32(const char* path, int symmetrize, int remove_diagonal, int transpose, int* m, int* n, int64_t* nnz,
                        int** row_ptr, int** col_idx, float** vals) {
  if (!path || !m || !n || !nnz || !row_ptr || !col_idx || !vals) {
    set_error("read_mtx: null argument");
    return B200SP_ERR_INVALID_ARGUMENT;
  }
  return read_mtx_impl<float>(path, symmetrize != 0, remove_diagonal != 0, transpose != 0, m, n, nnz, row_ptr, col_idx, vals);
}
int b200sp_write_crs_f64(const char* path, int m, int n, int64_t nnz, const int* row_ptr, const int* col_idx,
                         const double* vals) {
  return write_crs_impl<double>(path, m, n, nnz, row_ptr, col_idx, vals);
}
int b200sp_write_crs_f32(const char* path, int m, int n, int64_t nnz, const int* row_ptr, const int* col_idx,
                         const float* vals) {
  return write_crs_impl<float>(path, m, n, nnz, row_ptr, col_idx, vals);
}
void b200sp_host_free(void* p) { free(p); }

}  // extern "C"
