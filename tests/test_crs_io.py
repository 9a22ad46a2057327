"""Matrix-file readers / writers (kokkos-kernels_b200/csrc/crs_io.cpp) -- host code, no GPU:
the reference's own IO test (sparse/unit_test/Test_Sparse_IOUtils.hpp:131-175: a symmetric and an
asymmetric 6x6 fixture written as general / symmetric / hermitian / skew-symmetric MatrixMarket and read
back with read_kokkos_crst_matrix), round trips through .mtx and .bin, array format, pattern field, the
error cases read_mtx throws on, and a cross-check against scipy.io.mmread."""
import numpy as np
import pytest
import torch

from kokkos_kernels_b200 import B200SparseError, sparse as sp

SYM = np.array([[11, 12, 13, 14, 15, 16], [12, 2, 0, 0, 0, 0], [13, 0, 0, 0, 0, 0], [14, 0, 0, 4, 0, 0],
                [15, 0, 0, 0, 5, 0], [16, 0, 0, 0, 0, 6]], dtype=np.float64)
ASYM = np.array([[1, 0, 0, 9, 0, 0], [0, 2, 0, 0, 0, 0], [0, 0, 0, 0, 0, 8], [0, 0, 0, 4, 0, 0],
                 [0, 7, 0, 0, 5, 0], [0, 0, 0, 0, 0, 6]], dtype=np.float64)


def compress(D):
    rp, ci, v = [0], [], []
    for row in D:
        for j, x in enumerate(row):
            if x != 0:
                ci.append(j)
                v.append(x)
        rp.append(len(ci))
    return np.array(rp, np.int32), np.array(ci, np.int32), np.array(v)


def write_as_mtx(path, D, kind):
    """write_as_mtx of the reference test (:106-129): symmetric kinds list the lower triangle only."""
    src = np.tril(D) if kind != "general" else D
    rp, ci, v = compress(src)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real {kind}\n")
        f.write(f"{len(rp) - 1} {len(rp) - 1} {len(ci)}\n")
        for i in range(len(rp) - 1):
            for j in range(rp[i], rp[i + 1]):
                f.write(f"{i + 1} {ci[j] + 1} {v[j]}\n")


def as_np(A):
    return A.row_map.numpy(), A.entries.numpy(), A.values.numpy()


@pytest.mark.parametrize("kind,fixture", [("general", ASYM), ("symmetric", SYM), ("hermitian", SYM), ("skew-symmetric", SYM)])
def test_reference_io_fixture(tmp_path, kind, fixture):
    p = tmp_path / f"fix_{kind}.mtx"
    write_as_mtx(p, fixture, kind)
    A = sp.read_kokkos_crst_matrix(p)
    rp, ci, v = as_np(A)
    want = fixture.copy()
    if kind == "skew-symmetric":  # A(j,i) = -A(i,j) for the mirrored entries
        want = np.tril(fixture) - np.tril(fixture, -1).T
    erp, eci, ev = compress(want)
    assert A.numRows() == 6 and A.numCols() == 6
    assert np.array_equal(rp, erp) and np.array_equal(ci, eci) and np.array_equal(v, ev)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("ext", [".mtx", ".bin"])
def test_round_trip(tmp_path, dtype, ext):
    rng = np.random.default_rng(1)
    n = 300
    lens = rng.integers(0, 12, n)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    v = rng.standard_normal(len(ci))
    A = sp.CrsMatrix(torch.from_numpy(rp), torch.from_numpy(ci), torch.from_numpy(v).to(dtype), n)
    p = tmp_path / ("rt" + ext)
    sp.write_kokkos_crst_matrix(A, p)
    B = sp.read_kokkos_crst_matrix(p, dtype=dtype)
    assert np.array_equal(B.row_map.numpy(), rp) and np.array_equal(B.entries.numpy(), ci)
    assert np.array_equal(B.values.numpy(), A.values.numpy()), "17 significant digits / raw bytes: exact round trip"
    assert B.numRows() == n
    assert B.numCols() == (n if ext == ".mtx" else int(ci.max()) + 1)


def test_array_format_pattern_and_comments(tmp_path):
    p = tmp_path / "arr.mtx"
    p.write_text("%%MatrixMarket matrix array real general\n% a comment\n%another\n2 3\n1\n2\n3\n4\n5\n6\n")
    A = sp.read_kokkos_crst_matrix(p)
    assert A.numRows() == 2 and A.numCols() == 3
    assert A.row_map.tolist() == [0, 3, 6] and A.entries.tolist() == [0, 1, 2, 0, 1, 2]
    assert A.values.tolist() == [1, 3, 5, 2, 4, 6]          # column-major listing
    q = tmp_path / "pat.mtx"
    q.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n3 3 3\n2 1\n3 3\n3 1\n")
    B = sp.read_kokkos_crst_matrix(q)
    assert B.row_map.tolist() == [0, 2, 3, 5] and B.entries.tolist() == [1, 2, 0, 0, 2] and B.values.tolist() == [1.0] * 5
    r = tmp_path / "int.mtx"
    r.write_text("%%MatrixMarket matrix coordinate integer general\n2 2 2\n2 2 -7\n1 2 3\n")
    Cm = sp.read_kokkos_crst_matrix(r)
    assert Cm.row_map.tolist() == [0, 1, 2] and Cm.entries.tolist() == [1, 1] and Cm.values.tolist() == [3.0, -7.0]


@pytest.mark.parametrize("text,msg", [
    ("garbage\n1 1 1\n", "Invalid MM file"),
    ("%%MatrixMarket vector coordinate real general\n1 1\n", "vector"),
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1 0\n", "complex"),
    ("%%MatrixMarket matrix coordinate real symmetric\n2 3 1\n1 1 1\n", "non-square"),
    ("%%MatrixMarket matrix array real symmetric\n2 2\n1\n2\n3\n4\n", "array format"),
    ("%%MatrixMarket matrix coordinate real\n2 2 1\n1 1 1\n", "symmetry type"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1\n", "ends after"),
])
def test_read_mtx_errors(tmp_path, text, msg):
    p = tmp_path / "bad.mtx"
    p.write_text(text)
    with pytest.raises(B200SparseError, match=msg):
        sp.read_kokkos_crst_matrix(p)
    with pytest.raises(B200SparseError, match="opened"):
        sp.read_kokkos_crst_matrix(tmp_path / "missing.mtx")
    with pytest.raises(B200SparseError, match="extension"):
        sp.read_kokkos_crst_matrix(tmp_path / "x.txt")


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
def test_against_scipy(tmp_path):
    sio = pytest.importorskip("scipy.io")
    ssp = pytest.importorskip("scipy.sparse")
    M = ssp.random(200, 150, density=0.05, random_state=3, format="coo")
    p = tmp_path / "sci.mtx"
    sio.mmwrite(str(p), M, precision=17)
    A = sp.read_kokkos_crst_matrix(p)
    R = ssp.csr_matrix(sio.mmread(str(p)))
    R.sort_indices()
    assert A.numRows() == 200 and A.numCols() == 150
    assert np.array_equal(A.row_map.numpy(), R.indptr) and np.array_equal(A.entries.numpy(), R.indices)
    assert np.array_equal(A.values.numpy(), R.data)
