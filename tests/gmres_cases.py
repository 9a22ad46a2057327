"""The problem family of the reference's GMRES unit test (sparse/unit_test/Test_Sparse_gmres.hpp:52-60,86-110): a banded,
diagonally dominant matrix with about 10 entries per row, diagonal = diagDominance x the sum of the row's off-diagonal
magnitudes (kk_diagonally_dominant_sparseMatrix_generate, sparse/src/KokkosSparse_IOUtils.hpp:112-177: off-diagonals uniform
in [-50, 50) inside a band of 0.01 n around the diagonal, wrapped), rows sorted; B = 1, X = 0, m = 15."""
import numpy as np


def gmres_matrix(n, dominance=1.0, per=10, seed=13721, dtype=np.float64):
    rng = np.random.default_rng(seed)
    bw = max(int(0.01 * n), 2 * per)
    rp, ci, v = [0], [], []
    for row in range(n):
        cols = set()
        while len(cols) < per - 1:
            pos = int((rng.random() - 0.5) * bw + row) % n
            if pos != row:
                cols.add(pos)
        vals = {c: 100.0 * rng.random() - 50.0 for c in cols}
        vals[row] = dominance * sum(abs(t) for t in vals.values())
        for c in sorted(vals):
            ci.append(c)
            v.append(vals[c])
        rp.append(len(ci))
    return np.array(rp, np.int32), np.array(ci, np.int32), np.array(v, dtype=dtype)


def true_rel_res(oracle, A, b, x):
    rp, ci, v = A
    y = np.zeros(len(b), dtype=np.float64)
    oracle.spmv_serial(rp, ci, v.astype(np.float64), x.astype(np.float64), y, 1.0, 0.0)
    return np.linalg.norm(b.astype(np.float64) - y) / np.linalg.norm(b.astype(np.float64))


def crs_to_bsr(rp, ci, v, bs):
    """BsrMatrix(const CrsMatrix&, blockDim) (sparse/src/KokkosSparse_BsrMatrix.hpp:520-610): the block structure of the point
    matrix with explicit zeros where a block is only partly filled.  n must be a multiple of bs."""
    n = len(rp) - 1
    assert n % bs == 0
    mb = n // bs
    brp, bci, bv = [0], [], []
    for br in range(mb):
        cols = sorted({int(c) // bs for r in range(br * bs, br * bs + bs) for c in ci[rp[r]:rp[r + 1]]})
        pos = {c: k for k, c in enumerate(cols)}
        blk = np.zeros((len(cols), bs, bs), dtype=v.dtype)
        for lr in range(bs):
            r = br * bs + lr
            for q in range(rp[r], rp[r + 1]):
                blk[pos[int(ci[q]) // bs], lr, int(ci[q]) % bs] = v[q]
        bci.extend(cols)
        bv.append(blk.reshape(-1))
        brp.append(len(bci))
    return np.array(brp, np.int32), np.array(bci, np.int32), (np.concatenate(bv) if bv else np.zeros(0, v.dtype))
