"""Inputs and acceptance laws of the reference's SortCrs / spadd unit tests, shared by the oracle
(CPU) tests and the GPU parity tests."""
import numpy as np

_i = lambda a: np.array(a, dtype=np.int32)
_f = lambda a: np.array(a, dtype=np.float64)

# testSortAndMerge cases 0-4 (sparse/unit_test/Test_Sparse_SortCrs.hpp:195-290), values exactly
# representable in float
MERGE_CASES = {
    0: dict(nrows=5, ncols=7, rowmap=_i([0, 4, 4, 5, 7, 10]), entries=_i([4, 3, 5, 3, 6, 2, 2, 0, 1, 2]),
            values=_f([1.5, 4, 1, -3, 2, -1, -2, 0, 3.5, -2.25]), gold_rowmap=_i([0, 3, 3, 4, 5, 8]),
            gold_entries=_i([3, 4, 5, 6, 2, 0, 1, 2]), gold_values=_f([1, 1.5, 1, 2, -3, 0, 3.5, -2.25])),
    1: dict(nrows=5, ncols=7, rowmap=_i([0, 3, 3, 4, 5, 8]), entries=_i([4, 5, 3, 6, 2, 0, 1, 2]),
            values=_f([1.5, 4, 1, 2, -1, 0, 3.5, -2.25]), gold_rowmap=_i([0, 3, 3, 4, 5, 8]),
            gold_entries=_i([3, 4, 5, 6, 2, 0, 1, 2]), gold_values=_f([1, 1.5, 4, 2, -1, 0, 3.5, -2.25])),
    2: dict(nrows=5, ncols=7, rowmap=_i([0, 0, 0, 0, 0, 0]), entries=_i([]), values=_f([]),
            gold_rowmap=_i([0, 0, 0, 0, 0, 0]), gold_entries=_i([]), gold_values=_f([])),
    3: dict(nrows=0, ncols=0, rowmap=_i([]), entries=_i([]), values=_f([]), gold_rowmap=_i([]), gold_entries=_i([]),
            gold_values=_f([])),
    4: dict(nrows=0, ncols=0, rowmap=_i([0]), entries=_i([]), values=_f([]), gold_rowmap=_i([0]), gold_entries=_i([]),
            gold_values=_f([])),
}


def random_matrix(nrows, ncols, min_nnz, max_nnz, sort_rows, seed=0, dtype=np.float64):
    """randomMatrix of Test_Sparse_spadd.hpp:41-94: row lengths uniform in [min_nnz, max_nnz]; the
    columns of a row are the first entries of a shuffle of (j % ncols), so rows longer than ncols
    repeat columns; values uniform in [0, 1]."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_nnz, max_nnz + 1, nrows) if max_nnz >= min_nnz else np.zeros(nrows, np.int64)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    width = max(ncols, int(lens.max()) if nrows else 0)
    ci = np.empty(int(rp[-1]), dtype=np.int32)
    for i in range(nrows):
        idx = (np.arange(width) % max(ncols, 1)).astype(np.int32)
        rng.shuffle(idx)
        row = idx[: lens[i]]
        if sort_rows:
            row = np.sort(row)
        ci[rp[i]:rp[i + 1]] = row
    v = rng.uniform(0.0, 1.0, int(rp[-1])).astype(dtype)
    return rp, ci, v


def spadd_dense_check(A, B, Cm, ncols, alpha, beta):
    """test_spadd's row-by-row check (Test_Sparse_spadd.hpp:143-186): right count, sorted and unique
    columns, values within 1 ulp of the dense row sum."""
    rpA, ciA, vA = A
    rpB, ciB, vB = B
    rpC, ciC, vC = Cm
    eps = np.finfo(vC.dtype).eps
    for row in range(len(rpA) - 1):
        correct = np.zeros(ncols, dtype=vC.dtype)
        nonzeros = np.zeros(ncols, dtype=bool)
        for i in range(rpA[row], rpA[row + 1]):
            correct[ciA[i]] += vC.dtype.type(alpha) * vA[i]
            nonzeros[ciA[i]] = True
        for i in range(rpB[row], rpB[row + 1]):
            correct[ciB[i]] += vC.dtype.type(beta) * vB[i]
            nonzeros[ciB[i]] = True
        s, e = rpC[row], rpC[row + 1]
        assert e - s == int(nonzeros.sum()), f"A+B row {row} has {e - s} entries, expected {int(nonzeros.sum())}"
        cols = ciC[s:e]
        assert np.all(np.diff(cols) > 0), f"C row {row} is not sorted / unique"
        assert np.all(nonzeros[cols])
        want = correct[cols]
        tol = np.where(want == 0, eps, np.abs(want) * eps * 4)  # a few ulp: sums of up to ~4 terms per column
        assert np.all(np.abs(want - vC[s:e]) <= tol), f"A+B row {row}: values off"
