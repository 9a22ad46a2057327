"""Golden tables of the reference's merge-matrix unit test
(sparse/unit_test/Test_Sparse_MergeMatrix.hpp:129-346,411-505) against the oracle's
restatement of MergeMatrixDiagonal / diagonal_search."""
import numpy as np


CASES_VIEW_VIEW = {
    # name: (a, b, {diagonal: entries})
    "case_1": ([1, 2, 3, 4], [0, 1, 2, 3],
               {0: [], 1: [1], 2: [1, 0], 3: [1, 1, 0], 4: [1, 1, 0, 0], 5: [1, 1, 0], 6: [1, 0], 7: [1]}),
    "case_2": ([1, 2, 9], [0, 2, 2, 8, 8, 8],
               {0: [], 1: [1], 2: [1, 0], 3: [1, 0, 0], 4: [1, 0, 0], 5: [1, 0, 0], 6: [1, 0, 0], 7: [1, 0], 8: [1]}),
    "case_3": ([-1, 9, 9], [0, 2, 7], {0: [], 1: [0], 2: [1, 0], 3: [1, 1, 0], 4: [1, 1], 5: [1]}),
    "case_4": ([1, 6, 6], [-3, -1, 7], {0: [], 1: [1], 2: [1, 1], 3: [1, 1, 0], 4: [1, 0], 5: [0]}),
    "case_5": ([-3, -2, 2], [-2, 0, 1], {0: [], 1: [0], 2: [0, 0], 3: [1, 0, 0], 4: [1, 0], 5: [1]}),
}


def test_view_view_tables(oracle):
    for name, (a, b, table) in CASES_VIEW_VIEW.items():
        for d, exp in table.items():
            assert oracle.mmd_entries(a, b, d) == exp, (name, d)


def test_all_zero_all_one(oracle):
    a0, b = [0, 0, 0, 0], [0, 1, 2, 3]
    for d in range(len(a0) + len(b) - 1):
        assert oracle.mmd_entries(a0, b, d) == [0] * oracle.mmd_size(4, 4, d)
    a1, b1 = [1, 2, 3, 4], [0, 0, 0, 0]
    for d in range(7):
        assert oracle.mmd_entries(a1, b1, d) == [1] * oracle.mmd_size(4, 4, d)


def test_view_iota_tables(oracle):
    exp = {0: [], 1: [1], 2: [1, 0], 3: [1, 1, 0], 4: [1, 1, 0, 0], 5: [1, 1, 0], 6: [1, 0], 7: [1]}
    for d, e in exp.items():
        assert oracle.mmd_entries([1, 2, 3, 4], 4, d) == e
    for d in range(7):
        assert oracle.mmd_entries([0, 0, 0, 0], 4, d) == [0] * oracle.mmd_size(4, 4, d)
        assert oracle.mmd_entries([5, 6, 7, 8], 4, d) == [1] * oracle.mmd_size(4, 4, d)


def test_empty_shapes(oracle):
    assert oracle.mmd_size(0, 0, 0) == 0
    for d in range(4):
        assert oracle.mmd_size(5, 0, d) == 0
        assert oracle.mmd_size(0, 4, d) == 0


def test_diagonal_search_partitions_spmv_path(oracle):
    """diagonal_search over (row_ends, iota(nnz)) is the merge-path split used by
    SpmvMergeHierarchical (spmv_impl_merge.hpp:104-130): positions are monotone and
    consume exactly `d` path items."""
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 9, size=200)
    row_ends = np.cumsum(lens)
    nnz = int(row_ends[-1])
    prev = (0, 0)
    for d in range(0, 200 + nnz + 1, 7):
        ai, bi = oracle.diagonal_search(row_ends, nnz, d)
        assert ai + bi == d
        assert ai >= prev[0] and bi >= prev[1]
        # all rows before ai end at or before bi; row ai (if any) ends after bi-1
        if ai > 0:
            assert row_ends[ai - 1] <= bi
        if ai < 200 and bi > 0:
            assert row_ends[ai] > bi - 1
        prev = (ai, bi)
