"""BsrMatrix inputs and the acceptance law of the reference's BsrMatrix SpMV unit test
(sparse/unit_test/Test_Sparse_spmv_bsr.hpp), shared by the oracle tests, the emulated-kernel tests and the GPU
parity tests."""
import numpy as np

MAX_A = MAX_X = MAX_Y = 10.0  # getRandomBounds(10.0, ...) for real scalars (:57-80)

# (block rows, block columns) of test_spmv_random (:428) + the "tougher case" (:444-447); block sizes (:431)
SHAPES = [(10, 10), (10, 50), (50, 10)]
BLOCK_SIZES = [1, 2, 5, 9]
PRIME_CASE = (7, 11, 499)  # block size, block rows, block columns
COEFS_ALPHA = [0.0, 1.0, -1.0, 3.7]  # test_spmv_combos (:386-387)
COEFS_BETA = [0.0, 1.0, -1.0, -1.5]


def bsr_random(bs, mb, nb, seed=0, dtype=np.float64, min_blocks=0, max_blocks=None, sort=True):
    """Random block structure (block rows of min_blocks..max_blocks distinct block columns) with values uniform in
    [0, MAX_A) -- the role of bsr_random (:118-137: a random CrsMatrix expanded to blocks)."""
    rng = np.random.default_rng(seed)
    if max_blocks is None:
        max_blocks = max(1, min(nb, 8))
    max_blocks = min(max_blocks, nb)
    lens = rng.integers(min(min_blocks, max_blocks), max_blocks + 1, mb) if nb > 0 else np.zeros(mb, np.int64)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cols = []
    for l in lens:
        c = rng.choice(nb, int(l), replace=False)
        cols.append(np.sort(c) if sort else c)
    ci = (np.concatenate(cols) if cols else np.zeros(0)).astype(np.int32)
    v = rng.uniform(0.0, MAX_A, len(ci) * bs * bs).astype(dtype)
    return rp, ci, v


def op_max_nnz_per_row(bs, rp, ci, nb, trans):
    """opMaxNnzPerRow (:87-99): block size x the largest block-row degree of Op(A)."""
    if not trans:
        deg = int(np.diff(rp).max()) if len(rp) > 1 else 0
    else:
        deg = int(np.bincount(ci, minlength=max(nb, 1)).max()) if len(ci) else 0
    return bs * deg


def tolerance(dtype, alpha, beta, max_nnz_per_row):
    """:174-176: eps*|beta|*max_y + 10*eps*maxNnzPerRow*|alpha|*max_a*max_x."""
    eps = np.finfo(dtype).eps
    return eps * abs(beta) * MAX_Y + 10 * eps * max_nnz_per_row * abs(alpha) * MAX_A * MAX_X
