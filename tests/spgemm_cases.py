"""SpGEMM inputs that drive each path of the expand / sort / compress kernels (kokkos-kernels_b200/csrc/spgemm_esc.cuh):
shared by the emulated (CPU) and the GPU parity tests.  Every case is (name, (rpA, ciA, vA), (rpB, ciB, vB), m, n, k)."""
import numpy as np


def _csr(lens, cols_of_row, ncols, rng, lo=1.0, hi=50.0, dtype=np.float64, sort=True):
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = np.empty(int(rp[-1]), dtype=np.int32)
    for r, ln in enumerate(lens):
        c = cols_of_row(r, int(ln))
        ci[rp[r]:rp[r + 1]] = np.sort(c) if sort else c
    v = rng.uniform(lo, hi, len(ci)).astype(dtype)
    return rp, ci, v


def uniform_random(m, per, seed, dtype=np.float64, sort=True):
    """config-4 shape: exactly `per` distinct uniform-random columns per row (square); A*A has per^2 products per row,
    almost no duplicate columns -> the uniform-length fast path and the duplicate-free exit."""
    rng = np.random.default_rng(seed)
    lens = np.full(m, per)
    return _csr(lens, lambda r, ln: rng.choice(m, ln, replace=False), m, rng, dtype=dtype, sort=sort)


def banded(m, per, band, seed, dtype=np.float64):
    """narrow band: many duplicate columns per product row (runs to add up), dense bucket map"""
    rng = np.random.default_rng(seed)
    lens = rng.integers(max(1, per - 3), per + 4, m)

    def cols(r, ln):
        lo, hi = max(0, r - band), min(m, r + band + 1)
        return rng.choice(np.arange(lo, hi), min(ln, hi - lo), replace=False)

    lens = np.array([min(l, min(m, r + band + 1) - max(0, r - band)) for r, l in enumerate(lens)])
    return _csr(lens, cols, m, rng, dtype=dtype)


def cases(dtype=np.float64, big=False):
    out = []
    rng = np.random.default_rng(77)
    # 1. uniform 32 per row (1024 products per row: bin 1, T=128) -- config 4 in small
    A = uniform_random(3000 if not big else 200000, 32, 4, dtype)
    out.append(("uniform32", A, A, len(A[0]) - 1, len(A[0]) - 1, len(A[0]) - 1))
    # 2. uniform 16 per row, unsorted input rows (256 products: bin 0, one warp)
    A = uniform_random(2000, 16, 5, dtype, sort=False)
    out.append(("uniform16_unsorted", A, A, 2000, 2000, 2000))
    # 3. uniform 12 per row (L0 not a power of two, 144 products)
    A = uniform_random(1500, 12, 6, dtype)
    out.append(("uniform12", A, A, 1500, 1500, 1500))
    # 4. banded: heavy duplication, dense bucket map when the span is small
    A = banded(2500, 14, 20, 7, dtype)
    out.append(("banded", A, A, 2500, 2500, 2500))
    # 5. a wider band with longer rows: 2048..4096 products (bin 2, T=512), duplicates
    A = banded(600, 55, 300, 8, dtype)
    out.append(("banded_wide", A, A, 600, 600, 600))
    # 6. 80 per row uniform: 6400 products (bin 3, T=1024)
    A = uniform_random(700, 80, 9, dtype)
    out.append(("uniform80", A, A, 700, 700, 700))
    # 7. rectangular with empty rows in A and in B, single-entry rows, a row of A longer than half its bin
    m, n, k = 900, 700, 1100
    lensA = rng.integers(0, 9, m)
    lensA[3] = 0
    lensA[10] = 300  # 2 * nnz(A_i) = 600 decides the bin, its B rows are short
    lensB = rng.integers(0, 7, n)
    lensB[:20] = 0
    Am = _csr(lensA, lambda r, ln: rng.choice(n, ln, replace=False), n, rng, dtype=dtype)
    Bm = _csr(lensB, lambda r, ln: rng.choice(k, ln, replace=False), k, rng, dtype=dtype)
    out.append(("ragged_rect", Am, Bm, m, n, k))
    # 8. skew: one far column per row stretches the span, the rest sits in a narrow band (crowded buckets)
    m = 1200
    lens = np.full(m, 20)

    def skew_cols(r, ln):
        lo = max(0, min(m - 40, r))
        c = rng.choice(np.arange(lo, lo + 30), ln - 1, replace=False)
        return np.concatenate([c, [m - 1 - (r % 3)]]) if (m - 1 - (r % 3)) not in c else np.concatenate([c, [lo + 35]])

    As = _csr(lens, skew_cols, m, rng, dtype=dtype)
    out.append(("skewed_span", As, As, m, m, m))
    # 9. rows beyond the ESC capacity next to small ones (hash kernels + ESC in one product)
    m, n, k = 300, 2000, 30000
    lensA = rng.integers(1, 5, m)
    lensA[7] = 400
    lensB = np.full(n, 30)
    Am = _csr(lensA, lambda r, ln: rng.choice(n, ln, replace=False), n, rng, dtype=dtype)
    Bm = _csr(lensB, lambda r, ln: rng.choice(k, ln, replace=False), k, rng, dtype=dtype)
    out.append(("mixed_long_row", Am, Bm, m, n, k))
    return out
