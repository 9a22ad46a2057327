"""Row-block partition of a CrsMatrix across the GPUs of one box (BASELINE.json configs[4]).
The reference has no distributed layer (README.md:12-16); the nearest piece is
StaticCrsGraph::create_block_partitioning (sparse/src/KokkosSparse_StaticCrsGraph.hpp:340-353),
balanced row blocks by nnz, which `balanced_row_blocks` mirrors.  The only collective on the
data path is the all-gather of y."""
import numpy as np


def equal_row_blocks(n_rows, world):
    """Block r owns rows [n*r//world, n*(r+1)//world)."""
    return [(n_rows * r) // world for r in range(world + 1)]


def balanced_row_blocks(row_ptr, world):
    """Blocks with ~equal nnz: boundary r = first row whose offset >= r*nnz/world."""
    nnz = int(row_ptr[-1])
    targets = (np.arange(world + 1, dtype=np.int64) * nnz) // world
    b = np.searchsorted(row_ptr, targets, side="left")
    b[0], b[-1] = 0, len(row_ptr) - 1
    return [int(v) for v in np.maximum.accumulate(b)]


def piece_bounds(row_ptr, chunks, row_offset=0, even_rows=False):
    """Boundaries (local row indices, first = 0, last = nrows) of the pieces a row block is computed / pushed in by
    the pipelined all-gather (multigpu.RowBlockSpMV): about nrows/chunks rows each, every piece starting on a row
    whose first entry is 16-byte aligned in col_idx / vals (row_ptr % 4 == 0: the TMA-tiled kernel is handed a
    sub-matrix view) and -- even_rows, for the 16-byte multicast pushes -- whose global row index
    (row_offset + r) is even, so that the piece of y starts on a 16-byte boundary.  Pieces that cannot satisfy
    the constraints are merged into their predecessor."""
    nrows = len(row_ptr) - 1
    bounds = [0]
    for c in range(1, max(int(chunks), 1)):
        r = (nrows * c) // chunks
        while r < nrows and (row_ptr[r] % 4 != 0 or (even_rows and (row_offset + r) % 2 != 0)):
            r += 1
        if bounds[-1] < r < nrows:
            bounds.append(r)
    bounds.append(nrows)
    return bounds


def extract_shard(row_ptr, col_idx, values, r0, r1):
    """Rows [r0, r1) with offsets rebased to 0 (fits int32 per shard); columns stay global."""
    s, e = int(row_ptr[r0]), int(row_ptr[r1])
    rp = (row_ptr[r0:r1 + 1].astype(np.int64) - s).astype(np.int32)
    return rp, col_idx[s:e], values[s:e]


def allgather_y(y_local, x_next, bounds, rank, group=None):
    """x_next[bounds[r]:bounds[r+1]] <- rank r's y_local.  Equal blocks use one
    all_gather_into_tensor (NCCL on GPUs, gloo in the CPU tests); ragged blocks fall back to
    all_gather on padded pieces."""
    import torch
    import torch.distributed as dist

    world = len(bounds) - 1
    sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(x_next, y_local, group=group)
        return x_next
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=y_local.dtype, device=y_local.device)
    pad[: sizes[rank]] = y_local
    pieces = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(pieces, pad, group=group)
    for r in range(world):
        x_next[bounds[r]:bounds[r + 1]] = pieces[r][: sizes[r]]
    return x_next
