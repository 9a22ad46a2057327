"""Row-block-partitioned SpMV across the GPUs of one box with the all-gather of y pipelined behind
the compute (BASELINE.json configs[4]; no reference counterpart -- the reference is single-process).

Each rank owns rows [r0, r1) of A (row_ptr rebased, global columns) and a full copy of x.  The local
rows are cut into `chunks` pieces; piece c is computed by the TMA-tiled SpMV kernel straight into this
rank's slot of the next-x buffer and, as soon as it is done, pushed to the same slot of every peer's
next-x buffer by copy-engine transfers over NVLink (peer buffers are mapped through symmetric memory),
while piece c+1 is being computed.  A device-side barrier closes the step.  Results are bit-identical
to the single-GPU SpMV (same kernel, same per-row order).

modes: "pipelined" (default), "fused" (P2P stores issued by the SpMV kernel itself,
b200sp_spmv_scatter_f64_i32), "multicast" (as fused, but ONE store per y value to the NVSwitch multicast
address of the symmetric buffer -- on sm_100 multimem.st is a plain st.global to a multicast mapping, the
switch replicates it into all 8 copies, so every GPU sends its 80 MB once instead of 7 times; falls back
to "fused" when the symmetric-memory handle has no multicast pointer; first measurements: round 2),
"pipelined_mc" (pieces like "pipelined", but each finished piece is pushed ONCE to the multicast address by a
small SM kernel with 16-byte stores, b200sp_multicast_push, on one communication stream; falls back to
"pipelined" without a multicast pointer),
"nccl" (SpMV then all_gather_into_tensor)."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, partition, sparse as sp


class RowBlockSpMV:
    def __init__(self, rp, ci, va, n_total, r0, r1, device, mode="pipelined", chunks=8, tune=(-1, -1, -1)):
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.n_total, self.r0, self.r1, self.dev, self.mode = n_total, r0, r1, device, mode
        nrows = r1 - r0
        assert n_total % self.world == 0 and nrows == n_total // self.world, "equal row blocks expected"
        self.ci_d = torch.from_numpy(ci).to(device)
        self.va_d = torch.from_numpy(va).to(device)
        self.symm = None
        if mode in ("pipelined", "fused", "multicast", "pipelined_mc"):
            import torch.distributed._symmetric_memory as symm_mem

            self.x_next = symm_mem.empty(n_total, dtype=torch.float64, device=device)
            self.symm = symm_mem.rendezvous(self.x_next, dist.group.WORLD)
            self.peer_ptrs = [int(p) for p in self.symm.buffer_ptrs]
            self.mc_ptr = int(getattr(self.symm, "multicast_ptr", 0) or 0)
            if mode == "multicast" and self.mc_ptr == 0:
                self.mode = mode = "fused"  # no NVLS multicast mapping on this box
            if mode == "pipelined_mc" and self.mc_ptr == 0:
                self.mode = mode = "pipelined"
        else:
            self.x_next = torch.empty(n_total, dtype=torch.float64, device=device)
        self.y = self.x_next[r0:r1]
        # chunk boundaries on rows whose first entry is 16-byte aligned in col_idx / vals (TMA path)
        if mode not in ("pipelined", "pipelined_mc"):
            chunks = 1
        bounds = partition.piece_bounds(rp, chunks, row_offset=r0, even_rows=(mode == "pipelined_mc"))
        self.pieces = []
        for c0, c1 in zip(bounds[:-1], bounds[1:]):
            s0, s1 = int(rp[c0]), int(rp[c1])
            rpc = torch.from_numpy((rp[c0:c1 + 1].astype(np.int64) - s0).astype(np.int32)).to(device)
            A = sp.CrsMatrix(rpc, self.ci_d[s0:s1], self.va_d[s0:s1], n_total)
            h = sp.SPMVHandle(sp.SPMV_DEFAULT)
            h.tune(*tune)
            yv = self.y[c0:c1]
            dsts = []
            if self.symm is not None and mode in ("multicast", "pipelined_mc"):
                dsts = [self.mc_ptr + (r0 + c0) * 8]  # one store, replicated by the switch (incl. this rank's copy)
            elif self.symm is not None:
                dsts = [self.peer_ptrs[q] + (r0 + c0) * 8 for q in range(self.world) if q != self.rank]
            arr = (C.c_void_p * max(len(dsts), 1))(*[C.c_void_p(d) for d in dsts])
            self.pieces.append((A, h, yv, dsts, arr, torch.cuda.Event()))
        # whole-shard view (host-vector end-to-end path, parity checks)
        self.A_full = sp.CrsMatrix(torch.from_numpy(np.ascontiguousarray(rp)).to(device), self.ci_d, self.va_d, n_total)
        self.h_full = sp.SPMVHandle(sp.SPMV_DEFAULT)
        # one communication stream per peer: the pushes of a piece run concurrently on the copy engines
        self.comm = [torch.cuda.Stream(device=device) for _ in range(max(self.world - 1, 1))]
        self.comm_arr = (C.c_void_p * len(self.comm))(*[C.c_void_p(s.cuda_stream) for s in self.comm])

    def kernel_name(self):
        return self.pieces[0][1].last_kernel()

    def nnz(self):
        return self.ci_d.numel()

    def step(self, x):
        """x_next <- all-gather(A_local @ x); returns the next-x buffer (valid on every rank after the call's
        stream work completes)."""
        lib = _lib.sparse()
        cur = torch.cuda.current_stream()
        if self.mode in ("fused", "multicast"):
            A, h, yv, dsts, arr, ev = self.pieces[0]
            sp.spmv_scatter(h, 1.0, A, x, yv, dsts)
            self.symm.barrier(channel=0)
        elif self.mode == "pipelined_mc":
            cs = C.c_void_p(cur.cuda_stream)
            comm = self.comm[0]
            for A, h, yv, dsts, arr, ev in self.pieces:
                sp.spmv(h, "N", 1.0, A, x, 0.0, yv)
                ev.record(cur)
                comm.wait_event(ev)
                _lib.check(lib.b200sp_multicast_push(C.c_void_p(comm.cuda_stream), C.c_void_p(yv.data_ptr()),
                                                     C.c_void_p(dsts[0]), yv.numel() * 8, 16))
            _lib.check(lib.b200sp_peer_join(cs, self.comm_arr, 1))
            self.symm.barrier(channel=0)
        elif self.mode == "pipelined":
            cs = C.c_void_p(cur.cuda_stream)
            for A, h, yv, dsts, arr, ev in self.pieces:
                sp.spmv(h, "N", 1.0, A, x, 0.0, yv)
                _lib.check(lib.b200sp_peer_push_async(cs, self.comm_arr, len(dsts), arr, C.c_void_p(yv.data_ptr()),
                                                      yv.numel() * 8))
            _lib.check(lib.b200sp_peer_join(cs, self.comm_arr, len(self.comm) if self.world > 1 else 0))
            self.symm.barrier(channel=0)
        else:
            A, h, yv, dsts, arr, ev = self.pieces[0]
            sp.spmv(h, "N", 1.0, A, x, 0.0, yv)
            dist.all_gather_into_tensor(self.x_next, yv)
        return self.x_next

    def local_spmv_only(self, x):
        for A, h, yv, dsts, arr, ev in self.pieces:
            sp.spmv(h, "N", 1.0, A, x, 0.0, yv)


class RowBlockSpGEMM:
    """C = A * B with the rows of A (hence of C) sharded over the ranks and B replicated ("SpGEMM also
    shards by row blocks of A with B replicated; rowptr needs one exchange of P counts").  Every rank runs the single-GPU
    spgemm_symbolic / spgemm_numeric of the library on its row block -- the rows of C are independent, so the block is
    bit-identical to the same rows of the single-GPU product -- and ONE all_gather of the P block sizes gives each rank the
    offset of its block in the global C (no collective on the data path).  No reference counterpart (single-process library).

        op = RowBlockSpGEMM(A_local, B, group=None)   # A_local: rows [r0, r1) of A, row map rebased, global columns
        C_local = op.symbolic()                       # structure of this rank's rows; op.block_nnz / op.offset are known
        op.numeric()                                  # values; re-runnable with new values on the same structure
        op.global_row_map()                           # int64: C_local.row_map + op.offset (global nnz may exceed int32)
    """

    def __init__(self, A_local, B, group=None):
        if A_local.numCols() != B.numRows():
            raise sp.B200SparseError(f"RowBlockSpGEMM: inner dimensions differ: {A_local.numCols()} vs {B.numRows()}")
        self.A, self.B, self.group = A_local, B, group
        self.kh = sp.KokkosKernelsHandle()
        self.kh.create_spgemm_handle(sp.SPGEMM_KK)
        self.C = None
        self.block_nnz = None  # nnz of every rank's block of C
        self.offset = None     # where this rank's block starts in the global entries / values of C

    def symbolic(self):
        self.C = sp.spgemm_symbolic(self.kh, self.A, False, self.B, False)
        mine = int(self.kh.get_spgemm_handle().get_c_nnz())
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            counts = [None] * world
            dist.all_gather_object(counts, mine, group=self.group)
            rank = dist.get_rank(self.group)
        else:
            counts, rank = [mine], 0
        self.block_nnz = [int(c) for c in counts]
        self.offset = int(sum(self.block_nnz[:rank]))
        return self.C

    def numeric(self, A_values=None, B_values=None):
        if self.C is None:
            raise sp.B200SparseInvalidArgument("RowBlockSpGEMM.numeric: call symbolic first")
        A = self.A if A_values is None else sp.CrsMatrix(self.A.row_map, self.A.entries, A_values, self.A.numCols())
        B = self.B if B_values is None else sp.CrsMatrix(self.B.row_map, self.B.entries, B_values, self.B.numCols())
        sp.spgemm_numeric(self.kh, A, False, B, False, self.C)
        return self.C

    def global_row_map(self):
        return self.C.row_map.to(torch.int64) + self.offset
