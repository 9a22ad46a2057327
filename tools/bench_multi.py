"""Micro-benchmark of the fused SpMV + all-gather step on N GPUs (torchrun): plain SpMV, scatter SpMV
without / with the symmetric-memory barrier, barrier alone, NCCL all-gather alone."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kokkos_kernels_b200 import matgen, sparse as sp  # noqa: E402


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem

    grid = int(os.environ.get("GRID", "171"))
    rp, ci, va, n_total, r0, r1 = bench.build_shard(world, rank, grid)
    A = sp.CrsMatrix(torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev), torch.from_numpy(va).to(dev), n_total)
    x = torch.from_numpy(matgen.fill(n_total, -1, 1, 1)).to(dev)
    xn = symm_mem.empty(n_total, dtype=torch.float64, device=dev)
    hdl = symm_mem.rendezvous(xn, dist.group.WORLD)
    ptrs = list(hdl.buffer_ptrs)
    extra = [ptrs[q] + r0 * 8 for q in range(world) if q != rank]
    y = xn[r0:r1]
    h = sp.SPMVHandle()
    xn2 = torch.empty(n_total, dtype=torch.float64, device=dev)
    y2 = torch.empty(r1 - r0, dtype=torch.float64, device=dev)
    import ctypes as C
    from kokkos_kernels_b200 import _lib
    arr = (C.c_void_p * max(len(extra), 1))(*[C.c_void_p(int(q)) for q in extra])

    def push_all():
        _lib.check(_lib.sparse().b200sp_peer_push(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(y2.data_ptr()),
                                                  y2.numel() * 8, len(extra), arr))

    res = {
        "spmv_plain_ms": timeit(lambda: sp.spmv(h, "N", 1.0, A, x, 0.0, y2)),
        "spmv_scatter_ms": timeit(lambda: sp.spmv_scatter(h, 1.0, A, x, y, extra)),
        "spmv_scatter_barrier_ms": timeit(lambda: (sp.spmv_scatter(h, 1.0, A, x, y, extra), hdl.barrier(channel=0))),
        "peer_push_80MB_x7_ms": timeit(lambda: push_all()),
        "barrier_ms": timeit(lambda: hdl.barrier(channel=0)),
        "nccl_allgather_ms": timeit(lambda: dist.all_gather_into_tensor(xn2, y2)),
        "spmv_plus_nccl_ms": timeit(lambda: (sp.spmv(h, "N", 1.0, A, x, 0.0, y2), dist.all_gather_into_tensor(xn2, y2))),
    }
    if rank == 0:
        print({k: round(v, 4) for k, v in res.items()}, "world", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
