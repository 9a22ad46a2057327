#!/bin/bash
# First GPU call of the next round (single B200, nothing else on the GPU): the FIRST GPU RUN of everything written after round 1's
# GPU minutes were spent (validated under the CPU emulation only), then the headline bench.  About 10 minutes of box time:
#   1. harness suites bsr / cg / jacobi / solvers / spmv64 (oracle-checked, with timings next to the SpMV's)
#   2. `pytest -m gpu_next`: BsrMatrix SpMV / SpMM, CG / PCG, GMRES, point and two-stage Gauss-Seidel, spgemm_jacobi, 64-bit offsets,
#      host-vector deferred completion, the shim driver with every flag
#   3. the headline bench (SpMV config 2; its e2e leg times the stream-ordered and the deferred host-vector modes)
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# then tools/round2_promote.sh (gpu_next -> gpu) and, in a second call, tools/round2_second_call.sh (full-size timings of every suite incl. the 2.2e9-entry matrix, ncu).
# Outputs land in gpurun_out/ (copy the summaries to profiles/ afterwards).  On success change `gpu_next` to `gpu` in
# tests/test_gpu_{jacobi,bsr,cg,gmres,gs,gs2,spmv64,hostvec_defer}.py and tests/test_shim.py.
set -u
mkdir -p gpurun_out
G=./kokkos-kernels_b200/lib/gpu_check
O=gpurun_out/r02_gpu_check_first.jsonl
L=gpurun_out/r02_gpu_check_first.log
: > $L
for s in bsr cg jacobi solvers spmv64; do timeout 400 $G --suite $s --out $O >> $L 2>&1; done
timeout 600 python -m pytest tests -q -m gpu_next > gpurun_out/r02_pytest_gpu_next.log 2>&1; tail -n 5 gpurun_out/r02_pytest_gpu_next.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
grep -E "FAIL|summary" $L | head -60
tail -1 gpurun_out/r02_bench_n1.json | cut -c1-600
ls -la gpurun_out | tail -20
