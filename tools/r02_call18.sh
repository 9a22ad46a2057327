#!/bin/bash
# Round 2, GPU call 18 (single B200): BsrMatrix rank-1 product through the tensor-core kernel (bs 6..16, double) against the
# row-vector kernel at full size; the BsrMatrix suite on hardware.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c18
timeout 900 python -m pytest tests/test_gpu_bsr.py tests/test_gpu_gmres.py -q -x > ${O}_pytest.log 2>&1; tail -n 2 ${O}_pytest.log
G=./kokkos-kernels_b200/lib/gpu_check
timeout 400 $G --big --suite bsr --out ${O}_bsr_big.jsonl > ${O}_bsr_big_default.log 2>&1
B200SP_BSR_KERNEL=vector timeout 400 $G --big --suite bsr --out gpurun_out/scratch.jsonl > ${O}_bsr_big_vector.log 2>&1
grep -E "N_default" ${O}_bsr_big_default.log | cut -c1-200
grep -E "N_default" ${O}_bsr_big_vector.log | cut -c1-200 | grep -E "bs8|bs5|bs16|bs6|bs7"
