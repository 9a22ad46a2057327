#!/bin/bash
# Round 2, GPU call 2 (single B200): first run of the ESC SpGEMM kernels (parity + config-4 timing, A/B against the round-1 hash
# kernels), clean config-3 SpMM timings (row-limit / ring sweep), one ncu capture of the new kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spgemm_esc.py tests/test_gpu_spgemm.py tests/test_gpu_jacobi.py tests/test_gpu_hostvec_defer.py -q -x \
    > gpurun_out/r02c2_pytest.log 2>&1; tail -n 6 gpurun_out/r02c2_pytest.log
timeout 400 python tools/bench_spgemm.py --reps 3 --out gpurun_out/r02c2_spgemm_esc.json > gpurun_out/r02c2_spgemm_esc.log 2>&1; tail -n 4 gpurun_out/r02c2_spgemm_esc.log
B200SP_SPGEMM_SYMBOLIC=1 B200SP_SPGEMM_NUMERIC=1 timeout 400 python tools/bench_spgemm.py --reps 2 --out gpurun_out/r02c2_spgemm_hash.json \
    > gpurun_out/r02c2_spgemm_hash.log 2>&1; tail -n 3 gpurun_out/r02c2_spgemm_hash.log
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:esc_(sym|num)_kernel' -c 2 -f -o gpurun_out/r02c2_esc \
    python tools/bench_spgemm.py --n 400000 --reps 1 --out gpurun_out/scratch.json > gpurun_out/r02c2_ncu_esc.log 2>&1
timeout 500 python tools/bench_spmm.py --out gpurun_out/r02c2_spmm_default.json > gpurun_out/r02c2_spmm_default.log 2>&1; tail -n 3 gpurun_out/r02c2_spmm_default.log
for lm in 64 128; do
  B200SP_SPMM_LMAX=$lm timeout 500 python tools/bench_spmm.py --out gpurun_out/r02c2_spmm_lmax$lm.json > gpurun_out/r02c2_spmm_lmax$lm.log 2>&1; tail -n 3 gpurun_out/r02c2_spmm_lmax$lm.log
done
ls -la gpurun_out | tail -20
