#!/bin/bash
# Round 2, GPU call 3 (single B200): ESC v2 (symbolic by bucket sort, fixed-unroll ranks, 16-byte (column,value) pairs, block-aggregated
# binning), the item-based SpMM kernel (first run), and the new bench.py end to end.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spgemm_esc.py tests/test_gpu_spgemm.py tests/test_gpu_spmm.py tests/test_gpu_jacobi.py tests/test_gpu_spmv64.py -q -x \
    > gpurun_out/r02c3_pytest.log 2>&1; tail -n 6 gpurun_out/r02c3_pytest.log
timeout 400 python tools/bench_spgemm.py --reps 3 --out gpurun_out/r02c3_spgemm_esc.json > gpurun_out/r02c3_spgemm_esc.log 2>&1; tail -n 4 gpurun_out/r02c3_spgemm_esc.log
timeout 500 python tools/bench_spmm.py --out gpurun_out/r02c3_spmm_items.json > gpurun_out/r02c3_spmm_items.log 2>&1; tail -n 3 gpurun_out/r02c3_spmm_items.log
for lm in 32 64 256; do
  B200SP_SPMM_ITEM_LMAX=$lm timeout 500 python tools/bench_spmm.py --out gpurun_out/r02c3_spmm_items_lmax$lm.json > gpurun_out/r02c3_spmm_items_lmax$lm.log 2>&1; tail -n 3 gpurun_out/r02c3_spmm_items_lmax$lm.log
done
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:esc_(sym|num)_kernel' -c 2 -f -o gpurun_out/r02c3_esc \
    python tools/bench_spgemm.py --n 400000 --reps 1 --out gpurun_out/scratch.json > gpurun_out/r02c3_ncu_esc.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k 'regex:spmm_item_kernel' -c 1 -f -o gpurun_out/r02c3_spmm_item \
    python tools/bench_spmm.py --scale 22 --iters 2 --out gpurun_out/scratch.json > gpurun_out/r02c3_ncu_spmm.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c3_bench_n1.json 2> gpurun_out/r02c3_bench_n1.log; tail -n 5 gpurun_out/r02c3_bench_n1.log; cut -c1-1500 gpurun_out/r02c3_bench_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02c3_bench_ref.json 2> gpurun_out/r02c3_bench_ref.log; cut -c1-900 gpurun_out/r02c3_bench_ref.json
ls -la gpurun_out | tail -24
