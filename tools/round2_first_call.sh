#!/bin/bash
# First GPU call of the next round (single B200, nothing else on the GPU): clean timings of everything that was
# only measured under contention or not at all in round 1, plus ncu captures of the kernels VERDICT is likely to name.
#   gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (copy the summaries to profiles/ afterwards).
set -u
mkdir -p gpurun_out
G=./kokkos-kernels_b200/lib/gpu_check
O=gpurun_out/r02_gpu_check_first.jsonl
# 1. parity + clean timings of all harness suites at full size (config 4 SpGEMM, config 3 SpMM incl. the row-limit sweep)
$G --big --out $O > gpurun_out/r02_gpu_check_first.log 2>&1
$G --suite spmm --spmm-scale 23 --out $O >> gpurun_out/r02_gpu_check_first.log 2>&1
# (the `jacobi` and `bsr` suites of the --big run above are the first GPU runs of spgemm_jacobi and of the BsrMatrix kernels)
python -m pytest tests -x -q -m gpu_next > gpurun_out/r02_pytest_gpu_next.log 2>&1; tail -n 3 gpurun_out/r02_pytest_gpu_next.log
# -> on success: change `gpu_next` to `gpu` in tests/test_gpu_jacobi.py, tests/test_gpu_bsr.py, tests/test_shim.py
# 2. ncu: launch lists + one full capture per kernel of interest (never bench numbers)
ncu --metrics gpu__time_duration.sum --clock-control none --target-processes all -c 400 --csv \
    --log-file gpurun_out/r02_launches_spmm.csv $G --big --suite spmm --out gpurun_out/scratch.jsonl > /dev/null 2>&1
for k in spmm_seg_kernel spmm_tile_kernel num_hash_kernel sym_hash_kernel bsr_tile_kernel; do
  suite=spmm; [ "${k#num}" != "$k" ] && suite=spgemm_c4; [ "${k#sym}" != "$k" ] && suite=spgemm_c4; [ "${k#bsr}" != "$k" ] && suite=bsr
  timeout 120 ncu --set full --import-source on --clock-control none --target-processes all -k regex:$k -c 1 -f \
      -o gpurun_out/r02_$k $G --suite $suite --out gpurun_out/scratch.jsonl > gpurun_out/r02_ncu_$k.log 2>&1
done
# 3. the headline bench (SpMV config 2) for reference, same box
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
grep -E "FAIL|summary" gpurun_out/r02_gpu_check_first.log | head -40
tail -1 gpurun_out/r02_bench_n1.json | cut -c1-400
ls -la gpurun_out | tail -20
