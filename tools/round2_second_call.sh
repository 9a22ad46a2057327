#!/bin/bash
# Second GPU call of the next round (single B200): clean full-size timings and profiles.
#   1. every harness suite at full size (config 4 SpGEMM incl. numeric variants 4-6 / symbolic 2, config 3 SpMM incl. the row-limit
#      sweep, BsrMatrix bandwidths per block size, solvers at 2 M rows, and spmv64 --big: the 2.2e9-entry stacked stencil, 26 GB)
#   2. ncu: launch list + one full capture per kernel VERDICT is likely to name (never bench numbers)
#   gpurun --timeout 2400 -- 'bash tools/round2_second_call.sh'
set -u
mkdir -p gpurun_out
G=./kokkos-kernels_b200/lib/gpu_check
O=gpurun_out/r02_gpu_check_big.jsonl
L=gpurun_out/r02_gpu_check_big.log
: > $L
timeout 1200 $G --big --out $O >> $L 2>&1
timeout 300 $G --suite spmm --spmm-scale 23 --out $O >> $L 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --target-processes all -c 400 --csv \
    --log-file gpurun_out/r02_launches_spmm.csv $G --big --suite spmm --out gpurun_out/scratch.jsonl > /dev/null 2>&1
for k in spmm_seg_kernel spmm_tile_kernel num_hash_kernel sym_hash_kernel bsr_tile_e_kernel gs_set_kernel; do
  suite=spmm; [ "${k#num}" != "$k" ] && suite=spgemm_c4; [ "${k#sym}" != "$k" ] && suite=spgemm_c4; [ "${k#bsr}" != "$k" ] && suite=bsr
  [ "${k#gs_}" != "$k" ] && suite=solvers
  timeout 150 ncu --set full --import-source on --clock-control none --target-processes all -k regex:$k -c 1 -f \
      -o gpurun_out/r02_$k $G --suite $suite --out gpurun_out/scratch.jsonl > gpurun_out/r02_ncu_$k.log 2>&1
done
grep -E "FAIL|summary" $L | head -60
ls -la gpurun_out | tail -20
