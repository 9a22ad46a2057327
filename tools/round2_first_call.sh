#!/bin/bash
# First GPU call of the next round (single B200, nothing else on the GPU), most valuable first:
#   1. first GPU run of everything written after round 1's GPU minutes were spent (validated under the CPU emulation only):
#      harness suites bsr / cg / jacobi / solvers / spmv64 (with --big: a matrix past 2^31 entries), then `pytest -m gpu_next` (BsrMatrix SpMV/SpMM, CG / PCG, GMRES, Gauss-Seidel, spgemm_jacobi, shim --bsr --jacobi)
#   2. clean timings of all harness suites at full size (config 4 SpGEMM incl. numeric variants 4-6 / symbolic 2, config 3 SpMM
#      incl. the row-limit sweep), which round 1 only measured under contention or not at all
#   3. ncu: launch list + one full capture per kernel VERDICT is likely to name (never bench numbers)
#   4. the headline bench (SpMV config 2) on the same box
#   gpurun --timeout 1200 -- 'bash tools/round2_first_call.sh'
# Outputs land in gpurun_out/ (copy the summaries to profiles/ afterwards).  On success of step 1 change `gpu_next` to `gpu`
# in tests/test_gpu_{jacobi,bsr,cg,gmres,gs,gs2,spmv64}.py and tests/test_shim.py.
set -u
mkdir -p gpurun_out
G=./kokkos-kernels_b200/lib/gpu_check
O=gpurun_out/r02_gpu_check_first.jsonl
L=gpurun_out/r02_gpu_check_first.log
: > $L
for s in bsr cg jacobi solvers spmv64; do $G --suite $s --out $O >> $L 2>&1; done
python -m pytest tests -x -q -m gpu_next > gpurun_out/r02_pytest_gpu_next.log 2>&1; tail -n 3 gpurun_out/r02_pytest_gpu_next.log
$G --big --out $O >> $L 2>&1
$G --suite spmm --spmm-scale 23 --out $O >> $L 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --target-processes all -c 400 --csv \
    --log-file gpurun_out/r02_launches_spmm.csv $G --big --suite spmm --out gpurun_out/scratch.jsonl > /dev/null 2>&1
for k in spmm_seg_kernel spmm_tile_kernel num_hash_kernel sym_hash_kernel bsr_tile_e_kernel; do
  suite=spmm; [ "${k#num}" != "$k" ] && suite=spgemm_c4; [ "${k#sym}" != "$k" ] && suite=spgemm_c4; [ "${k#bsr}" != "$k" ] && suite=bsr
  timeout 120 ncu --set full --import-source on --clock-control none --target-processes all -k regex:$k -c 1 -f \
      -o gpurun_out/r02_$k $G --suite $suite --out gpurun_out/scratch.jsonl > gpurun_out/r02_ncu_$k.log 2>&1
done
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
grep -E "FAIL|summary" $L | head -60
tail -1 gpurun_out/r02_bench_n1.json | cut -c1-400
ls -la gpurun_out | tail -20
