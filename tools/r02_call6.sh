#!/bin/bash
# Round 2, GPU call 6 (single B200): symbolic with the bitmap pre-filter (phase trace), numeric configurations 3..7, ncu of both
# at the full config-4 size, full-size config-3 parity, BsrMatrix bandwidths at full size (default vs vector kernel).
set -u
mkdir -p gpurun_out
B200SP_SPGEMM_TRACE=1 B200SP_ESC_CFG=3 timeout 300 python tools/bench_spgemm.py --reps 3 --out gpurun_out/r02c6_spgemm_trace.json > gpurun_out/r02c6_spgemm_trace.log 2>&1
grep -E "spgemm_symbolic\]|'rep'" gpurun_out/r02c6_spgemm_trace.log | cut -c1-150 | tail -24
for cfg in 3 4 5 6 7; do
  B200SP_ESC_CFG=$cfg timeout 300 python tools/bench_spgemm.py --reps 3 --out gpurun_out/r02c6_spgemm_cfg$cfg.json > gpurun_out/r02c6_spgemm_cfg$cfg.log 2>&1
  echo "cfg $cfg: $(grep "'rep': 2" gpurun_out/r02c6_spgemm_cfg$cfg.log | cut -c1-120)"
done
B200SP_ESC_CFG=3 timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:esc_(sym|num)_kernel' -c 2 -f -o gpurun_out/r02c6_esc \
    python tools/bench_spgemm.py --reps 1 --out gpurun_out/scratch.json > gpurun_out/r02c6_ncu_esc.log 2>&1
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_spgemm_esc.py -q -x -k "config3_full_size or esc_cases or run_to_run" > gpurun_out/r02c6_pytest.log 2>&1; tail -n 4 gpurun_out/r02c6_pytest.log
G=./kokkos-kernels_b200/lib/gpu_check
timeout 400 $G --big --suite bsr --out gpurun_out/r02c6_bsr_big.jsonl > gpurun_out/r02c6_bsr_big_default.log 2>&1
B200SP_BSR_KERNEL=vector timeout 400 $G --big --suite bsr --out gpurun_out/r02c6_bsr_big_vector.jsonl > gpurun_out/r02c6_bsr_big_vector.log 2>&1
grep -E "N_default|N_vector|N_walk" gpurun_out/r02c6_bsr_big_default.log | cut -c1-220 | head -30
ls -la gpurun_out | tail -8
