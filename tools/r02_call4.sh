#!/bin/bash
# Round 2, GPU call 4 (single B200): ESC configuration A/B on config 4, SpMM item kernel with L2 evict-first hints, the whole
# `-m gpu` suite, reference arm with the thread-count fix.
set -u
mkdir -p gpurun_out
for cfg in 0 1 2 3; do
  B200SP_ESC_CFG=$cfg timeout 300 python tools/bench_spgemm.py --reps 3 --out gpurun_out/r02c4_spgemm_cfg$cfg.json > gpurun_out/r02c4_spgemm_cfg$cfg.log 2>&1
  echo "cfg $cfg: $(grep "'rep': 2" gpurun_out/r02c4_spgemm_cfg$cfg.log | cut -c1-140)"
done
timeout 500 python tools/bench_spmm.py --out gpurun_out/r02c4_spmm_items.json > gpurun_out/r02c4_spmm_items.log 2>&1; tail -n 3 gpurun_out/r02c4_spmm_items.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02c4_bench_ref.json 2> gpurun_out/r02c4_bench_ref.log; cut -c1-1200 gpurun_out/r02c4_bench_ref.json
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02c4_pytest_gpu.log 2>&1; tail -n 8 gpurun_out/r02c4_pytest_gpu.log
ls -la gpurun_out | tail -12
