#!/bin/bash
# Round 2, GPU call 17 (single B200): SpMM item kernel with / without the L2 evict-last hint on the gathers of X.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c17
for k in 0 1 0 1; do
  B200SP_SPMM_X_KEEP=$k timeout 300 python tools/bench_spmm.py --scale 23 --iters 20 --out ${O}_spmm_keep$k.json > ${O}_spmm_keep$k.log 2>&1
  echo "keep=$k: $(grep "LayoutRight" ${O}_spmm_keep$k.log | cut -c1-160)"
done
timeout 600 python -m pytest tests/test_gpu_spmm.py -q -x > ${O}_pytest.log 2>&1; tail -n 2 ${O}_pytest.log
