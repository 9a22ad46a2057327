#!/bin/bash
# Round 2, GPU call 25 (single B200): the lane-group triangular solve (sptrsv.cu) -- parity tests, then its timing against call 24's
# one-thread-per-row kernel (11.2 / 19.5 ms at 96^3 / 160^3), group sizes 8 / 16 / 32.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c25
timeout 600 python -m pytest tests/test_gpu_sptrsv.py tests/test_gpu_gs2.py tests/test_shim.py -x -q -m gpu > ${O}_pytest.log 2>&1; tail -n 3 ${O}_pytest.log
timeout 200 python tools/bench_sptrsv.py --grid 96 --out ${O}_sptrsv_96.json > ${O}_sptrsv_96.log 2>&1; tail -n 3 ${O}_sptrsv_96.log
for g in 8 16 32; do
  B200SP_SPTRSV_GROUP=$g timeout 200 python tools/bench_sptrsv.py --grid 160 --iters 5 --out ${O}_sptrsv_160_g$g.json > ${O}_sptrsv_160_g$g.log 2>&1
  echo "group $g:"; tail -n 3 ${O}_sptrsv_160_g$g.log
done
