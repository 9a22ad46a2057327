#!/bin/bash
# Round 2, GPU call 27 (single B200): the lane-group sorted spadd kernels -- parity (tests/test_gpu_crs_utils.py) and timing against the
# one-thread-per-row kernels (tools/bench_spadd.py).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c27
timeout 200 python -m pytest tests/test_gpu_crs_utils.py -x -q -m gpu > ${O}_pytest.log 2>&1; tail -n 3 ${O}_pytest.log
timeout 150 python tools/bench_spadd.py --grid 64 --out ${O}_spadd_64.json > ${O}_spadd_64.log 2>&1; tail -n 6 ${O}_spadd_64.log
