#!/bin/bash
# Round 2, GPU call 28 (single B200, the round's last minutes): the lane-group utility kernels (sorted / unsorted spadd, sort_and_merge
# fill) and the pool-keeping scratch scope -- parity of the files that use them, then tools/bench_spadd.py.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c28
timeout 60 python -m pytest tests/test_gpu_crs_utils.py -x -q -m gpu > ${O}_pytest_crs.log 2>&1; tail -n 2 ${O}_pytest_crs.log
timeout 70 python tools/bench_spadd.py --grid 64 --out ${O}_spadd_64.json > ${O}_spadd_64.log 2>&1; tail -n 8 ${O}_spadd_64.log | cut -c1-250
timeout 60 python -m pytest tests/test_gpu_spgemm.py tests/test_gpu_gs.py tests/test_gpu_jacobi.py -x -q -m gpu > ${O}_pytest_more.log 2>&1; tail -n 2 ${O}_pytest_more.log
