#!/bin/bash
# Round 2, GPU call 23 (single B200): the whole `pytest -m gpu` suite with the sptrsv / classic Gauss-Seidel additions, smoke().
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c23
timeout 1500 python -m pytest tests/ -x -q -m gpu > ${O}_pytest_gpu.log 2>&1; tail -n 3 ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -n 1 ${O}_smoke.log
timeout 300 python -m pytest tests/test_gpu_sptrsv.py -q --durations=5 > ${O}_pytest_sptrsv.log 2>&1; tail -n 8 ${O}_pytest_sptrsv.log
