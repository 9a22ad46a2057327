#!/bin/bash
# Round 2, GPU call 24 (single B200): the whole `pytest -m gpu` suite on the tree with the SPTRSV shim specialisations and the chained
# small levels of the triangular solve; first timing of sptrsv (tools/bench_sptrsv.py).
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c24
timeout 900 python -m pytest tests/ -x -q -m gpu > ${O}_pytest_gpu.log 2>&1; tail -n 3 ${O}_pytest_gpu.log
timeout 240 python tools/bench_sptrsv.py --grid 96 --out ${O}_sptrsv_96.json > ${O}_sptrsv_96.log 2>&1; tail -n 4 ${O}_sptrsv_96.log
timeout 240 python tools/bench_sptrsv.py --grid 160 --iters 5 --out ${O}_sptrsv_160.json > ${O}_sptrsv_160.log 2>&1; tail -n 4 ${O}_sptrsv_160.log
