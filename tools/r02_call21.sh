#!/bin/bash
# Round 2, GPU call 21 (single B200): cost of the (now software-pipelined) tile forwarding on one GPU; SpMV suite; headline kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c21
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_hostvec_defer.py -q -x > ${O}_pytest.log 2>&1; tail -n 2 ${O}_pytest.log
for i in 1 2; do timeout 300 python tools/bench_forward.py --out ${O}_forward$i.json > ${O}_forward$i.log 2>&1; tail -n 1 ${O}_forward$i.log | cut -c1-420; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > ${O}_bench_n1.json 2> ${O}_bench_n1.log; cut -c1-200 ${O}_bench_n1.json
