#!/bin/bash
# Round 2, GPU call 26 (single B200): bench.py as the driver runs it (the NVML clock sampler's first run on hardware), then the whole
# `pytest -m gpu` suite on the final tree.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c26
( time timeout 420 python bench.py > ${O}_bench_n1.json 2> ${O}_bench_n1.log ) 2> ${O}_bench_time.txt; tail -n 3 ${O}_bench_time.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c26_bench_n1.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "clocks", "gpu_launches")})
print("e2e", d.get("e2e", {}).get("value"), "roofline", d.get("roofline", {}).get("frac"))
PY
timeout 600 python -m pytest tests/ -x -q -m gpu > ${O}_pytest_gpu.log 2>&1; tail -n 3 ${O}_pytest_gpu.log
