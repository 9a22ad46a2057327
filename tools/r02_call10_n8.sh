#!/bin/bash
# Round 2, GPU call 10 (8 GPUs, charged 8x: one run only): bench.py at N = 8 as the driver launches it (transport auto-selection
# incl. NVSwitch multicast, sliced end-to-end leg, bit-compare of every rank's gathered vector).
set -u
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29548 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02c10_bench_n$N.json 2> gpurun_out/r02c10_bench_n$N.log
grep -E "collective\]|Error|error|Traceback" gpurun_out/r02c10_bench_n$N.log | head -8; cut -c1-300 gpurun_out/r02c10_bench_n$N.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c10_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), json.dumps(d.get("collective")))
except Exception as e:
    print("no result:", e)
PY
tail -4 gpurun_out/r02c10_bench_n$N.log
