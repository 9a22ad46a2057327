#!/bin/bash
# Round 2, GPU call 22 (8 GPUs, charged 8x): bench.py at N = 8 with the software-pipelined tile forwarding.
set -u
mkdir -p gpurun_out
N=8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02c22_bench_n$N.json 2> gpurun_out/r02c22_bench_n$N.log
grep -E "collective\]|REJECTED|Error|error|Traceback" gpurun_out/r02c22_bench_n$N.log | head -8
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c22_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), json.dumps(d.get("collective"))[:500])
except Exception as e:
    print("no result:", e)
PY
