#!/bin/bash
# Round 2, GPU call 7 (8 GPUs): bench.py at N = 8 and N = 4 as the driver launches it (transport auto-selection, sliced end-to-end leg)
set -u
mkdir -p gpurun_out
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N \
      bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02c7_bench_n$N.json 2> gpurun_out/r02c7_bench_n$N.log
  grep -E "collective\]|Error|error|Traceback" gpurun_out/r02c7_bench_n$N.log | head -5; cut -c1-200 gpurun_out/r02c7_bench_n$N.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c7_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], json.dumps(d.get("collective")))
except Exception as e:
    print("no result:", e)
PY
done
tail -3 gpurun_out/r02c7_bench_n8.log
