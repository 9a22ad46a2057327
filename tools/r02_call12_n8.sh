#!/bin/bash
# Round 2, GPU call 12 (8 GPUs, charged 8x): bench.py at N = 8 with the transport auto-selection (now incl. multicast_fwd), then the
# piece pipeline with more push CTAs / fewer pieces.
set -u
mkdir -p gpurun_out
N=8
run() {  # name, extra args
  local name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$((RANDOM % 10)) \
      bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/r02c12_bench_n${N}_$name.json 2> gpurun_out/r02c12_bench_n${N}_$name.log
  grep -E "collective\]|Error|error|Traceback" gpurun_out/r02c12_bench_n${N}_$name.log | head -4
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02c12_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), json.dumps(d.get("collective"))[:400])
except Exception as e:
    print("$name: no result:", e)
PY
}
run auto
run mc128 --collective pipelined_mc --push-ctas 128 --no-cpu
run mc64c4 --collective pipelined_mc --push-ctas 64 --chunks 4 --no-cpu
