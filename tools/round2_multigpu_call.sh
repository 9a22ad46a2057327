#!/bin/bash
# Round-2 multi-GPU call: A/B the all-gather strategies of the row-block SpMV (config 5 family) on N GPUs.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_multigpu_call.sh 8'
# (charged N x box time: start with N=2 to see that multicast / pipelined_mc run at all, then N=8.)
set -u
N=${1:-8}
mkdir -p gpurun_out
for mode in pipelined multicast pipelined_mc fused nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --collective $mode \
      > gpurun_out/r02_bench_n${N}_${mode}.json 2> gpurun_out/r02_bench_n${N}_${mode}.log
  echo "$mode: $(tail -1 gpurun_out/r02_bench_n${N}_${mode}.json | python -c 'import sys,json; d=json.loads(sys.stdin.read() or "{}"); print(d.get("value"), "GFLOP/s", d.get("ms_per_step"), "ms/step", d.get("config",{}).get("collective"))' 2>/dev/null)"
done
