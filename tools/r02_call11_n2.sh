#!/bin/bash
# Round 2, GPU call 11 (2 GPUs): first run of the tile-forwarding fused all-gather (multicast_fwd) -- every transport against the
# oracle, then bench.py --gpus 2 with the transport auto-selection.
set -u
mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    tools/multigpu_check.py > gpurun_out/r02c11_mgpu_check.log 2>&1; grep -E "^OK|^SKIP|Error|error|assert" gpurun_out/r02c11_mgpu_check.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02c11_bench_n$N.json 2> gpurun_out/r02c11_bench_n$N.log
grep -E "collective\]|Error|error|Traceback" gpurun_out/r02c11_bench_n$N.log | head; cut -c1-200 gpurun_out/r02c11_bench_n$N.json
tail -3 gpurun_out/r02c11_mgpu_check.log
