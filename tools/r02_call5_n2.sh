#!/bin/bash
# Round 2, GPU call 5 (2 GPUs): first run of the reworked multi-GPU path -- every all-gather transport against the oracle
# (tools/multigpu_check.py), then bench.py --gpus 2 with the transport auto-selection and the sliced end-to-end leg, then the
# reference arm as the driver launches it (under torchrun: only rank 0 works, with all host threads).
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    tools/multigpu_check.py > gpurun_out/r02c5_mgpu_check.log 2>&1; grep -E "^OK|^SKIP|Error|error|assert" gpurun_out/r02c5_mgpu_check.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02c5_bench_n$N.json 2> gpurun_out/r02c5_bench_n$N.log
grep -E "collective|Error|error" gpurun_out/r02c5_bench_n$N.log | head; cut -c1-2500 gpurun_out/r02c5_bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/r02c5_bench_ref_n$N.json 2> gpurun_out/r02c5_bench_ref_n$N.log
cut -c1-1200 gpurun_out/r02c5_bench_ref_n$N.json
tail -5 gpurun_out/r02c5_mgpu_check.log
