#!/bin/bash
# Round 2, GPU call 19 (2 GPUs): every all-gather transport incl. multicast_fwd (final form) against the oracle, RowBlockSpGEMM on
# GPUs for the first time, nvidia-smi topology of the box.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02c19_topo.log 2>&1; head -6 gpurun_out/r02c19_topo.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
    tools/multigpu_check.py > gpurun_out/r02c19_mgpu_check.log 2>&1; grep -E "^OK|^SKIP|FAIL|Error|error|assert" gpurun_out/r02c19_mgpu_check.log | head -20
