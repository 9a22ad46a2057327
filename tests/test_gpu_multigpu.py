"""Row-block SpMV over several GPUs (BASELINE.json configs[4] family, kokkos-kernels_b200/multigpu.py): every all-gather
transport against the host oracle, incl. chained steps over the two next-x buffers and the host-vector form.  Needs >= 2 GPUs
(skipped on a single-GPU box); the worker is tools/multigpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_block_spmv_transports(cuda):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    oks = [ln.split()[1] for ln in out.stdout.splitlines() if ln.startswith("OK ")]
    assert "nccl" in oks and "pipelined" in oks, out.stdout[-2000:]
