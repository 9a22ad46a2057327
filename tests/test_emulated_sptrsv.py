"""tests/test_gpu_sptrsv.py run on the CPU: the level-set triangular solve and the classic two-stage Gauss-Seidel kernels executed
under the CUDA-on-CPU emulation (tools/emu, TEST INFRASTRUCTURE) against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_sptrsv_suite_under_emulation():
    env = dict(os.environ, B200SP_TEST_EMULATED="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_sptrsv.py"), "-m", "gpu", "-x", "-q"],
                         capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-500:]
