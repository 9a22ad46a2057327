#!/bin/bash
# After tools/round2_first_call.sh has passed on a B200 (gpurun_out/r02_pytest_gpu_next.log ends in "N passed", no FAIL in
# gpurun_out/r02_gpu_check_first.log): turn the `gpu_next` tests into regular `gpu` tests.  Files whose tests failed stay as they are --
# pass the ones to promote, or nothing for all of them.
set -eu
cd "$(dirname "$0")/.."
files=("$@")
if [ ${#files[@]} -eq 0 ]; then
  files=(tests/test_gpu_jacobi.py tests/test_gpu_bsr.py tests/test_gpu_cg.py tests/test_gpu_gmres.py tests/test_gpu_gs.py tests/test_gpu_gs2.py
         tests/test_gpu_spmv64.py tests/test_gpu_hostvec_defer.py tests/test_shim.py)
fi
for f in "${files[@]}"; do
  sed -i 's/pytest\.mark\.gpu_next/pytest.mark.gpu/g' "$f"
  echo "promoted $f"
done
grep -rn "gpu_next" tests/*.py | grep -v conftest.py || echo "no gpu_next markers left"
