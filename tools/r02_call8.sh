#!/bin/bash
# Round 2, GPU call 8 (single B200): persistent ESC kernels with the row pipeline (A/B against one CTA per row), register-budget
# variants; cooperative SpMM item kernel (batch 8 / 4 / first kernel); parity of both on hardware; ncu of the three kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c8
timeout 1200 python -m pytest tests/test_gpu_spgemm_esc.py tests/test_gpu_spgemm.py tests/test_gpu_spmm.py tests/test_gpu_bsr.py tests/test_gpu_jacobi.py -q -x > ${O}_pytest.log 2>&1; tail -n 3 ${O}_pytest.log
run_spgemm() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python tools/bench_spgemm.py --reps 3 --out ${O}_spgemm_$name.json > ${O}_spgemm_$name.log 2>&1
  echo "spgemm $name: $(grep "'rep': 2" ${O}_spgemm_$name.log | cut -c1-140)"
}
run_spgemm default B200SP_SPGEMM_TRACE=1
grep -E "spgemm_symbolic\]" ${O}_spgemm_default.log | tail -5
run_spgemm oneperrow B200SP_ESC_PERSIST=0
run_spgemm sym1_num4 B200SP_ESC_SYM_CFG=1 B200SP_ESC_CFG=4
run_spgemm sym2_num7 B200SP_ESC_SYM_CFG=2 B200SP_ESC_CFG=7
run_spgemm sym4_num1 B200SP_ESC_SYM_CFG=4 B200SP_ESC_CFG=1
for c in 8 4 0; do
  B200SP_SPMM_ITEM_COOP=$c timeout 300 python tools/bench_spmm.py --scale 23 --out ${O}_spmm_coop$c.json > ${O}_spmm_coop$c.log 2>&1
  echo "spmm coop=$c: $(grep "LayoutRight" ${O}_spmm_coop$c.log | cut -c1-200)"
done
timeout 300 python tools/bench_bsr_mm.py --out ${O}_bsr_mm.json > ${O}_bsr_mm.log 2>&1; grep -E "'bs'|max \|" ${O}_bsr_mm.log | cut -c1-230
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:spmm_item_coop_kernel' -c 1 -f -o ${O}_spmm_coop \
    python tools/bench_spmm.py --scale 23 --out gpurun_out/scratch.json > ${O}_ncu_spmm.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:esc_(sym|num)_kernel' -c 2 -f -o ${O}_esc \
    python tools/bench_spgemm.py --reps 1 --out gpurun_out/scratch.json > ${O}_ncu_esc.log 2>&1
ls -la gpurun_out | tail -12
