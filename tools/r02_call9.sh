#!/bin/bash
# Round 2, GPU call 9 (single B200): ESC kernels with the cp.async row pipeline (A/B: one CTA per row, register budgets), parity on
# hardware, per-kernel launch times of the SpMM call, ncu of the ESC kernels, then bench.py as the driver runs it.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c9
timeout 1200 python -m pytest tests/test_gpu_spgemm_esc.py tests/test_gpu_spgemm.py tests/test_gpu_jacobi.py tests/test_gpu_bsr.py -q -x > ${O}_pytest.log 2>&1; tail -n 3 ${O}_pytest.log
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${O}_spmm_launches.csv \
    python tools/bench_spmm.py --scale 23 --iters 2 --out gpurun_out/scratch.json > ${O}_spmm_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02c9_spmm_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:60]].append(float(r[-1]))
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us")
PY
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:esc_(sym|num)_kernel' -c 2 -f -o ${O}_esc \
    python tools/bench_spgemm.py --reps 1 --out gpurun_out/scratch.json > ${O}_ncu_esc.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.log; tail -c 600 ${O}_bench_n1.log; cut -c1-400 ${O}_bench_n1.json
ls -la gpurun_out | tail -8
