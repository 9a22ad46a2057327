#!/bin/bash
# Round 2, GPU call 15 (single B200): after the last kernel edits (forwarding as a template parameter of the tile kernel, 8-CTA
# register budget of the symbolic kernel): the suites they touch, SpGEMM timing, fresh ncu --set full captures of the headline
# SpMV kernel and of the SpMM item kernel at full size (DRAM traffic for bench.py's roofline.traffic), bench.py.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c15
timeout 900 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_spgemm_esc.py tests/test_gpu_spgemm.py tests/test_gpu_hostvec_defer.py tests/test_shim.py -q -x > ${O}_pytest.log 2>&1; tail -n 3 ${O}_pytest.log
B200SP_SPGEMM_TRACE=1 timeout 300 python tools/bench_spgemm.py --reps 3 --out ${O}_spgemm.json > ${O}_spgemm.log 2>&1
grep -E "spgemm_symbolic\]" ${O}_spgemm.log | tail -5; echo "spgemm: $(grep "'rep': 2" ${O}_spgemm.log | cut -c1-140)"
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:spmv_tile_kernel' -s 3 -c 1 -f -o ${O}_spmv_tile \
    python bench.py --steps 3 --warmup 1 --no-cpu --no-secondary > ${O}_ncu_spmv.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:spmm_item_coop_kernel' -c 1 -f -o ${O}_spmm_coop \
    python tools/bench_spmm.py --scale 23 --iters 2 --out gpurun_out/scratch.json > ${O}_ncu_spmm.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.log; tail -c 300 ${O}_bench_n1.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c15_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
for s in d["secondary"]:
    print(s["metric"], s.get("value"), s.get("ms"), s.get("ms_symbolic"), s.get("ms_numeric"), s.get("layout_left",{}).get("ms"))
PY
ls -la gpurun_out | tail -6
