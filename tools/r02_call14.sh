#!/bin/bash
# Round 2, GPU call 14 (single B200): SpMM after the hub-row reduce (launch list), the suites it touches, bench.py.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c14
timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_spmv.py -q -x > ${O}_pytest.log 2>&1; tail -n 3 ${O}_pytest.log
timeout 300 python tools/bench_spmm.py --scale 23 --out ${O}_spmm.json > ${O}_spmm.log 2>&1; grep "Layout" ${O}_spmm.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${O}_spmm_launches.csv \
    python tools/bench_spmm.py --scale 23 --iters 2 --out gpurun_out/scratch.json > ${O}_spmm_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02c14_spmm_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:60]].append(float(r[-1]))
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us")
PY
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.log; tail -c 300 ${O}_bench_n1.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c14_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"].get("note","")[-90:])
for s in d["secondary"]:
    print(s["metric"], s.get("value"), s.get("ms"), s.get("ms_symbolic"), s.get("ms_numeric"), s.get("layout_left",{}).get("ms"))
PY
