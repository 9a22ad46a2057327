#!/bin/bash
# Round 2, GPU call 13 (single B200): rehearsal of what the driver runs at round end -- the whole `pytest -m gpu` suite, smoke(),
# bench.py (own arm and --impl reference), plus the ncu launch list of bench.py for profiles/.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c13
timeout 1500 python -m pytest tests/ -x -q -m gpu > ${O}_pytest_gpu.log 2>&1; tail -n 4 ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -n 2 ${O}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.log; tail -c 400 ${O}_bench_n1.log; cut -c1-300 ${O}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > ${O}_bench_ref.json 2> ${O}_bench_ref.log; cut -c1-400 ${O}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_bench.csv \
    python bench.py --steps 3 --warmup 1 --no-cpu > ${O}_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02c13_launches_bench.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:70]].append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
    print(f"{k:72s} n={len(v):3d} avg={sum(v)/len(v)/1e3:9.1f} us share={100*sum(v)/tot:5.1f}%")
PY
ls -la gpurun_out | tail -6
