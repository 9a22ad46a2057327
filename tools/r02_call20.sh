#!/bin/bash
# Round 2, GPU call 20 (single B200): final rehearsal of the round-end run (whole `pytest -m gpu`, smoke(), both bench arms) with the
# final tree, the cost of tile forwarding on one GPU, and an ncu --set full capture of the forwarding form of the tile kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02c20
timeout 1500 python -m pytest tests/ -x -q -m gpu > ${O}_pytest_gpu.log 2>&1; tail -n 3 ${O}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -n 1 ${O}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_n1.json 2> ${O}_bench_n1.log; tail -c 200 ${O}_bench_n1.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c20_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
for s in d["secondary"]:
    print(s["metric"], s.get("value"), s.get("ms"), s.get("ms_symbolic"), s.get("ms_numeric"), s.get("layout_left",{}).get("ms"), s["roofline"].get("traffic"))
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > ${O}_bench_ref.json 2> ${O}_bench_ref.log; cut -c1-200 ${O}_bench_ref.json
timeout 300 python tools/bench_forward.py --out ${O}_forward.json > ${O}_forward.log 2>&1; tail -n 1 ${O}_forward.log | cut -c1-400
timeout 400 ncu --set full --import-source on --clock-control none -k 'regex:spmv_tile_kernel<.*true>' -s 3 -c 1 -f -o ${O}_spmv_tile_fwd \
    python tools/bench_forward.py --iters 2 --out gpurun_out/scratch.json > ${O}_ncu_fwd.log 2>&1; tail -n 2 ${O}_ncu_fwd.log | cut -c1-200
ls -la gpurun_out | tail -5
