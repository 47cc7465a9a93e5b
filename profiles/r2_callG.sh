#!/bin/bash
# Round 2, call G: two-list prediction costs of small PUs on the lane-per-row-segment code (parity + timing); where the c2
# end-to-end step spends its host time.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_predcost.py tests/test_gpu_me.py tests/test_gpu_lookahead.py -m gpu -q -x) > gpurun_out/tG.log 2>&1
tail -n 5 gpurun_out/tG.log | cut -c1-400
timeout 200 python profiles/la_e2e_probe2.py 2>&1 | tail -6
timeout 300 python bench.py --config c2 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_c2_G.json 2> gpurun_out/bench_c2_G.err; python -c "import json; d=json.load(open('gpurun_out/bench_c2_G.json')); print('c2', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])"
(time timeout 600 python bench.py --steps 3 --warmup 3 --cpu-seconds 8 --no-primitives > gpurun_out/bench_c3_G.json 2> gpurun_out/bench_c3_G.err) 2>&1 | grep real
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_c3_G.json"))
print("c3", round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("checks_equal"))
p = d.get("pred_cost") or {}
print("pred", p.get("jobs"), p.get("ms"), p.get("jobs_per_s"), p.get("checks_equal"), d.get("pred_cost_error"))
P
tail -n 3 gpurun_out/bench_c3_G.err
