#!/bin/bash
# Round 2, call D: packed all-angs, split job-list pixelcmp kernels, full search, prediction-cost kernel tweaks.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/tD.log 2>&1
tail -n 8 gpurun_out/tD.log | cut -c1-400
timeout 300 python profiles/primitive_bench.py --only "k_pixelcmp sad,k_pixelcmp sse,k_pixelcmp satd 8,k_intra" --frames 24 --reps 5 --json gpurun_out/prims_D.json 2>&1 | grep -v "^batch" | awk '{print $1,$2,$3,$4,$5,$(NF-6),$(NF-5),$(NF-4),$(NF-3)}' | head -40
(time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-primitives > gpurun_out/bench_c3_D.json 2> gpurun_out/bench_c3_D.err) 2>&1 | grep real
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_c3_D.json"))
print("c3", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms"].items()})
p = d.get("pred_cost") or {}
print("pred", p.get("jobs"), p.get("ms"), p.get("jobs_per_s"), d.get("pred_cost_error"))
P
tail -n 3 gpurun_out/bench_c3_D.err
du -sh gpurun_out
