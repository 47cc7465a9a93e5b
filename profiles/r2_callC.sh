#!/bin/bash
# Round 2, call C: job-list pixelcmp (pow2 paths, chunk heuristic), 16-bit lowres init, warp-per-row border extension,
# prediction-cost leg of the bench line (GPU vs the reference's Predict / MotionEstimate on a CPU sample).
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/tC.log 2>&1
tail -n 8 gpurun_out/tC.log | cut -c1-400
timeout 300 python profiles/primitive_bench.py --only k_pixelcmp,k_lowres,k_intra_filter --frames 24 --reps 5 --json gpurun_out/prims_C.json 2>&1 | grep -v "^batch" | awk '{print $1,$2,$3,$4,$5,$(NF-6),$(NF-5),$(NF-4),$(NF-3)}' | head -80
(time timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 12 --no-primitives > gpurun_out/bench_c3_C.json 2> gpurun_out/bench_c3_C.err) 2>&1 | grep real
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_c3_C.json"))
print("c3", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms"].items()}, d.get("checks_equal"))
print(json.dumps(d.get("pred_cost"), indent=None)[:1500]); print(d.get("pred_cost_error"))
P
tail -n 3 gpurun_out/bench_c3_C.err
du -sh gpurun_out
