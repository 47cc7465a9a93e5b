#!/bin/bash
# Round 2, call B: sub-group job-list pixelcmp, one-launch border extension, coalesced inverse-transform stores, staged 8x8 TUs,
# warp-per-job intra filter: whole GPU suite + the primitive table at 24 frames + the c3 line.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8) > gpurun_out/tB.log 2>&1
tail -n 12 gpurun_out/tB.log | cut -c1-400
(time timeout 500 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_B.json 2> gpurun_out/bench_c3_B.err) 2>&1 | grep real
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_c3_B.json"))
print("c3", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d["stages_ms"].items()})
for p in sorted(d.get("primitives", []), key=lambda r: r["frac"]):
    print("%-44s %2d %7.1f GB/s %5.1f%%" % (p["kernel"], p["depth"], p["GBps"], 100 * p["frac"]))
print(d.get("primitives_error"))
P
tail -n 3 gpurun_out/bench_c3_B.err
du -sh gpurun_out
