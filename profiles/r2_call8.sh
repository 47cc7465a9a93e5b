#!/bin/bash
mkdir -p gpurun_out /tmp/nc
(time timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "frame or window or me_batch or chroma") > gpurun_out/t_frame.log 2>&1
tail -n 6 gpurun_out/t_frame.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 15 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-chroma > gpurun_out/bench_c3_luma.json 2> gpurun_out/bench_c3_luma.err
for f in c3 c3_luma; do python - "$f" <<'P'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1])); print(sys.argv[1], round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("stages_ms"), d.get("checks_equal"), d.get("cpu_baseline", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
done
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio
timeout 300 ncu --metrics $M --clock-control none -k regex:k_me -c 9 --csv --log-file gpurun_out/me_launches_r2e.csv python profiles/run_small.py 1920 1088 1 1 1 > gpurun_out/me_launches_r2e.log 2>&1
