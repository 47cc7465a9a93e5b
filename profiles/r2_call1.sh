#!/bin/bash
# Round 2, GPU call 1: parity of the shared-memory-window integer search + the new bench line + a first ncu look.
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -x -k "frame or window") > gpurun_out/t_frame.log 2>&1
tail -n 25 gpurun_out/t_frame.log
(time timeout 600 python -m pytest tests -m gpu -q --maxfail=12 -k "not frame and not window") > gpurun_out/t_rest.log 2>&1
tail -n 6 gpurun_out/t_rest.log
timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 15 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
X265CU_ME_WINDOW=0 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_nowin.json 2> gpurun_out/bench_c3_nowin.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-chroma > gpurun_out/bench_c3_luma.json 2> gpurun_out/bench_c3_luma.err
for f in c3 c3_nowin c3_luma; do python - "$f" <<'P'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1])); print(sys.argv[1], round(d["value"]), "e2e", round(d["e2e"]["value"]), d["stages_ms"], d.get("checks_equal"))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
done
tail -n 5 gpurun_out/bench_c3.err
M=gpu__time_duration.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem
timeout 300 ncu --metrics $M --clock-control none -k regex:k_me -c 16 --csv --log-file gpurun_out/me_launches_r2.csv python profiles/run_small.py 1920 1088 1 2 > gpurun_out/me_launches_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_me_window -s 2 -c 2 -o gpurun_out/mew_r2 python profiles/run_small.py 1920 1088 1 2 > gpurun_out/mew_ncu.log 2>&1
ls -la gpurun_out | tail -n 12
