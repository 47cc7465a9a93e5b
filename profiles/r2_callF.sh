#!/bin/bash
# Round 2, call F: per-op blockop instantiations; the ncu launch list of the bench command and the per-launch counters of the
# final ME kernels; c2 re-measured.
mkdir -p gpurun_out /tmp/nc
(time timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "batched or table or testbench or frame or predcost") > gpurun_out/tF.log 2>&1
tail -n 5 gpurun_out/tF.log | cut -c1-300
timeout 200 python profiles/primitive_bench.py --only k_blockop --frames 24 --reps 5 2>&1 | grep k_blockop
timeout 300 python bench.py --config c2 --steps 3 --warmup 2 --cpu-seconds 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print('c2', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d.get('checks_equal'), d['cpu_baseline']['value'])"
# launch list of the bench command (per-launch times under ncu are cold-cache and serialised: compare SHARES)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-primitives --no-pred > gpurun_out/launches_r2_bench.log 2>&1
tail -n 2 gpurun_out/launches_r2_bench.log | cut -c1-200
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_me -c 9 --csv --log-file gpurun_out/me_launches_r2f.csv python profiles/run_small.py 1920 1088 1 1 1 > gpurun_out/me_launches_r2f.log 2>&1
# one full capture of the (new) dominant chroma sub-pel kernel, raw page only
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_me_chroma -s 2 -c 1 -o /tmp/nc/mechroma_f python profiles/run_small.py 1920 1088 1 1 1 > gpurun_out/mechroma_f.log 2>&1
ncu -i /tmp/nc/mechroma_f.ncu-rep --page raw --csv > gpurun_out/mechroma_f_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv | tail -5; du -sh gpurun_out
