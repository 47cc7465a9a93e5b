#!/bin/bash
# profiles only: capture on the box, convert to CSV there, bring back compressed CSV (the .ncu-rep files are too big)
mkdir -p gpurun_out /tmp/nc
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_me -c 9 --csv --log-file gpurun_out/me_launches_r2d.csv python profiles/run_small.py 1920 1088 1 1 1 > gpurun_out/me_launches_r2d.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_me_chroma -c 4 -o /tmp/nc/mechroma python profiles/run_small.py 1920 1088 1 1 1 > gpurun_out/mechroma_ncu.log 2>&1
ncu -i /tmp/nc/mechroma.ncu-rep --page raw --csv > gpurun_out/mechroma_raw.csv 2>/dev/null
for i in 0 1 2 3; do ncu -i /tmp/nc/mechroma.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/mechroma_src$i.csv.gz; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_me_window -c 3 -o /tmp/nc/mew python profiles/run_small.py 1920 1088 1 1 > gpurun_out/mew_ncu.log 2>&1
ncu -i /tmp/nc/mew.ncu-rep --page raw --csv > gpurun_out/mew_raw.csv 2>/dev/null
for i in 0 1 2; do ncu -i /tmp/nc/mew.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/mew_src$i.csv.gz; done
ls -la gpurun_out | tail -n 14; du -sh gpurun_out
