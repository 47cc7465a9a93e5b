#!/bin/bash
# First GPU call of the next round (one B200, ~3 min): run it as  gpurun --timeout 420 -- 'bash profiles/next_round_first_call.sh'
# 1. the two parity tests written after round 1's GPU minutes were spent (non-strict xfail until they have run once):
#    weights analysis (host logic over verified kernels) and the analyser's chroma stage (glue over the verified k_me_chroma)
# 2. the bench line in its three modes (default / chroma-SATD on / CTU-row shards at N=1)
# 3. the l1tex breakdown of the integer-search kernel: which stage of the L1 data path saturates
mkdir -p gpurun_out
(time timeout 120 python -m pytest tests/test_gpu_weights.py tests/test_gpu_x_frame_chroma.py -q -rxX --runxfail) > gpurun_out/pending_tests.log 2>&1
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 120 python bench.py --chroma --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_chroma.json 2> gpurun_out/bench_chroma.err
timeout 120 python bench.py --shard rows --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_rows.json 2> gpurun_out/bench_rows.err
M=l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,l1tex__m_xbar2l1tex_read_sectors.sum,l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
timeout 200 ncu --metrics $M --clock-control none -k regex:k_me -c 12 --csv --log-file gpurun_out/me_l1tex.csv python profiles/run_small.py 1024 576 1 2 > gpurun_out/me_l1tex.log 2>&1
/usr/local/cuda/bin/nvcc -arch=sm_100a -O3 -o /tmp/l1_probe profiles/l1_probe.cu && timeout 60 /tmp/l1_probe > gpurun_out/l1_probe.txt 2>&1
# A/B of the per-burst segment width: variant library built on the box (~80 s), parity first, then the bench
/usr/local/cuda/bin/nvcc -DME_SEG_PER_BURST -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -o /tmp/libx265cu_seg.so x265_b200/csrc/x265cu.cu > gpurun_out/seg_build.log 2>&1 \
  && X265CU_LIB=/tmp/libx265cu_seg.so timeout 120 python -m pytest tests/test_gpu_me.py tests/test_gpu_frame.py tests/test_gpu_lookahead.py -x -q > gpurun_out/seg_tests.log 2>&1 \
  && X265CU_LIB=/tmp/libx265cu_seg.so timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_seg.json 2> gpurun_out/bench_seg.err
tail -n 3 gpurun_out/seg_tests.log
tail -n 8 gpurun_out/pending_tests.log
for f in default seg chroma rows; do python - "$f" <<'P'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1])); print(sys.argv[1], round(d["value"]), d["stages_ms"])
except Exception as e:
    print(sys.argv[1], "failed", e)
P
done
