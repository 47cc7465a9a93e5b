#!/bin/bash
# Round 2 final single-GPU call: whole GPU test suite, the bench lines, ncu DRAM bytes for the ME phases and the primitives.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --maxfail=10) > gpurun_out/t_all.log 2>&1
tail -n 8 gpurun_out/t_all.log | cut -c1-300
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"
timeout 400 python bench.py --config c4 --steps 3 --warmup 2 --cpu-seconds 15 --no-primitives > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"
timeout 300 python bench.py --config c2 --steps 3 --warmup 2 --cpu-seconds 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-chroma --no-primitives > gpurun_out/bench_c3_luma.json 2> gpurun_out/bench_c3_luma.err
timeout 400 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu --no-primitives --no-pred > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"
for f in c3 c3_luma c4 c2 c5; do python - "$f" <<'P'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1])); print(sys.argv[1], round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("stages_ms"), d.get("checks_equal"), d.get("cpu_baseline", {}).get("value"), len(d.get("primitives", [])), d.get("primitives_error"))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
done
tail -n 3 gpurun_out/bench_c3.err gpurun_out/bench_c4.err gpurun_out/bench_c2.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_me -c 9 --csv --log-file gpurun_out/me_traffic_2160.csv python profiles/run_small.py 3840 2160 1 1 1 > gpurun_out/me_traffic_2160.log 2>&1
timeout 500 ncu --metrics $M --clock-control none -k "regex:^(void )?(lf::)?k_" --csv --log-file gpurun_out/prims_ncu.csv python profiles/primitive_bench.py --single --frames 24 --depth 8,10 --json gpurun_out/prims_single.json > gpurun_out/prims_ncu.log 2>&1
tail -n 3 gpurun_out/prims_ncu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
du -sh gpurun_out
