mkdir -p gpurun_out
(time timeout 200 python -m pytest tests/test_gpu_me.py tests/test_gpu_frame.py -x -q) > gpurun_out/t_v2.log 2>&1
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
X265CU_LIB=$PWD/ab/libx265cu_v1.so timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
tail -n 4 gpurun_out/t_v2.log
python - <<'P'
import json
for n in ("v2","v1"):
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%n)); print(n, d["value"], d["stages_ms"], d["checks"]["me_cost_sum"])
    except Exception as e: print(n, "failed", e)
P
