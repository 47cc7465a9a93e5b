#!/bin/bash
# Round 2, call A: column-split chroma SATD (parity + A/B of the job order), DCT matrix through shared memory.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_me.py tests/test_gpu_x_frame_chroma.py tests/test_gpu_frame.py tests/test_gpu_testbench.py tests/test_gpu_batched.py tests/test_gpu_table.py -m gpu -q -x) > gpurun_out/tA.log 2>&1
tail -n 6 gpurun_out/tA.log | cut -c1-300
for o in 0 1 2; do
  X265CU_ME_ORDER=$o timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-primitives > gpurun_out/bench_c3_ord$o.json 2> gpurun_out/bench_c3_ord$o.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-primitives --no-chroma > gpurun_out/bench_c3_luma.json 2> gpurun_out/bench_c3_luma.err
timeout 400 python bench.py --config c4 --steps 3 --warmup 2 --cpu-seconds 12 --no-primitives > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for f in c3_ord0 c3_ord1 c3_ord2 c3_luma c4; do python - "$f" <<'P'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1])); print(sys.argv[1], round(d["value"]), "e2e", round(d["e2e"]["value"]), {k: round(v, 2) for k, v in d.get("stages_ms").items()}, d.get("checks_equal"))
except Exception as e:
    print(sys.argv[1], "failed", e)
P
done
tail -n 3 gpurun_out/bench_c4.err
timeout 300 python profiles/primitive_bench.py --only k_transform,k_quant --frames 24 --reps 5 --json gpurun_out/prims_transform24.json 2>&1 | tail -n 24
du -sh gpurun_out
