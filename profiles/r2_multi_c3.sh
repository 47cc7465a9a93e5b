#!/bin/bash
# Round 2, final multi-GPU line: bash profiles/r2_multi_c3.sh N   (under gpurun --gpus N) -- c3 only: frame shard (weak) with the
# CTU-row shard (strong) attached; per-step barrier before every timed step (rank skew no longer lands in the step time).
N=${1:-2}
mkdir -p gpurun_out
P=$((29600 + N))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale2_c3_n$N.json 2> gpurun_out/scale2_c3_n$N.err; echo "c3 N=$N rc=$?"
python - gpurun_out/scale2_c3_n$N.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], d["n_gpus"], round(d["value"]), d["unit"], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
                                            "rows:", (round(d["strong_scaling_rows"]["value"]), round(d["strong_scaling_rows"]["ms_per_step"], 2)) if "strong_scaling_rows" in d else None)
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
tail -n 4 gpurun_out/scale2_c3_n$N.err
