#!/bin/bash
# Round 2 multi-GPU call: bash profiles/r2_multi.sh N   (run under gpurun --gpus N)
# 1. the default N-GPU line (frame shard, weak) with the CTU-row shard (strong) attached   2. lookahead frames sharded (c2)
N=${1:-2}
mkdir -p gpurun_out
P=$((29500 + N))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_c3_n$N.json 2> gpurun_out/scale_c3_n$N.err; echo "c3 N=$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --config c2 --gpus $N --steps 3 --warmup 2 > gpurun_out/scale_c2_n$N.json 2> gpurun_out/scale_c2_n$N.err; echo "c2 N=$N rc=$?"
if [ "$N" = "8" ]; then
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+40)) bench.py --config c5 --gpus $N --steps 2 --warmup 1 --shard rows > gpurun_out/scale_c5_rows_n$N.json 2> gpurun_out/scale_c5_rows_n$N.err; echo "c5 rows N=$N rc=$?"
fi
if [ "$N" = "4" ]; then
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+40)) bench.py --config c4 --gpus $N --steps 3 --warmup 2 > gpurun_out/scale_c4_n$N.json 2> gpurun_out/scale_c4_n$N.err; echo "c4 N=$N rc=$?"
fi
for f in gpurun_out/scale_*_n$N.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[1], d["n_gpus"], round(d["value"]), d["unit"], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]),
                                            "rows:", (round(d["strong_scaling_rows"]["value"]), round(d["strong_scaling_rows"]["ms_per_step"], 2)) if "strong_scaling_rows" in d else None)
except Exception as e:
    print(sys.argv[1], "failed", e)
PY
done
tail -n 4 gpurun_out/scale_*_n$N.err
