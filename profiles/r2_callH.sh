#!/bin/bash
# Round 2, call H: inverse IMMA transform with the input staged through a padded shared tile (parity + bandwidth).
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_table.py tests/test_gpu_testbench.py tests/test_gpu_frame.py tests/test_gpu_encoder.py -m gpu -q -x) > gpurun_out/tH.log 2>&1
tail -n 5 gpurun_out/tH.log | cut -c1-300
timeout 200 python profiles/primitive_bench.py --only k_transform --frames 24 --reps 5 2>&1 | grep k_transform
true
