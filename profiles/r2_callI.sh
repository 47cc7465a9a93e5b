#!/bin/bash
# Round 2, call I: inverse 32x32 transform with its loads issued ahead of the shared stores (parity + bandwidth).
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_table.py tests/test_gpu_testbench.py tests/test_gpu_frame.py -m gpu -q -x) > gpurun_out/tI.log 2>&1
tail -n 4 gpurun_out/tI.log | cut -c1-300
timeout 200 python profiles/primitive_bench.py --only "k_transform" --frames 24 --reps 5 --depth 8 2>&1 | grep k_transform
