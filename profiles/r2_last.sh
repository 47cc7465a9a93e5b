#!/bin/bash
# Round 2, last call: the whole GPU suite on the final tree.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -q -x) > gpurun_out/t_last.log 2>&1
tail -n 4 gpurun_out/t_last.log | cut -c1-300
