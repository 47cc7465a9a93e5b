mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --maxfail=4 -k "frame or window or me_batch") > gpurun_out/t_frame.log 2>&1
tail -n 40 gpurun_out/t_frame.log | cut -c1-400
(time timeout 900 python -m pytest tests/test_gpu_testbench.py tests/test_gpu_table.py -q --maxfail=4) > gpurun_out/t_tb.log 2>&1
tail -n 30 gpurun_out/t_tb.log | cut -c1-600
