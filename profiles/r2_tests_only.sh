mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -k "frame or window or me_batch") > gpurun_out/t_frame.log 2>&1
tail -n 60 gpurun_out/t_frame.log
