mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/fulltests.log 2>&1
cat gpurun_out/fulltests.log
