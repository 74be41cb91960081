mkdir -p gpurun_out
(python tools/l1_probe.py; timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_umma.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5; RNC_GRAPH=0 python tools/step_breakdown.py) > gpurun_out/l1.log 2>&1
cat gpurun_out/l1.log
