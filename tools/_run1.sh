mkdir -p gpurun_out
(python tools/l1_probe.py | head -3; timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -x -q -m gpu 2>&1 | tail -5; RNC_GRAPH=0 python tools/step_breakdown.py) > gpurun_out/stats2.log 2>&1
cat gpurun_out/stats2.log
