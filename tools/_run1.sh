export RNC_GRAPH=0
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_encoder.py tests/test_gpu_parity.py tests/test_gpu_bench_config.py -x -q -m gpu 2>&1 | tail -15; python tools/iter_kernels.py; python tools/step_breakdown.py) > gpurun_out/tma_epi.log 2>&1
cat gpurun_out/tma_epi.log
