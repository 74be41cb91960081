export RNC_GRAPH=0
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_encoder.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5; python tools/l1_probe.py | head -4; python tools/iter_kernels.py | grep "128->64\|256->18\|sum"; python tools/step_breakdown.py | grep "step\|enc\|update") > gpurun_out/fused64.log 2>&1
cat gpurun_out/fused64.log
