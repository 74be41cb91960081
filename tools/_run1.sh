export RNC_GRAPH=0
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_umma.py tests/test_gpu_encoder.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3; python tools/l1_probe.py 2>/dev/null | head -3; python tools/iter_kernels.py) > gpurun_out/desc.log 2>&1
cat gpurun_out/desc.log
