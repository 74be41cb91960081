mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_encoder.py -x -q -m gpu 2>&1 | tail -15; python bench.py --steps 5 --warmup 3 2>&1 | tail -3) > gpurun_out/bench_mid.log 2>&1
cat gpurun_out/bench_mid.log | cut -c1-1500
