mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5; python bench.py --mode train --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400; python tools/train_profile.py 2>&1 | grep -v "^-" | cut -c1-62,140-260 | head -34) > gpurun_out/train2.log 2>&1
cat gpurun_out/train2.log
