mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
bash tools/profile_r02.sh
