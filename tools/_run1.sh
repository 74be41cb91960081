export RNC_GRAPH=0
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; for r in 1 2; do echo "--- new"; python tools/iter_kernels.py | tail -12; python tools/step_breakdown.py | grep "step\|update\|enc"; echo "--- old"; (cd _old && python tools/iter_kernels.py | tail -1; python tools/step_breakdown.py | grep "step\|update\|enc"); done) > gpurun_out/ab3.log 2>&1
cat gpurun_out/ab3.log
