mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/final_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r02_train_n1.json 2> gpurun_out/r02_train_n1.err
python bench.py --model raft --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_raft.json 2> gpurun_out/r02_bench_raft.err
RNC_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
cat gpurun_out/final_tests.log; cut -c1-300 gpurun_out/r02_bench.json
