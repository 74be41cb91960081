#!/bin/bash
# Round-2 profile captures on the GPU box (run under gpurun from the repo root); summaries are copied into profiles/ afterwards.
set -u
mkdir -p gpurun_out
# 1. launch list of the bench command: every launch with its device time (cold-cache, serialised: compare SHARES)
RNC_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1300 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
# 2. the dominant kernel of the roofline line (corr lookup) and the NCUP chain, full sets, from the same command
RNC_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:corr_lookup_umma -s 70 -c 1 -o gpurun_out/prof_r02_lookup_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lookup_final.log 2>&1
RNC_GRAPH=0 ncu --set full --clock-control none -k regex:ncup_fused -s 2 -c 1 -o gpurun_out/prof_r02_ncup_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ncup_final.log 2>&1
# 3. one update iteration, full set (per-layer tensor activity)
RNC_GRAPH=0 ncu --set full --clock-control none -k regex:"conv_umma|corr_lookup_umma|flow_tap|flow_im2col" -s 400 -c 16 -o gpurun_out/prof_r02_iter \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_iter_final.log 2>&1
# 4. the clean numbers (never taken under a profiler)
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/r02_train_n1.json 2> gpurun_out/r02_train_n1.err
python bench.py --model raft --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_raft.json 2> gpurun_out/r02_bench_raft.err
tail -c 200 gpurun_out/r02_bench.err gpurun_out/ncu_lookup_final.log gpurun_out/ncu_iter_final.log
