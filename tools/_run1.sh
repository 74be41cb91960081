export RNC_GRAPH=0
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc_launches.csv python tools/enc_launches.py > gpurun_out/enc.log 2>&1
tail -3 gpurun_out/enc.log
python tools/step_breakdown.py
