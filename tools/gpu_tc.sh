#!/bin/bash
# tensor-core K3 check: accuracy vs the FFMA path (small + full shapes), per-plane times, per-kernel launch list
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 120 python tools/tc_regnet_check.py --small --no-time > gpurun_out/${TAG}_tc_small.jsonl 2> gpurun_out/${TAG}_tc_small.err; cat gpurun_out/${TAG}_tc_small.jsonl; tail -2 gpurun_out/${TAG}_tc_small.err
timeout 300 python tools/tc_regnet_check.py --batch 8 --planes 4 > gpurun_out/${TAG}_tc_check.jsonl 2> gpurun_out/${TAG}_tc_check.err; cat gpurun_out/${TAG}_tc_check.jsonl; tail -2 gpurun_out/${TAG}_tc_check.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/${TAG}_tc_launches.csv python tools/tc_regnet_check.py --batch 8 --planes 2 --stages 3,2,1 --no-time > gpurun_out/${TAG}_tc_ncu.log 2>&1
python tools/launch_table.py gpurun_out/${TAG}_tc_launches.csv | grep -v "at::\|pack_conv" | tee gpurun_out/${TAG}_tc_launches.txt
