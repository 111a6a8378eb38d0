#!/bin/bash
# Run on the GPU box (via gpurun): bench lines + ncu launch list + ncu --set full captures of K2 and the K3 convs.
# usage: tools/gpu_profile.sh <tag>   (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --batch 8 > gpurun_out/${TAG}_bench_b8.json 2> gpurun_out/${TAG}_bench_b8.err
python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_b1.json 2> gpurun_out/${TAG}_bench_b1.err
export ADAMVS_BENCH_PROFILING=1
# launch list of the same command shape (B=1, 1 warm-up + 1 timed step of the resident loop, then e2e steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
# full-set captures: K2 (three stage launches of the second forward) and stage-3 K3 convs
ncu --set full --clock-control none --import-source on -k regex:fused_volume_kernel -s 3 -c 3 -f -o gpurun_out/${TAG}_k2 \
    python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_kernel -s 1008 -c 6 -f -o gpurun_out/${TAG}_k3 \
    python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_k3.log 2>&1
tail -c 600 gpurun_out/${TAG}_bench_b8.json
