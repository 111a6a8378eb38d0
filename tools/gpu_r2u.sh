#!/bin/bash
# round-2 session U: ncu --set full of the K2 stage-2 kernel inside bench (current build)
set -u
TAG=${1:-r2u}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
ADAMVS_BENCH_PROFILING=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'warp_volume_tma' -s 6 -c 1 -f -o gpurun_out/${TAG}_k2 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_k2.ncu-rep --page raw --csv > gpurun_out/${TAG}_k2_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k2.ncu-rep --page source --csv > gpurun_out/${TAG}_k2_source.csv 2>/dev/null
ls -la gpurun_out | grep ${TAG}
