#!/bin/bash
# round-2 session V: ncu --set full of the TMA-fed tail kernel (stage 3 form and x2 form) inside bench, + launch list at B=16
set -u
TAG=${1:-r2v}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
# launches of tail_tma_kernel in one forward: 48 (stage 1) + 32 (stage 2) + 8 (stage 3) = 88; skip one forward, take stage 2's first and stage 3's first
ADAMVS_BENCH_PROFILING=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'tail_tma' -s 138 -c 1 -f -o gpurun_out/${TAG}_tail_s2 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_ncu_s2.log 2>&1
ADAMVS_BENCH_PROFILING=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'tail_tma' -s 170 -c 1 -f -o gpurun_out/${TAG}_tail_s3 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_ncu_s3.log 2>&1
for r in tail_s2 tail_s3; do ncu -i gpurun_out/${TAG}_$r.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_${r}_source.csv.gz; done
ADAMVS_BENCH_PROFILING=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_launches.log 2>&1
ls -la gpurun_out | grep ${TAG}
