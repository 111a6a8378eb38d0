#!/bin/bash
# round-2 session Z6: ncu --set full of every K3 kernel of one stage-3 plane at the bench batch (B = 32), final build
set -u
TAG=${1:-r2z6}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
python tools/prof_kernels.py --batch 32 --planes 4 --iters 3 --only k3 --stages 1,2,3 2>&1 | tail -3
# launches per regnet_red call with 2 planes: 7 pack + hyp_lines + 2 x 7 + regress; 3 calls (2 warm-up + 1): capture the second plane of the last call
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'conv3x3_tc|conv3x3_v2|tail_tma' -s 35 -c 7 -f -o gpurun_out/${TAG}_k3_s3 \
  python tools/prof_kernels.py --batch 32 --planes 2 --iters 1 --only k3 --stages 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_k3_s3.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${TAG}_k3_s3_source.csv.gz
ls -la gpurun_out | grep ${TAG}
