#!/bin/bash
# round-2 session E: source-level stall profile of the K3 gate / candidate tcgen05 kernels (stage 3, B=16, second plane)
set -u
TAG=${1:-r2e}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
ncu --set full --clock-control none --import-source on -k regex:'conv3x3_tc' -s 6 -c 3 -f -o gpurun_out/${TAG}_k3 \
  python tools/tc_regnet_check.py --batch 16 --planes 2 --stages 3 --no-time > gpurun_out/${TAG}_ncu_k3.log 2>&1
ncu -i gpurun_out/${TAG}_k3.ncu-rep --page raw --csv > gpurun_out/${TAG}_k3_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_k3.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_k3_source.csv.gz
rm -f gpurun_out/${TAG}_k3.ncu-rep
ls -la gpurun_out | grep ${TAG}
