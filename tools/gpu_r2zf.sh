#!/bin/bash
# round-2 session ZF: 48-wide K2 boxes: parity of everything on K1/K2/K5 (+ forced cuts), then ncu --set full of K2 stage 2 inside bench
set -u
TAG=${1:-r2zf}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pair_score or fused_volume or variance or warp or rough or behind or forward or cost_volume or msred or config4" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
ADAMVS_BENCH_PROFILING=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:'warp_volume_tma' -s 6 -c 1 -f -o gpurun_out/${TAG}_k2 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
