#!/bin/bash
# round-2 session S: TMA-fed persistent tail kernel - parity, then A/B in bench (ADAMVS_TAIL_CFG=0 old | 16 | 24 | auto)
set -u
TAG=${1:-r2s}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
for cfg in 16 24; do
  ( ADAMVS_TAIL_CFG=$cfg timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "regnet or forward or reproducible" ) > gpurun_out/${TAG}_pytest_tail${cfg}.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_tail${cfg}.log
done
for cfg in 0 auto; do
  ADAMVS_TAIL_CFG=$cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_tail_${cfg}.json 2> gpurun_out/${TAG}_bench_tail_${cfg}.err
  echo "tail cfg $cfg"; python tools/show_bench.py gpurun_out/${TAG}_bench_tail_${cfg}.json 2>/dev/null | head -12
done
