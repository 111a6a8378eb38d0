#!/bin/bash
# round-2 session C: GPU suite (after CUDA is up), K3 L2 sub-batch sweep in bench
set -u
TAG=${1:-r2c}
mkdir -p gpurun_out
nvidia-smi -L
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; echo "cuda not up yet ($i)"; sleep 5; done
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -25 gpurun_out/${TAG}_pytest_gpu.log
for MB in 0 48 72 110; do
  ADAMVS_K3_L2_MB=$MB python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_l2_${MB}.json 2> gpurun_out/${TAG}_bench_l2_${MB}.err
  echo "== L2_MB=$MB"; tail -2 gpurun_out/${TAG}_bench_l2_${MB}.err; python tools/show_bench.py gpurun_out/${TAG}_bench_l2_${MB}.json | head -8
done
