#!/bin/bash
# round-2 session ZB: env / flag experiments on the final build: L2 sub-batching at 320 MB (stage 2 in two halves), batch 48 / 64
set -u
TAG=${1:-r2zb}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
echo "== base"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_base.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_base.json | grep "value\|regnet_red"
echo "== L2 320 MB"; ADAMVS_K3_L2_MB=320 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_l2_320.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_l2_320.json | grep "value\|regnet_red"
for b in 48 64; do
  echo "== batch $b"; timeout 600 python bench.py --batch $b --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_b$b.json 2> gpurun_out/${TAG}_b$b.err; python tools/show_bench.py gpurun_out/${TAG}_b$b.json | grep "value\|regnet_red\|featurenet\|fused"
done
