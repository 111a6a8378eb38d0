#!/bin/bash
# round-2 session Z8: same-box A/B of the tcgen05 conv's input TMA ring: up to 6 slots (shipped) | up to 9
set -u
TAG=${1:-r2z8}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
for v in base slot9 base slot9; do
  cp adamvs_b200/build/lib_$v.so adamvs_b200/libadamvs_b200.so
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  echo "== $v"; python tools/show_bench.py gpurun_out/${TAG}_bench_$v.json 2>/dev/null | grep "value\|regnet_red\|conv3x3/all"
done
