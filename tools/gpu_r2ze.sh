#!/bin/bash
# round-2 session ZE: K2 source boxes 32 wide (shipped) | 48 wide (row-to-row bank shift of 16): parity of K1/K2/K5, then same-box A/B
set -u
TAG=${1:-r2ze}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
cp adamvs_b200/build/lib_bw48.so adamvs_b200/libadamvs_b200.so
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pair_score or fused_volume or variance or warp or rough or behind or forward_matches or cost_volume" ) 2>&1 | tail -2
for v in base bw48 base bw48; do
  cp adamvs_b200/build/lib_$v.so adamvs_b200/libadamvs_b200.so
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  echo "== $v"; python tools/show_bench.py gpurun_out/${TAG}_bench_$v.json 2>/dev/null | grep "value\|fused_volume\|pair_score"
done
