#!/bin/bash
# round-2 session Z9: alternating tile order along the K3 kernel chain (ADAMVS_K3_ZIGZAG=0 off | default on): parity, then same-box A/B
set -u
TAG=${1:-r2z9}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "regnet or reproducible or forward_matches or msred_vs" ) 2>&1 | tail -2
for v in 0 1 0 1; do
  ADAMVS_K3_ZIGZAG=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_zz$v.json 2> gpurun_out/${TAG}_bench_zz$v.err
  echo "== zigzag $v"; python tools/show_bench.py gpurun_out/${TAG}_bench_zz$v.json 2>/dev/null | grep "value\|regnet_red"
done
