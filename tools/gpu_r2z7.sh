#!/bin/bash
# round-2 session Z7: same-box A/B/C of the tcgen05 conv: shipped | separate epilogue-operand TMA producer | + 4 operand stages
set -u
TAG=${1:-r2z7}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
cp adamvs_b200/libadamvs_b200.so adamvs_b200/build/lib_base.so
for v in base eprod eprod_na4 base eprod; do
  cp adamvs_b200/build/lib_$v.so adamvs_b200/libadamvs_b200.so
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  echo "== $v"; python tools/show_bench.py gpurun_out/${TAG}_bench_$v.json 2>/dev/null | grep "value\|regnet_red"
done
cp adamvs_b200/build/lib_eprod_na4.so adamvs_b200/libadamvs_b200.so
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "regnet_red or reproducible or forward_matches" ) 2>&1 | tail -2
