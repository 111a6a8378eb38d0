#!/bin/bash
# round-2 session ZI: MS-REDNet GRU convolutions (levels 1-3) on the tensor-core kernel: parity (both paths forced), then bench A/B
set -u
TAG=${1:-r2zi}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "msred" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -12 gpurun_out/${TAG}_pytest.log
for m in ffma auto; do
  ADAMVS_K3_MATH=$m timeout 900 python bench.py --model msrednet --steps 2 --warmup 2 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_$m.json 2> gpurun_out/${TAG}_bench_$m.err
  echo "== $m"; python tools/show_bench.py gpurun_out/${TAG}_bench_$m.json 2>/dev/null | head -6
done
