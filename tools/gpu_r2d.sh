#!/bin/bash
# round-2 session D: K3 parity subset under NSET=4, NSET sweep + batch 64 in bench
set -u
TAG=${1:-r2d}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; echo "cuda not up yet ($i)"; sleep 5; done
( time python -m pytest tests -m gpu -q -x -k "regnet_red or forward_matches or bit_reproducible or full_size or native_conv3x3 or feature_net" ) > gpurun_out/${TAG}_pytest_k3.log 2>&1; tail -6 gpurun_out/${TAG}_pytest_k3.log
python tools/tc_regnet_check.py --batch 8 --planes 4 2>&1 | tail -3
for NS in 2 3 4; do
  ADAMVS_TC_NSET=$NS python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_nset${NS}.json 2> gpurun_out/${TAG}_bench_nset${NS}.err
  echo "== NSET=$NS"; tail -2 gpurun_out/${TAG}_bench_nset${NS}.err; python tools/show_bench.py gpurun_out/${TAG}_bench_nset${NS}.json 2>/dev/null | head -9
done
python bench.py --steps 6 --warmup 3 --batch 64 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_b64.json 2> gpurun_out/${TAG}_bench_b64.err
echo "== B=64"; tail -2 gpurun_out/${TAG}_bench_b64.err; python tools/show_bench.py gpurun_out/${TAG}_bench_b64.json 2>/dev/null | head -9
