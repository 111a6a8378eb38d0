#!/bin/bash
# round-2 session G: K2 parity subset, bench with plane lanes G=2 / G=4 (auto)
set -u
TAG=${1:-r2g}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time python -m pytest tests -m gpu -q -x -k "pair_score or cost_volume or fused_volume or variance_volume or forward_matches or warp_box or full_size" ) > gpurun_out/${TAG}_pytest_k2.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_k2.log
for G in 2 0; do
  ADAMVS_WARP_G=$G python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_g${G}.json 2> gpurun_out/${TAG}_bench_g${G}.err
  echo "== WARP_G=$G (0 = auto)"; tail -2 gpurun_out/${TAG}_bench_g${G}.err; python tools/show_bench.py gpurun_out/${TAG}_bench_g${G}.json 2>/dev/null | grep -E "value|roofline|fused_volume|pair_score"
done
