#!/bin/bash
# round-2 session Y: TMA-fed persistent transposed conv - parity, then bench A/B (ADAMVS_DECONV_CFG=0 keeps the plain kernel)
set -u
TAG=${1:-r2y}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "deconv or feature_net or forward_small or golden or context" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for cfg in 0 1; do
  ADAMVS_DECONV_CFG=$cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_deconv${cfg}.json 2> gpurun_out/${TAG}_bench_deconv${cfg}.err
  echo "deconv cfg $cfg"; python tools/show_bench.py gpurun_out/${TAG}_bench_deconv${cfg}.json 2>/dev/null | grep "value\|featurenet\|deconv\|pair_unet\|context"
done
