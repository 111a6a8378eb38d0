#!/bin/bash
# round-2 session Z5: the new parity cases (ragged tail tiles, forced tail / transposed-conv variants, six-forward reproducibility)
set -u
TAG=${1:-r2z5}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "regnet_red_vs_oracle or variants_forced or reproducible or deconv" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log
