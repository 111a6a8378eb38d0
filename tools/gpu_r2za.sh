#!/bin/bash
# round-2 session ZA: parity subset after the tile-order change
set -u
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "regnet or reproducible or forward or msred_vs or conv3x3 or feature" ) > gpurun_out/r2za_pytest.log 2>&1; tail -3 gpurun_out/r2za_pytest.log
