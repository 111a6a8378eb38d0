#!/bin/bash
# round-2 session Z3: context_head with three source columns per thread - parity + bench
set -u
TAG=${1:-r2z3}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "context or feature_net or forward or golden or reproducible" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | grep "value\|featurenet\|context"
