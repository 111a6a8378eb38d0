#!/bin/bash
set -u
TAG=${1:-r2i}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time python -m pytest tests -m gpu -q -x -k "regnet or forward_matches or bit_reproducible or full_size or native_conv3x3 or feature_net or msred" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
python tools/tc_regnet_check.py --batch 32 --planes 4 --stages 3 2>&1 | tail -1 | cut -c1-600
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; python tools/show_bench.py gpurun_out/${TAG}_bench.json | head -12
python bench.py --workload scene256 --steps 2 --warmup 1 > gpurun_out/${TAG}_scene256.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_scene256.json
python bench.py --model msrednet --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_msred_b32.json 2> gpurun_out/${TAG}_msred_b32.err; tail -2 gpurun_out/${TAG}_msred_b32.err; python tools/show_bench.py gpurun_out/${TAG}_msred_b32.json | head -8
