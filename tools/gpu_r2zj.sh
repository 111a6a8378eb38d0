#!/bin/bash
# round-2 session ZJ: full GPU suite once more on the final tree + the MS-REDNet bench line with its reference arms
set -u
TAG=${1:-r2zj}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --model msrednet --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_msrednet_b64.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_msrednet_b64.json 2>/dev/null | head -8
python bench.py --model msrednet --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_msrednet_b32.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_msrednet_b32.json 2>/dev/null | head -1
