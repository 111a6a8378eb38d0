#!/bin/bash
# round-2 session Z4 (2 GPUs): the final build under torch.distributed.run - weak-scaling line and the 256-view scene
set -u
TAG=${1:-r2z4}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
python tools/show_bench.py gpurun_out/${TAG}_bench_n2.json 2>/dev/null | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload scene256 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_scene256_n2.json 2> gpurun_out/${TAG}_scene_n2.err
cut -c1-200 gpurun_out/${TAG}_bench_scene256_n2.json
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "device" 2>&1 | tail -2
