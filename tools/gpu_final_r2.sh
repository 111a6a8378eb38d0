#!/bin/bash
# Round-2 closing session: what the driver runs at round end (full GPU suite, smoke, both bench arms at 20/5) + the other workloads.
set -u
TAG=${1:-r2z}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_cpu.json 2>/dev/null; cut -c1-180 gpurun_out/${TAG}_bench_reference_cpu.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_adamvs_b64.json 2> gpurun_out/${TAG}_bench_adamvs_b64.err; python tools/show_bench.py gpurun_out/${TAG}_bench_adamvs_b64.json 2>/dev/null | head -24
python bench.py --math tf32 --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_adamvs_b64_tf32.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_adamvs_b64_tf32.json 2>/dev/null | head -1
python bench.py --workload config4 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config4_b4.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_config4_b4.json 2>/dev/null | head -2
python bench.py --model msrednet --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_msrednet_b64.json 2>/dev/null; python tools/show_bench.py gpurun_out/${TAG}_bench_msrednet_b64.json 2>/dev/null | head -2
python bench.py --workload scene256 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_scene256_n1.json 2>/dev/null; cut -c1-160 gpurun_out/${TAG}_bench_scene256_n1.json
