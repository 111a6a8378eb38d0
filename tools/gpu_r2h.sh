#!/bin/bash
# round-2 session H: full GPU suite, headline bench (full line), scene256 N=1, ncu launch list of one bench step
set -u
TAG=${1:-r2h}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -16 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_b32.json 2> gpurun_out/${TAG}_bench_b32.err; tail -2 gpurun_out/${TAG}_bench_b32.err; python tools/show_bench.py gpurun_out/${TAG}_bench_b32.json
python bench.py --workload scene256 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_scene256_n1.json 2> gpurun_out/${TAG}_bench_scene256_n1.err; tail -2 gpurun_out/${TAG}_bench_scene256_n1.err; cut -c1-400 gpurun_out/${TAG}_bench_scene256_n1.json
ADAMVS_BENCH_PROFILING=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --batch 16 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_ncu_launch.log 2>&1
ls -la gpurun_out | grep ${TAG}
