#!/bin/bash
# round-2 session A: full GPU parity suite, smoke, headline bench (with library_bar + CPU reference), config4 bench
set -u
TAG=${1:-r2a}
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -25 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_b32.json 2> gpurun_out/${TAG}_bench_b32.err; tail -2 gpurun_out/${TAG}_bench_b32.err; python tools/show_bench.py gpurun_out/${TAG}_bench_b32.json
python bench.py --workload config4 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_config4.json 2> gpurun_out/${TAG}_bench_config4.err; tail -2 gpurun_out/${TAG}_bench_config4.err; python tools/show_bench.py gpurun_out/${TAG}_bench_config4.json
