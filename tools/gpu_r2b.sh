#!/bin/bash
# round-2 session B: full GPU suite, bench B=32, ncu --set full of K2 (stage 2) inside bench
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
export_rep() { for r in "$@"; do ncu -i $r.ncu-rep --page raw --csv > ${r}_raw.csv 2>/dev/null; rm -f $r.ncu-rep; done; }
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( time python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -25 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_bench_b32.json 2> gpurun_out/${TAG}_bench_b32.err; tail -2 gpurun_out/${TAG}_bench_b32.err; python tools/show_bench.py gpurun_out/${TAG}_bench_b32.json
ADAMVS_BENCH_PROFILING=1 ncu --set full --clock-control none -k regex:'warp_volume_dm' -s 6 -c 1 -f -o gpurun_out/${TAG}_k2_bench \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-bar > gpurun_out/${TAG}_ncu_k2_bench.log 2>&1
export_rep gpurun_out/${TAG}_k2_bench
ls -la gpurun_out | grep ${TAG}
