#!/bin/bash
# Round-end evidence: ncu --set full of the K3 tensor-core kernels and of K2 inside bench, ncu launch list of one bench step.
TAG=${1:-rXX}
mkdir -p gpurun_out
export_rep() { for r in "$@"; do ncu -i $r.ncu-rep --page raw --csv > ${r}_raw.csv 2>/dev/null; rm -f $r.ncu-rep; done; }
for S in 3 2; do
  ncu --set full --clock-control none -k regex:'conv3x3_tc|conv3x3_v2|tail_regress' -s 14 -c 7 -f -o gpurun_out/${TAG}_k3_s${S} \
    python tools/tc_regnet_check.py --batch 8 --planes 2 --stages $S --no-time > gpurun_out/${TAG}_ncu_k3_s${S}.log 2>&1
  export_rep gpurun_out/${TAG}_k3_s${S}
done
ADAMVS_BENCH_PROFILING=1 ncu --set full --clock-control none -k regex:'warp_volume_tma' -s 6 -c 1 -f -o gpurun_out/${TAG}_k2_bench \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_k2_bench.log 2>&1
export_rep gpurun_out/${TAG}_k2_bench
ADAMVS_BENCH_PROFILING=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ls -la gpurun_out | grep ${TAG}
