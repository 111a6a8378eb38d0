#!/bin/bash
# One GPU session: parity tests, stand-alone kernel timings, cuDNN-part variants, bench, ncu captures.
# usage: tools/gpu_round.sh <tag> [what...]   what in: test kern feat bench ncu_k2 ncu_k3 launches
set -u
TAG=${1:-rXX}; shift
WHAT=${*:-test kern bench}
mkdir -p gpurun_out
# .ncu-rep files with imported source are tens of MB each and gpurun_out/ is capped at 64 MiB: export the
# raw-metric and per-instruction source pages as (gzipped) CSV on the box and drop the report.
export_rep() { for r in "$@"; do ncu -i $r.ncu-rep --page raw --csv > ${r}_raw.csv 2>/dev/null; ncu -i $r.ncu-rep --page source --csv 2>/dev/null | gzip -9 > ${r}_source.csv.gz; rm -f $r.ncu-rep; done; }
for w in $WHAT; do
case $w in
test) python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log;;
kern) python tools/prof_kernels.py --batch 8 > gpurun_out/${TAG}_kernels_b8.jsonl 2> gpurun_out/${TAG}_kernels_b8.err; cat gpurun_out/${TAG}_kernels_b8.jsonl;;
kern1) python tools/prof_kernels.py --batch 1 > gpurun_out/${TAG}_kernels_b1.jsonl 2> gpurun_out/${TAG}_kernels_b1.err; cat gpurun_out/${TAG}_kernels_b1.jsonl;;
feat) python tools/prof_featurenet.py --batch 8 > gpurun_out/${TAG}_featurenet.jsonl 2> gpurun_out/${TAG}_featurenet.err; cat gpurun_out/${TAG}_featurenet.jsonl;;
bench) python bench.py --steps 10 --warmup 3 --batch 8 > gpurun_out/${TAG}_bench_b8.json 2> gpurun_out/${TAG}_bench_b8.err; python tools/show_bench.py gpurun_out/${TAG}_bench_b8.json;;
benchms) python bench.py --model msrednet --steps 5 --warmup 3 --batch 4 > gpurun_out/${TAG}_bench_msred_b4.json 2> gpurun_out/${TAG}_bench_msred_b4.err; python tools/show_bench.py gpurun_out/${TAG}_bench_msred_b4.json; tail -3 gpurun_out/${TAG}_bench_msred_b4.err;;
bench1) python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_b1.json 2> gpurun_out/${TAG}_bench_b1.err; python tools/show_bench.py gpurun_out/${TAG}_bench_b1.json;;
ncu_k2) ncu --set full --clock-control none --import-source on -k regex:'fused_volume_kernel|pair_score_kernel|warp_volume' -s 2 -c 10 -f -o gpurun_out/${TAG}_k2 \
    python tools/prof_kernels.py --batch 8 --only k1,k2 --iters 1 > gpurun_out/${TAG}_ncu_k2.log 2>&1; export_rep gpurun_out/${TAG}_k2;;
ncu_k3s2) for S in 2; do ncu --set full --clock-control none --import-source on -k regex:"conv3x3|upconv|regress" -s 49 -c 7 -f -o gpurun_out/${TAG}_k3_s${S} python tools/prof_kernels.py --batch 8 --only k3 --stages $S --planes 3 --iters 1 > gpurun_out/${TAG}_ncu_k3_s${S}.log 2>&1; export_rep gpurun_out/${TAG}_k3_s${S}; done;;
ncu_k3s3) for S in 3; do ncu --set full --clock-control none --import-source on -k regex:'conv3x3|upconv|regress' -s 56 -c 8 -f -o gpurun_out/${TAG}_k3_s${S} \
    python tools/prof_kernels.py --batch 8 --only k3 --stages $S --planes 3 --iters 1 > gpurun_out/${TAG}_ncu_k3_s${S}.log 2>&1; export_rep gpurun_out/${TAG}_k3_s${S}; done;;
ncu_k3) for S in 3 2 1; do ncu --set full --clock-control none --import-source on -k regex:'conv3x3|upconv|regress' -s 56 -c 8 -f -o gpurun_out/${TAG}_k3_s${S} \
    python tools/prof_kernels.py --batch 8 --only k3 --stages $S --planes 3 --iters 1 > gpurun_out/${TAG}_ncu_k3_s${S}.log 2>&1; export_rep gpurun_out/${TAG}_k3_s${S}; done;;
launches) ADAMVS_BENCH_PROFILING=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1;;
esac
done
