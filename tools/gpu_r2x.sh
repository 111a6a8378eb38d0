#!/bin/bash
# round-2 session X: 3-product tf32 split (A_lo meets the W_hi rows only) - parity of everything on the tensor-core conv, then bench
set -u
TAG=${1:-r2x}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "regnet or forward or reproducible or conv3x3 or feature or golden or config4" ) > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -12
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2x_bench.json").read().strip().splitlines()[-1])
print(json.dumps(j.get("library_bar", {}).get("parity_ours_vs_reference_fp32")))
PY
