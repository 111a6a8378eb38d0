#!/bin/bash
# round-2 session Z2: hunt a run-to-run difference: the bit-reproducibility test repeated under each kernel selection
set -u
TAG=${1:-r2z2}
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do python -c "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)" && break; sleep 5; done
cat > gpurun_out/repro_loop.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from tests.helpers import load_golden, rebuild_case
from tests.test_gpu_parity import _model, _dev, _to_dev
from adamvs_b200 import cascade
for name, cls in (("small_d8", "stream"), ("batch2_d8", "whole")):
    g = load_golden(name)
    sd, imgs, proj, dv2, dv3, meta = rebuild_case(g)
    m = _model(cls, sd, meta["ndepths"], meta["num_depth"])
    dv = dv3 if cls == "whole" else dv2
    args = (imgs.to(_dev()), _to_dev(proj), dv.to(_dev()))
    ref = None
    bad = {}
    for it in range(int(sys.argv[1])):
        cap = {}
        with torch.no_grad():
            out = cascade._forward(m, *args, cap)
        cur = {}
        for st in ("stage1", "stage2", "stage3"):
            for k, v in cap[st].items():
                cur[f"{st}/{k}"] = v.clone()
            cur[f"{st}/depth"] = out[st]["depth"].clone()
        if ref is None: ref = cur; continue
        for k in cur:
            if not torch.equal(cur[k], ref[k]):
                bad.setdefault(k, []).append(it)
    print(name, "differences:", {k: (len(v), v[:6]) for k, v in bad.items()})
PY
for env in "X=1" "ADAMVS_CONV2D_MATH=ffma"; do
  echo "== $env"; env $env timeout 600 python gpurun_out/repro_loop.py 24 2>&1 | tail -3
done
