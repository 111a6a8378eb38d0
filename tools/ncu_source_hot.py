#!/usr/bin/env python
"""Summarise the per-instruction page of an ncu capture (gzipped CSV made by tools/gpu_round.sh):
per kernel, stall-reason totals and the hottest SASS instructions.
    python tools/ncu_source_hot.py gpurun_out/X_source.csv.gz [kernel-substring] [top-n]"""
import csv, gzip, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = csv.reader(gzip.open(path, "rt"))
kern = None; hdr = None; data = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1]; hdr = None; data[kern] = []; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) >= len(hdr) - 1 and kern:
        data[kern].append(dict(zip(hdr, r)))
for k, ins in data.items():
    if want not in k: continue
    tot = sum(int(i["# Samples"] or 0) for i in ins)
    print("=" * 110); print(k[:200]); print(f"instructions {len(ins)}, samples {tot}")
    stalls = collections.Counter()
    for i in ins:
        for key, v in i.items():
            if key.startswith("stall_") and "Not Issued" not in key and v not in ("", "-"):
                stalls[key] += int(v)
    print("stall samples:", ", ".join(f"{a}={b} ({100*b/max(1,tot):.0f}%)" for a, b in stalls.most_common(8)))
    cls = collections.Counter(); exe = collections.Counter()
    for i in ins:
        op = i["Source"].split()[0] if i["Source"].split() else "?"
        if op.startswith("@"): op = i["Source"].split()[1]
        op = op.split(".")[0]
        cls[op] += int(i["# Samples"] or 0); exe[op] += int(i["Instructions Executed"] or 0)
    te = sum(exe.values())
    print("by opcode (samples | executed):", ", ".join(f"{a}={b} ({100*b/max(1,tot):.0f}%|{100*exe[a]/max(1,te):.0f}%)" for a, b in cls.most_common(10)))
    for i in sorted(ins, key=lambda i: -int(i["# Samples"] or 0))[:topn]:
        st = {k2[6:]: int(v) for k2, v in i.items() if k2.startswith("stall_") and "Not Issued" not in k2 and v not in ("", "-", "0")}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"  {int(i['# Samples']):6d} {100*int(i['# Samples'])/max(1,tot):5.1f}%  {i['Source'].strip()[:70]:70s} {top}")
