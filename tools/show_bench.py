#!/usr/bin/env python
"""Pretty-print bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json, sys
for path in sys.argv[1:]:
    j = json.loads(open(path).read().strip().splitlines()[-1])
    print(path, "value %.1f maps/s  e2e %.1f  ms/step %.2f  launches %s  B=%s" % (
        j["value"], j["e2e"]["value"], j["ms_per_step"], j.get("gpu_launches"), j["config"].get("batch_per_gpu_per_step")), j.get("clocks"))
    r = j.get("roofline", {})
    print("  roofline", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("kernel", "achieved", "peak", "frac", "traffic")})
    for k, v in sorted(j.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
        extra = ""
        if "achieved_GBps" in v: extra = "%.0f GB/s (%.1f%% hbm)" % (v["achieved_GBps"], 100 * v["frac_of_hbm_peak"])
        if "achieved_TFLOPs" in v: extra = "%.1f TFLOP/s fp32-equivalent (%.1f%% of ffma peak)" % (v["achieved_TFLOPs"], 100 * v["frac_of_ffma_peak"])
        if "tf32_TFLOPs_issued" in v: extra += ", %.0f tf32 TFLOP/s issued (%.1f%% of tf32 tensor peak)" % (v["tf32_TFLOPs_issued"], 100 * v.get("tf32_issue_utilisation", v.get("frac_of_tf32_tensor_peak", 0)))
        print("   %-28s %8.3f ms  %5.1f%%  %s" % (k, v["ms_per_step"], 100 * v["share_of_step"], extra))
    if "cpu_baseline" in j: print("  cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"], j["cpu_baseline"]["kind"])
    if "library_bar" in j: print("  library_bar", json.dumps(j["library_bar"]))
