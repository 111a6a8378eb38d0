#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, Python/PyTorch,
CPU) on seeded synthetic inputs.  Runs only in the build container (the reference does not travel to
the GPU box); the fixtures it writes are committed and are what pins oracle/adamvs_oracle.py.

    python tests/golden/make_golden.py            # rewrites every fixture

Accommodations needed to execute the reference on CPU:
  * models/adamvs.py:175-176,448-459 call .cuda() inside forward -> Tensor.cuda patched to identity
    for the duration of this script;
  * torch.meshgrid / F.upsample deprecation warnings are silenced.
Weights and inputs come from adamvs_b200.synth (numpy PCG64 streams), so a fixture stores only the
seeds, the calibration scalars and the reference's outputs.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

CASES = {
    # name: (B, H, W, ndepths, num_depth, gain, weight_seed, input_seed, store_intermediates)
    "small_d8": (1, 64, 96, (8, 4, 2), 32, 20.0, 11, 3, True),
    "batch2_d8": (2, 64, 96, (8, 4, 2), 32, 20.0, 12, 5, False),
    "full_d48": (1, 64, 128, (48, 32, 8), 192, 60.0, 13, 7, False),
}


def _bind_reference(module):
    """The reference's models/<module>.py.  Its models/ has no __init__.py and this repo has a package of the same
    name, so `import models.adamvs` would resolve to OUR drop-in: bind the reference's directory explicitly to a
    private package name instead (relative imports inside the reference keep working)."""
    import importlib
    import types
    if "refmodels" not in sys.modules:
        pkg = types.ModuleType("refmodels")
        pkg.__path__ = [os.path.join(REF, "models")]
        sys.modules["refmodels"] = pkg
    mod = importlib.import_module("refmodels." + module)
    assert mod.__file__.startswith(REF), mod.__file__
    return mod


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    from adamvs_b200 import synth            # imported before the reference shadows nothing of ours
    ref_adamvs = _bind_reference("adamvs")
    torch.Tensor.cuda = lambda self, *a, **k: self             # CPU shim for the hard-coded .cuda()
    torch.set_num_threads(os.cpu_count())

    for name, (B, H, W, ndepths, num_depth, gain, wseed, iseed, store) in CASES.items():
        imgs, proj, dv2 = synth.make_sample(B, H, W, 5, seed=iseed)
        interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / num_depth
        dv3 = torch.cat([dv2, torch.full((B, 1), interval)], 1)
        sd = synth.fill_state_dict(synth.state_dict_shapes(ndepths[0]), wseed)

        whole = ref_adamvs.AdaMVSNet(ndepths=list(ndepths), depth_intervals_ratio=[4.0, 2.0, 1.0]).eval()
        whole.load_state_dict(sd)
        with torch.no_grad():
            f = whole.feature(imgs[:, 0])
        fstd = {k: float(f[k].std()) for k in ("stage1", "stage2", "stage3")}
        sd = synth.calibrate_state_dict(sd, fstd, gain)
        whole.load_state_dict(sd)
        stream = ref_adamvs.Infer_AdaMVSNet(num_depth=num_depth, ndepths=list(ndepths),
                                            depth_intervals_ratio=[4.0, 2.0, 1.0]).eval()
        stream.load_state_dict(sd)

        blob = {"meta_B": B, "meta_H": H, "meta_W": W, "meta_ndepths": np.array(ndepths),
                "meta_num_depth": num_depth, "meta_gain": gain, "meta_wseed": wseed, "meta_iseed": iseed,
                "meta_fstd": np.array([fstd["stage1"], fstd["stage2"], fstd["stage3"]], dtype=np.float64)}

        grabbed = {}
        hooks = []
        if store:
            for i in range(3):
                hooks.append(whole.DepthNet[i].reg_fuse.register_forward_hook(
                    lambda m, a, o, i=i: grabbed.__setitem__(f"whole_s{i + 1}_fused", a[0].detach().clone())
                    or grabbed.__setitem__(f"whole_s{i + 1}_logits", o.detach().clone())))
            scores = []
            hooks.append(whole.DepthNet[0].reg.register_forward_hook(
                lambda m, a, o: scores.append((a[0].detach().clone(), o.detach().clone()))))
        with torch.no_grad():
            ow = whole(imgs, proj, dv3)
            if store:
                grabbed["whole_s1_pair_score"] = torch.stack([s[0] for s in scores], 1)
                grabbed["whole_s1_pair_logits"] = torch.stack([s[1] for s in scores], 1)
                feats = [whole.feature(imgs[:, v]) for v in range(5)]
                for k in ("stage1", "stage2", "stage3"):
                    grabbed[f"features_{k}"] = torch.stack([f[k] for f in feats], 1)
            for h in hooks:
                h.remove()
            os_ = stream(imgs, proj, dv2)
        for tag, out in (("whole", ow), ("stream", os_)):
            for s in ("stage1", "stage2", "stage3"):
                blob[f"{tag}_{s}_depth"] = out[s]["depth"].numpy()
                blob[f"{tag}_{s}_conf"] = out[s]["photometric_confidence"].numpy()
                blob[f"{tag}_{s}_pair_conf4"] = torch.stack(out[s]["pair_confidence"][:4], 1).numpy()
                blob[f"{tag}_{s}_pair_conf_len"] = len(out[s]["pair_confidence"])
                if len(out[s]["pair_result"]):
                    blob[f"{tag}_{s}_pair_result"] = torch.stack(out[s]["pair_result"], 1).numpy()
            assert out["depth"] is out["stage3"]["depth"]
        for k, v in grabbed.items():
            blob[k] = v.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB",
              "conf range s3:", float(ow["photometric_confidence"].min()), float(ow["photometric_confidence"].max()))


MSRED_CASES = {
    # name: (B, H, W, ndepths, num_depth, gain, weight_seed, input_seed, store_intermediates)
    "msred_small_d8": (1, 64, 96, (8, 4, 2), 32, 4.0, 41, 9, True),
    "msred_batch2_d6": (2, 64, 128, (6, 4, 2), 24, 4.0, 42, 10, False),
}


def main_msred():
    """Same for MS-REDNet (BASELINE config 5): the reference's CascadeREDNet / Infer_CascadeREDNet."""
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    from adamvs_b200 import synth
    import contextlib
    import io
    ref = _bind_reference("msrednet")
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.set_num_threads(os.cpu_count())
    for name, (B, H, W, ndepths, num_depth, gain, wseed, iseed, store) in MSRED_CASES.items():
        imgs, proj, dv2 = synth.make_sample(B, H, W, 5, seed=iseed)
        interval = (synth.DEPTH_MAX - synth.DEPTH_MIN) / num_depth
        dv3 = torch.cat([dv2, torch.full((B, 1), interval)], 1)
        sd = synth.fill_state_dict(synth.msred_state_dict_shapes(), wseed)
        with contextlib.redirect_stdout(io.StringIO()):
            whole = ref.CascadeREDNet(ndepths=list(ndepths), depth_interals_ratio=[4.0, 2.0, 1.0]).eval()
            stream = ref.Infer_CascadeREDNet(num_depth=num_depth, ndepths=list(ndepths), depth_interals_ratio=[4.0, 2.0, 1.0]).eval()
        whole.load_state_dict(sd)
        with torch.no_grad():
            f = whole.feature(imgs[:, 0])
        fstd = {k: float(f[k].std()) for k in ("stage1", "stage2", "stage3")}
        sd = synth.calibrate_msred_state_dict(sd, fstd, gain)
        whole.load_state_dict(sd)
        stream.load_state_dict(sd)
        blob = {"meta_B": B, "meta_H": H, "meta_W": W, "meta_ndepths": np.array(ndepths), "meta_num_depth": num_depth,
                "meta_gain": gain, "meta_wseed": wseed, "meta_iseed": iseed,
                "meta_fstd": np.array([fstd["stage1"], fstd["stage2"], fstd["stage3"]], dtype=np.float64)}
        grabbed, hooks = {}, []
        if store:
            for i in range(3):
                hooks.append(whole.cost_regularization[i].register_forward_hook(
                    lambda m, a, o, i=i: grabbed.__setitem__(f"whole_s{i + 1}_variance", a[0].detach().clone())
                    or grabbed.__setitem__(f"whole_s{i + 1}_logits", o.detach().clone())))
        with torch.no_grad():
            ow = whole(imgs, proj, dv3)
            for h in hooks:
                h.remove()
            os_ = stream(imgs, proj, dv2)
        for tag, out in (("whole", ow), ("stream", os_)):
            for s in ("stage1", "stage2", "stage3"):
                blob[f"{tag}_{s}_depth"] = out[s]["depth"].numpy()
                blob[f"{tag}_{s}_conf"] = out[s]["photometric_confidence"].numpy()
        for k, v in grabbed.items():
            blob[k] = v.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "stream conf range s3:",
              float(os_["photometric_confidence"].min()), float(os_["photometric_confidence"].max()),
              "depth range s3:", float(os_["depth"].min()), float(os_["depth"].max()))


if __name__ == "__main__":
    if "--msred-only" not in sys.argv:
        main()
    main_msred()
