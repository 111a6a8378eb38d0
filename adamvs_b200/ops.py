"""ctypes binding of libadamvs_b200.so (the C ABI in include/adamvs_b200.h) for torch tensors.

PyTorch is plumbing here: it owns device memory and streams; every op below marshals
``tensor.data_ptr()`` + shapes + ``torch.cuda.current_stream()`` into one C call.  There is no CPU path
and no torch fallback: without the compiled library or a CUDA tensor these functions raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadamvs_b200.so")

HYP_PLANES, HYP_PER_PIXEL = 0, 1
EPS_NUMERATOR, EPS_DENOMINATOR = 0, 1
PROB_SOFTMAX, PROB_EXP_EPS = 0, 1
INTERVAL_LAST_COLUMN, INTERVAL_FROM_RANGE = 0, 1
MATH_FFMA, MATH_TC_FP32, MATH_TC_TF32, MATH_AUTO = 0, 1, 2, 3

EXPORTS = (
    "adamvs_abi_version", "adamvs_cascade_prepare", "adamvs_pair_score_f32", "adamvs_resize_bilinear_f32",
    "adamvs_fused_volume_f32", "adamvs_regnet_red_workspace_floats", "adamvs_regnet_red_f32",
    "adamvs_regnet_red_ex_f32",
    "adamvs_softmax_regress_f32", "adamvs_variance_volume_f32", "adamvs_regnet_msred_workspace_floats",
    "adamvs_regnet_msred_f32", "adamvs_conv3x3_supported", "adamvs_conv3x3_f32",
    "adamvs_deconv3x3_supported", "adamvs_deconv3x3_f32",
    "adamvs_context_head_supported", "adamvs_context_head_f32",
    "adamvs_deconv3x3_res_f32", "adamvs_context_pool_supported", "adamvs_context_pool_f32",
    "adamvs_conv2d_f32", "adamvs_conv2d_wgrad_f32", "adamvs_pair_score_bwd_f32", "adamvs_fused_volume_bwd_f32",
    "adamvs_softmax_expect_f32", "adamvs_softmax_expect_bwd_f32",
)


class RegnetWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "conv1_w", "gates1_w", "gates1_b", "cand1_w", "cand1_b", "conv2_w", "gates2_w", "gates2_b",
        "cand2_w", "cand2_b", "up1_w", "up1_b", "out_w", "out_b")]


class MsredWeights(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in ("conv1_w", "conv2_w", "conv3_w")] +
                [(n, ctypes.c_void_p * 4) for n in ("gate_w", "gate_b", "rnorm_w", "rnorm_b", "unorm_w", "unorm_b",
                                                     "out_w", "out_b", "onorm_w", "onorm_b")] +
                [(n, ctypes.c_void_p) for n in ("up3_w", "up2_w", "up1_w", "prob_w", "prob_b")])


_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) the in-tree shared library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m adamvs_b200.build` "
                "(adamvs_b200 has no CPU or torch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, ci, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        L.adamvs_abi_version.restype = ci
        L.adamvs_cascade_prepare.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ctypes.POINTER(ci),
                                             ctypes.POINTER(ctypes.c_double), vp, vp, vp]
        L.adamvs_pair_score_f32.argtypes = [vp, vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_resize_bilinear_f32.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_fused_volume_f32.argtypes = [vp, vp, ci, vp, ci, vp, vp, ci, vp, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_regnet_red_workspace_floats.argtypes = [ci, ci, ci, ci, ci, ci]
        L.adamvs_regnet_red_workspace_floats.restype = cs
        L.adamvs_regnet_red_f32.argtypes = [vp, ctypes.POINTER(RegnetWeights), ci, vp, ci, vp, ci, ci, vp, cs,
                                            vp, vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_regnet_red_ex_f32.argtypes = [vp, ctypes.POINTER(RegnetWeights), ci, vp, ci, vp, ci, ci, ci, vp, cs,
                                               vp, vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_softmax_regress_f32.argtypes = [vp, ci, vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_variance_volume_f32.argtypes = [vp, vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_regnet_msred_workspace_floats.argtypes = [ci, ci, ci, ci, ci]
        L.adamvs_regnet_msred_workspace_floats.restype = cs
        L.adamvs_regnet_msred_f32.argtypes = [vp, ctypes.POINTER(MsredWeights), ci, vp, ci, vp, ci, vp, cs,
                                              vp, vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_deconv3x3_supported.argtypes = [ci, ci]
        L.adamvs_deconv3x3_f32.argtypes = [vp, vp, vp, ci, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_conv3x3_supported.argtypes = [ci, ci, ci, ci]
        L.adamvs_context_head_supported.argtypes = [ci, ci, ci]
        L.adamvs_context_head_f32.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_conv3x3_f32.argtypes = [vp, ci, vp, ci, vp, vp, ci, ci, vp, ci, ci, ci, ci, vp]
        L.adamvs_deconv3x3_res_f32.argtypes = [vp, vp, vp, ci, vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_context_pool_supported.argtypes = [ci, ci]
        L.adamvs_context_pool_f32.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
        L.adamvs_conv2d_f32.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_conv2d_wgrad_f32.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_pair_score_bwd_f32.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_fused_volume_bwd_f32.argtypes = [vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.adamvs_softmax_expect_f32.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
        L.adamvs_softmax_expect_bwd_f32.argtypes = [vp, vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp]
        for name in EXPORTS:
            if name not in ("adamvs_abi_version", "adamvs_regnet_red_workspace_floats",
                            "adamvs_regnet_msred_workspace_floats"):
                getattr(L, name).restype = ci
        if L.adamvs_abi_version() != 1:
            raise RuntimeError("libadamvs_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


class AdamvsError(RuntimeError):
    pass


# ---- instrumentation used by bench.py (off by default; never changes what is launched) -------------
LAUNCHES = [0]          # kernels of this library enqueued so far (counted from the known launch lists)
_timing = None          # None | dict: "<op>/<tag>" -> [(start_event, end_event), ...]
_tag = ""


def set_timing(sink):
    global _timing
    _timing = sink


def set_tag(tag: str):
    global _tag
    _tag = tag


class _timed:
    def __init__(self, name, launches):
        self.name, self.launches = name, launches

    def __enter__(self):
        LAUNCHES[0] += self.launches
        if _timing is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _timing is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _timing.setdefault(f"{self.name}/{_tag}", []).append((self.e0, e1))
        return False


timed = _timed      # cascade.py brackets the torch/cuDNN parts with it (0 launches of ours)


def _check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise AdamvsError(f"{what}: argument error {rc} (ADAMVS_EINVAL=-1, ADAMVS_ENOSPACE=-2)")
    raise AdamvsError(f"{what}: CUDA error {rc}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise AdamvsError(f"{name} must be a CUDA tensor (adamvs_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise AdamvsError(f"{name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _guard(t: torch.Tensor):
    """The C library launches on the CURRENT device (it never calls cudaSetDevice; its SM-count / attribute caches are
    per device): make the tensors' device current for the duration of the call, so that
    predict_scene(device=cuda:1) from a process whose current device is cuda:0 launches where the pointers live.
    _stream() is evaluated inside, i.e. it is that device's current stream."""
    if not t.is_cuda:
        raise AdamvsError("adamvs_b200 runs on CUDA tensors only (no CPU fallback)")
    return torch.cuda.device(t.device)


def _same_device(*ts):
    dev = None
    for t in ts:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise AdamvsError(f"all tensors of one call must live on one device, got {dev} and {t.device}")


_default_math = None    # None = the library default (exact split); bench.py --math tf32 sets MATH_TC_TF32


def set_default_math(mode):
    """Arithmetic of regnet_red when the caller does not pass `math` (None = library default, the fp32-parity path)."""
    global _default_math
    _default_math = mode


class Hyp:
    """Depth-hypothesis source for one stage (reference get_depth_range_samples, module.py:646-663)."""

    def __init__(self, mode: int, src: torch.Tensor, half_range: Optional[torch.Tensor] = None):
        self.mode = mode
        self.src = _f32c(src, "hyp_src")
        self.ncol = int(src.shape[1]) if mode == HYP_PLANES else 0
        self.half_range = half_range

    def args(self):
        return self.mode, _p(self.src), self.ncol, _p(self.half_range)


def cascade_prepare(proj: Sequence[torch.Tensor], depth_values: torch.Tensor, interval_mode: int, num_depth: int,
                    ndepths: Sequence[int], ratios: Sequence[float]):
    """-> relproj [3,B,V-1,12], half_range [3] (device)."""
    p1, p2, p3 = (_f32c(p, "proj") for p in proj)
    dv = _f32c(depth_values, "depth_values")
    B, V = p1.shape[0], p1.shape[1]
    _same_device(p1, p2, p3, dv)
    relproj = torch.empty((3, B, V - 1, 12), device=p1.device, dtype=torch.float32)
    half = torch.empty((3,), device=p1.device, dtype=torch.float32)
    nd = (ctypes.c_int * 3)(*[int(x) for x in ndepths])
    rt = (ctypes.c_double * 3)(*[float(x) for x in ratios])
    with _guard(p1), _timed("cascade_prepare", 2):
        _check(lib().adamvs_cascade_prepare(_p(p1), _p(p2), _p(p3), _p(dv), dv.shape[1], B, V, interval_mode,
                                            int(num_depth), nd, rt, _p(relproj), _p(half), _stream()), "cascade_prepare")
    return relproj, half


def pair_score(feat: torch.Tensor, relproj: torch.Tensor, hyp: Hyp, D: int) -> torch.Tensor:
    feat = _f32c(feat, "feat")
    B, V, C, h, w = feat.shape
    out = torch.empty((B, V - 1, D, h, w), device=feat.device, dtype=torch.float32)
    with _guard(feat), _timed("pair_score", 1):
        _check(lib().adamvs_pair_score_f32(_p(feat), _p(_f32c(relproj, "relproj")), *hyp.args(), _p(out),
                                           B, V, C, D, h, w, _stream()), "pair_score")
    return out


def resize_bilinear(x: torch.Tensor, ho: int, wo: int) -> torch.Tensor:
    """x [..., hi, wi] -> [..., ho, wo], align_corners=False."""
    x = _f32c(x, "x")
    lead = x.shape[:-2]
    hi, wi = x.shape[-2:]
    n = 1
    for s in lead:
        n *= int(s)
    out = torch.empty((*lead, ho, wo), device=x.device, dtype=torch.float32)
    with _guard(x), _timed("resize_bilinear", 1):
        _check(lib().adamvs_resize_bilinear_f32(_p(x), _p(out), n, hi, wi, ho, wo, _stream()), "resize_bilinear")
    return out


def fused_volume(feat: torch.Tensor, relproj: torch.Tensor, hyp: Hyp, weights: torch.Tensor, eps_mode: int,
                 D: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    feat = _f32c(feat, "feat")
    B, V, C, h, w = feat.shape
    weights = _f32c(weights, "weights")
    assert tuple(weights.shape) == (B, V - 1, h, w), (weights.shape, (B, V - 1, h, w))
    _same_device(feat, relproj, hyp.src, weights, out)
    if out is None:
        out = torch.empty((B, C, D, h, w), device=feat.device, dtype=torch.float32)
    with _guard(feat), _timed("fused_volume", 1):
        _check(lib().adamvs_fused_volume_f32(_p(feat), _p(_f32c(relproj, "relproj")), *hyp.args(), _p(weights), eps_mode,
                                             _p(out), B, V, C, D, h, w, _stream()), "fused_volume")
    return out


def regnet_workspace_floats(B, C, D, h, w, out_up) -> int:
    return int(lib().adamvs_regnet_red_workspace_floats(B, C, D, h, w, int(out_up)))


def regnet_red(volume: torch.Tensor, weights: dict, hyp: Hyp, out_up: bool, prob_mode: int,
               workspace: Optional[torch.Tensor] = None, want_logits: bool = False, math: Optional[int] = None):
    """volume [B,C,D,h,w] -> depth, conf [B,Ho,Wo] (+ logits [B,D,Ho,Wo] when asked).
    `weights`: name -> contiguous fp32 CUDA tensor in the reference layouts (see RegnetWeights).
    `math`: None = the library default, or MATH_FFMA / MATH_TC_FP32 / MATH_TC_TF32 / MATH_AUTO (include/adamvs_b200.h)."""
    volume = _f32c(volume, "volume")
    B, C, D, h, w = volume.shape
    Ho, Wo = (2 * h, 2 * w) if out_up else (h, w)
    need = regnet_workspace_floats(B, C, D, h, w, out_up)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), device=volume.device, dtype=torch.float32)
    depth = torch.empty((B, Ho, Wo), device=volume.device, dtype=torch.float32)
    conf = torch.empty((B, Ho, Wo), device=volume.device, dtype=torch.float32)
    logits = torch.empty((B, D, Ho, Wo), device=volume.device, dtype=torch.float32) if want_logits else None
    keep = {k: _f32c(v, k) for k, v in weights.items()}
    st = RegnetWeights(**{k: v.data_ptr() for k, v in keep.items()})
    with _guard(volume), _timed("regnet_red", 9 + 7 * D):         # weight packs + hypothesis lines, 7 kernels per plane, regression
        if math is None:
            math = _default_math
        if math is None:
            rc = lib().adamvs_regnet_red_f32(_p(volume), ctypes.byref(st), *hyp.args(), int(out_up), prob_mode,
                                             _p(workspace), workspace.numel(), _p(depth), _p(conf), _p(logits),
                                             B, C, D, h, w, _stream())
        else:
            rc = lib().adamvs_regnet_red_ex_f32(_p(volume), ctypes.byref(st), *hyp.args(), int(out_up), prob_mode, int(math),
                                                _p(workspace), workspace.numel(), _p(depth), _p(conf), _p(logits),
                                                B, C, D, h, w, _stream())
        _check(rc, "regnet_red")
    return (depth, conf, logits) if want_logits else (depth, conf)


def softmax_regress(logits: torch.Tensor, hyp: Hyp, prob_mode: int, n_per_batch: int = 1):
    """logits [N,D,h,w] -> depth, conf [N,h,w]."""
    logits = _f32c(logits, "logits")
    N, D, h, w = logits.shape
    depth = torch.empty((N, h, w), device=logits.device, dtype=torch.float32)
    conf = torch.empty((N, h, w), device=logits.device, dtype=torch.float32)
    with _guard(logits), _timed("softmax_regress", 1):
        _check(lib().adamvs_softmax_regress_f32(_p(logits), *hyp.args(), prob_mode, _p(depth), _p(conf),
                                                N, n_per_batch, D, h, w, _stream()), "softmax_regress")
    return depth, conf


# ---- MS-REDNet (config 5) ----------------------------------------------------------------------------

def variance_volume(feat: torch.Tensor, relproj: torch.Tensor, hyp: Hyp, D: int, out: Optional[torch.Tensor] = None):
    """feat [B,V,C,h,w] -> variance over the reference and the warped source views [B,C,D,h,w]."""
    feat = _f32c(feat, "feat")
    B, V, C, h, w = feat.shape
    if out is None:
        out = torch.empty((B, C, D, h, w), device=feat.device, dtype=torch.float32)
    with _guard(feat), _timed("variance_volume", 1):
        _check(lib().adamvs_variance_volume_f32(_p(feat), _p(_f32c(relproj, "relproj")), *hyp.args(), _p(out),
                                                B, V, C, D, h, w, _stream()), "variance_volume")
    return out


def regnet_msred(volume: torch.Tensor, weights: dict, hyp: Hyp, prob_mode: int,
                 workspace: Optional[torch.Tensor] = None, want_logits: bool = False):
    """volume [B,C,D,h,w] -> depth, conf [B,h,w] (+ logits [B,D,h,w] when asked).  `weights`: field name of
    MsredWeights -> tensor, or list of 4 tensors (index 0..3 = conv_gru1..conv_gru4) for the per-level fields."""
    volume = _f32c(volume, "volume")
    B, C, D, h, w = volume.shape
    need = int(lib().adamvs_regnet_msred_workspace_floats(B, C, D, h, w))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), device=volume.device, dtype=torch.float32)
    depth = torch.empty((B, h, w), device=volume.device, dtype=torch.float32)
    conf = torch.empty((B, h, w), device=volume.device, dtype=torch.float32)
    logits = torch.empty((B, D, h, w), device=volume.device, dtype=torch.float32) if want_logits else None
    keep = []
    st = MsredWeights()
    for name, _ in MsredWeights._fields_:
        v = weights[name]
        if isinstance(v, (list, tuple)):
            ts = [_f32c(t, name) for t in v]
            keep.extend(ts)
            setattr(st, name, (ctypes.c_void_p * 4)(*[t.data_ptr() for t in ts]))
        else:
            t = _f32c(v, name)
            keep.append(t)
            setattr(st, name, t.data_ptr())
    with _guard(volume), _timed("regnet_msred", 11 + 25 * D):
        _check(lib().adamvs_regnet_msred_f32(_p(volume), ctypes.byref(st), *hyp.args(), prob_mode,
                                             _p(workspace), workspace.numel(), _p(depth), _p(conf), _p(logits),
                                             B, C, D, h, w, _stream()), "regnet_msred")
    return (depth, conf, logits) if want_logits else (depth, conf)


# ---- 3x3 convolutions of FeatureNet0 / CostRegNet2D -----------------------------------

def conv3x3_supported(ca: int, cb: int, cout: int, stride: int) -> bool:
    return bool(lib().adamvs_conv3x3_supported(int(ca), int(cb), int(cout), int(stride)))


def pack_conv3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """[COUT,CIN,3,3] -> [CIN,9,COUT] (the layout the kernels keep resident in shared memory)."""
    return w.permute(1, 2, 3, 0).reshape(w.shape[1], 9, w.shape[0]).contiguous()


def conv3x3(xa: torch.Tensor, xb: Optional[torch.Tensor], wpk: torch.Tensor, bias: torch.Tensor, relu: bool, stride: int = 1):
    """act(conv3x3(cat(xa, xb)) + bias): xa [N,CA,h,w], xb [N,CB,h,w] or None, wpk from pack_conv3x3_weight."""
    xa = _f32c(xa, "xa")
    N, CA, h, w = xa.shape
    CB = 0
    if xb is not None:
        xb = _f32c(xb, "xb")
        CB = xb.shape[1]
    COUT = wpk.shape[2]
    out = torch.empty((N, COUT, h // stride, w // stride), device=xa.device, dtype=torch.float32)
    with _guard(xa), _timed("conv3x3", 1):
        _check(lib().adamvs_conv3x3_f32(_p(xa), CA, _p(xb), CB, _p(_f32c(wpk, "wpk")), _p(_f32c(bias, "bias")), int(relu),
                                        int(stride), _p(out), N, COUT, h, w, _stream()), "conv3x3")
    return out


def context_head_supported(cx: int, cctx: int, cout: int) -> bool:
    return bool(lib().adamvs_context_head_supported(int(cx), int(cctx), int(cout)))


def context_head(x: torch.Tensor, ctx_a: torch.Tensor, ctx_c: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """conv1x1(cat(up(ctx_a), up(ctx_c), x), weight) with up() = bilinear resize to x's size (align_corners=False):
    x [N,CX,h,w], ctx_a / ctx_c [N,CCTX,*,*], weight [COUT, 2*CCTX+CX(,1,1)] -> [N,COUT,h,w]."""
    x, ctx_a, ctx_c = _f32c(x, "x"), _f32c(ctx_a, "ctx_a"), _f32c(ctx_c, "ctx_c")
    N, CX, h, w = x.shape
    CCTX = ctx_a.shape[1]
    COUT = weight.shape[0]
    wt = _f32c(weight.reshape(COUT, -1), "weight")
    if ctx_c.shape[1] != CCTX or wt.shape[1] != 2 * CCTX + CX or ctx_a.shape[0] != N or ctx_c.shape[0] != N:
        raise ValueError("context_head: inconsistent shapes")
    out = torch.empty((N, COUT, h, w), device=x.device, dtype=torch.float32)
    with _guard(x), _timed("context_head", 1):
        _check(lib().adamvs_context_head_f32(_p(x), _p(ctx_a), _p(ctx_c), _p(wt), _p(out), N, CX, CCTX, COUT, h, w,
                                             ctx_a.shape[2], ctx_a.shape[3], ctx_c.shape[2], ctx_c.shape[3], _stream()),
               "context_head")
    return out


def deconv3x3_supported(cin: int, cout: int) -> bool:
    return bool(lib().adamvs_deconv3x3_supported(int(cin), int(cout)))


def pack_deconv3x3_weight(w: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d weight [CIN,COUT,3,3] -> [CIN,9,COUT]."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], 9, w.shape[1]).contiguous()


def deconv3x3(x: torch.Tensor, wpk: torch.Tensor, bias: torch.Tensor, relu: bool, residual: Optional[torch.Tensor] = None):
    """act(ConvTranspose2d(k=3, s=2, p=1, output_padding=1)(x) + bias) [+ residual]: x [N,CIN,h,w] -> [N,COUT,2h,2w]."""
    x = _f32c(x, "x")
    N, CIN, h, w = x.shape
    COUT = wpk.shape[2]
    out = torch.empty((N, COUT, 2 * h, 2 * w), device=x.device, dtype=torch.float32)
    if residual is not None:
        residual = _f32c(residual, "residual")
        assert tuple(residual.shape) == tuple(out.shape), (residual.shape, out.shape)
    with _guard(x), _timed("deconv3x3", 1):
        _check(lib().adamvs_deconv3x3_res_f32(_p(x), _p(_f32c(wpk, "wpk")), _p(_f32c(bias, "bias")), int(relu), _p(residual), _p(out),
                                              N, CIN, COUT, h, w, _stream()), "deconv3x3")
    return out


def context_pool_supported(c: int, co: int) -> bool:
    return bool(lib().adamvs_context_pool_supported(int(c), int(co)))


def context_pool(x: torch.Tensor, wa: torch.Tensor, ba: torch.Tensor, wc: torch.Tensor, bc: torch.Tensor):
    """relu(conv1x1(avgpool4(x), wa) + ba), relu(conv1x1(avgpool8(x), wc) + bc): x [N,C,h,w]; wa, wc [CO,C(,1,1)]."""
    x = _f32c(x, "x")
    N, C, h, w = x.shape
    CO = wa.shape[0]
    a = torch.empty((N, CO, h // 4, w // 4), device=x.device, dtype=torch.float32)
    c = torch.empty((N, CO, h // 8, w // 8), device=x.device, dtype=torch.float32)
    with _guard(x), _timed("context_pool", 1):
        _check(lib().adamvs_context_pool_f32(_p(x), _p(_f32c(wa.reshape(CO, C), "wa")), _p(_f32c(ba, "ba")),
                                             _p(_f32c(wc.reshape(CO, C), "wc")), _p(_f32c(bc, "bc")), _p(a), _p(c),
                                             N, C, CO, h, w, _stream()), "context_pool")
    return a, c


def polyphase_5x5_s2_weight(w: torch.Tensor) -> torch.Tensor:
    """A 5x5 stride-2 padding-2 convolution equals a 3x3 stride-1 padding-1 convolution over the pixel-unshuffled
    input (channel order c*4 + py*2 + px, as F.pixel_unshuffle): tap ky = 2*dy + 2 + py of the original kernel sits
    at offset dy of phase py (and likewise in x); the 11 of 36 positions without a counterpart are zero.
    w [COUT,CIN,5,5] -> [COUT,4*CIN,3,3]."""
    cout, cin = w.shape[:2]
    out = w.new_zeros(cout, cin, 2, 2, 3, 3)
    for py in range(2):
        for dy in range(-1, 2):
            ky = 2 * dy + 2 + py
            if not 0 <= ky < 5:
                continue
            for px in range(2):
                for dx in range(-1, 2):
                    kx = 2 * dx + 2 + px
                    if 0 <= kx < 5:
                        out[:, :, py, px, dy + 1, dx + 1] = w[:, :, ky, kx]
    return out.reshape(cout, cin * 4, 3, 3)
