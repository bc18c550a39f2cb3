"""ctypes binding of libdfa_b200.so (the C ABI declared in include/dfa_b200.h).

This is the host side of the product path: PyTorch supplies device memory and the current
stream, the library does the work.  There is no fallback: if the shared library is missing the
import raises, and every non-zero return code raises `DfaError`.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfa_b200.so")

F32, BF16 = 0, 1
BWD_ACCUMULATE, BWD_OVERWRITE_SMALL, BWD_ZERO_GRAD_FEAT = 0, 1, 2

# every symbol include/dfa_b200.h declares (tests check the library exports each of them)
SYMBOLS = ("dfa_version", "dfa_error_string", "dfa_debug_reload_knobs", "dfa_forward", "dfa_backward", "dfa_forward_fused",
           "dfa_debug_indices",
           "dfa_flatten_maps", "dfa_keypoints_project", "dfa_keypoints_project_backward",
           "dfa_softmax_weights", "dfa_softmax_weights_backward", "dfa_softmax_weights_split",
           "dfa_softmax_weights_split_backward",
           "dfa_msda_forward", "dfa_msda_backward", "dfa_msda_forward_raw",
           "dfa_forward_host_workspace_bytes", "dfa_forward_host", "dfa_forward_host_stats")


class DfaError(RuntimeError):
    pass


class Dims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("batch_size", "num_cams", "num_feat", "num_embeds",
                                              "num_scale", "num_anchors", "num_pts", "num_groups")]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found — build it with `python -m simpb_b200.build` "
                          "(the CUDA path has no fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    dp = ctypes.POINTER(Dims)
    lib.dfa_version.restype = i32
    lib.dfa_error_string.restype = ctypes.c_char_p
    lib.dfa_error_string.argtypes = [i32]
    lib.dfa_debug_reload_knobs.restype = None
    lib.dfa_debug_reload_knobs.argtypes = []
    lib.dfa_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp, dp, vp]
    lib.dfa_backward.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, dp, i32, vp]
    lib.dfa_forward_fused.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, dp, vp]
    lib.dfa_forward_fused.restype = i32
    lib.dfa_debug_indices.argtypes = [vp, vp, vp, vp, vp, dp, vp]
    lib.dfa_flatten_maps.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32, vp]
    lib.dfa_keypoints_project.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.dfa_keypoints_project_backward.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32,
                                                   i32, vp]
    f32 = ctypes.c_float
    lib.dfa_softmax_weights.argtypes = [vp, vp, f32, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.dfa_softmax_weights_backward.argtypes = [vp, vp, f32, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.dfa_softmax_weights_split.argtypes = [vp, vp, vp, f32, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.dfa_softmax_weights_split_backward.argtypes = [vp, vp, vp, f32, vp, vp, vp, i32, i32, i32, i32,
                                                       i32, i32, vp]
    lib.dfa_msda_forward.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32,
                                     vp, vp]
    lib.dfa_msda_backward.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32,
                                      i32, i32, i32, vp, i32, vp]
    lib.dfa_msda_forward_raw.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32,
                                         vp, vp]
    lib.dfa_forward_host_workspace_bytes.restype = i64
    lib.dfa_forward_host_workspace_bytes.argtypes = [i32, dp]
    lib.dfa_forward_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, dp, vp, i64, vp]
    lib.dfa_forward_host_stats.argtypes = [vp, i32, dp, vp, vp, vp, vp]
    for name in ("dfa_forward", "dfa_backward", "dfa_debug_indices", "dfa_flatten_maps",
                 "dfa_keypoints_project", "dfa_keypoints_project_backward", "dfa_softmax_weights",
                 "dfa_softmax_weights_backward", "dfa_softmax_weights_split",
                 "dfa_softmax_weights_split_backward", "dfa_msda_forward", "dfa_msda_backward", "dfa_msda_forward_raw",
                 "dfa_forward_host", "dfa_forward_host_stats", "dfa_msda_forward_raw"):
        getattr(lib, name).restype = i32
    return lib


lib = _load()


def reload_knobs():
    """Tests / tools: the library caches its DFA_* tuning environment variables on first use; call this
    after changing one of them."""
    lib.dfa_debug_reload_knobs()


def check(rc, what):
    if rc != 0:
        raise DfaError("%s failed: %s (code %d)" % (what, lib.dfa_error_string(rc).decode(), rc))


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def feat_dtype(t):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise DfaError("mc_ms_feat must be float32 or bfloat16, got %s" % t.dtype)


def _need(t, name, dtype=None, cuda=True):
    if not isinstance(t, torch.Tensor):
        raise DfaError("%s must be a tensor" % name)
    if cuda and not t.is_cuda:
        raise DfaError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if not t.is_contiguous():
        raise DfaError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise DfaError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def op_dims(feat, shape, start, loc, w):
    """Same size extraction as deformable_aggregation.cpp:40-47, plus the shape checks the
    reference omits."""
    if feat.dim() != 3 or shape.dim() != 3 or shape.shape[2] != 2 or start.dim() != 2 \
            or loc.dim() != 5 or loc.shape[4] != 2 or w.dim() != 6:
        raise DfaError("deformable_aggregation: wrong tensor ranks")
    bs, num_feat, C = feat.shape
    K, L = shape.shape[:2]
    A, P = loc.shape[1:3]
    G = w.shape[5]
    if tuple(start.shape) != (K, L) or tuple(loc.shape) != (bs, A, P, K, 2) \
            or tuple(w.shape) != (bs, A, P, K, L, G):
        raise DfaError("deformable_aggregation: inconsistent shapes feat=%s shape=%s start=%s "
                       "loc=%s weights=%s" % (tuple(feat.shape), tuple(shape.shape),
                                              tuple(start.shape), tuple(loc.shape), tuple(w.shape)))
    return Dims(bs, K, num_feat, C, L, A, P, G)


def forward(feat, shape, start, loc, w, out=None):
    _need(feat, "mc_ms_feat"); _need(shape, "spatial_shape", torch.int32)
    _need(start, "scale_start_index", torch.int32)
    _need(loc, "sampling_location", torch.float32); _need(w, "weights", torch.float32)
    d = op_dims(feat, shape, start, loc, w)
    if out is None:
        out = torch.empty(d.batch_size, d.num_anchors, d.num_embeds, device=feat.device,
                          dtype=torch.float32)
    if d.batch_size == 0 or d.num_anchors == 0:      # nothing to aggregate (the reference launches 0 threads)
        return out
    with torch.cuda.device(feat.device):
        check(lib.dfa_forward(feat.data_ptr(), feat_dtype(feat), shape.data_ptr(), start.data_ptr(),
                              loc.data_ptr(), w.data_ptr(), out.data_ptr(), ctypes.byref(d),
                              stream_ptr(feat.device)), "dfa_forward")
    return out


def backward(feat, shape, start, loc, w, grad_out, grad_feat=None, grad_loc=None, grad_w=None,
             flags=None, need_feat=True):
    """With no buffers given: allocates them, lets the kernel write the two small gradients in full
    and zero-fills grad_feat on the stream (one memset instead of the reference's three).  With every
    buffer given (and no flags): the reference contract, accumulate into caller-zeroed buffers.  A
    caller that passes only SOME of the buffers must say what it wants with `flags` — guessing could
    overwrite a buffer it expected to be accumulated into.
    need_feat=False skips the feature gradient (scatter and zero-fill) and returns None for it."""
    _need(feat, "mc_ms_feat"); _need(shape, "spatial_shape", torch.int32)
    _need(start, "scale_start_index", torch.int32)
    _need(loc, "sampling_location", torch.float32); _need(w, "weights", torch.float32)
    _need(grad_out, "grad_output", torch.float32)
    d = op_dims(feat, shape, start, loc, w)
    if tuple(grad_out.shape) != (d.batch_size, d.num_anchors, d.num_embeds):
        raise DfaError("grad_output must be [bs, anchors, embeds]")
    if flags is None:
        all_given = (grad_feat is not None or not need_feat) and grad_loc is not None and grad_w is not None
        none_given = grad_feat is None and grad_loc is None and grad_w is None
        if not (all_given or none_given):
            raise DfaError("dfa_backward: some gradient buffers were passed and some were not; "
                           "pass `flags` (BWD_ACCUMULATE or BWD_OVERWRITE_SMALL [| BWD_ZERO_GRAD_FEAT])")
        flags = BWD_ACCUMULATE if all_given else BWD_OVERWRITE_SMALL
        if grad_feat is None and need_feat:
            grad_feat = torch.empty(feat.shape, device=feat.device, dtype=torch.float32)
            flags |= BWD_ZERO_GRAD_FEAT
    grad_loc = torch.empty_like(loc) if grad_loc is None else grad_loc
    grad_w = torch.empty_like(w) if grad_w is None else grad_w
    if not need_feat:
        grad_feat = None
    for t, n in ((grad_feat, "grad_mc_ms_feat"), (grad_loc, "grad_sampling_location"),
                 (grad_w, "grad_weights")):
        if t is not None:
            _need(t, n, torch.float32)
    if (grad_feat is not None and grad_feat.numel() != feat.numel()) or grad_loc.numel() != loc.numel() \
            or grad_w.numel() != w.numel():
        raise DfaError("gradient buffer sizes do not match their inputs")
    if d.batch_size == 0 or d.num_anchors == 0:      # empty batch / no anchors: no contribution
        if grad_feat is not None and (flags & BWD_ZERO_GRAD_FEAT):
            grad_feat.zero_()
        return grad_feat, grad_loc, grad_w
    with torch.cuda.device(feat.device):
        check(lib.dfa_backward(feat.data_ptr(), feat_dtype(feat), shape.data_ptr(), start.data_ptr(),
                               loc.data_ptr(), w.data_ptr(), grad_out.data_ptr(),
                               grad_feat.data_ptr() if grad_feat is not None else None,
                               grad_loc.data_ptr(), grad_w.data_ptr(),
                               ctypes.byref(d), int(flags), stream_ptr(feat.device)), "dfa_backward")
    return grad_feat, grad_loc, grad_w


def debug_indices(shape, start, loc):
    _need(shape, "spatial_shape", torch.int32); _need(start, "scale_start_index", torch.int32)
    _need(loc, "sampling_location", torch.float32)
    bs, A, P, K, _ = loc.shape
    L = shape.shape[1]
    d = Dims(bs, K, 1, 1, L, A, P, 1)
    valid = torch.empty(bs, A, P, K, device=loc.device, dtype=torch.uint8)
    rows = torch.empty(bs, A, P, K, L, 4, device=loc.device, dtype=torch.int32)
    with torch.cuda.device(loc.device):
        check(lib.dfa_debug_indices(shape.data_ptr(), start.data_ptr(), loc.data_ptr(),
                                    valid.data_ptr(), rows.data_ptr(), ctypes.byref(d),
                                    stream_ptr(loc.device)), "dfa_debug_indices")
    return valid, rows


def flatten_maps(maps, out_dtype=torch.float32, out=None):
    """maps: list over levels of contiguous float32 CUDA tensors [bs, K, C, H_l, W_l]."""
    L = len(maps)
    bs, K, C = maps[0].shape[:3]
    for m in maps:
        _need(m, "feature map", torch.float32)
        if m.dim() != 5 or tuple(m.shape[:3]) != (bs, K, C):
            raise DfaError("feature maps must all be [bs, cams, C, H_l, W_l]")
    rows = K * sum(int(m.shape[3]) * int(m.shape[4]) for m in maps)
    if out is None:
        out = torch.empty(bs, rows, C, device=maps[0].device, dtype=out_dtype)
    ptrs = (ctypes.c_void_p * L)(*[m.data_ptr() for m in maps])
    hw = (ctypes.c_int32 * (2 * L))(*[int(v) for m in maps for v in m.shape[3:5]])
    with torch.cuda.device(out.device):
        check(lib.dfa_flatten_maps(ptrs, hw, L, bs, K, C, out.data_ptr(), feat_dtype(out),
                                   stream_ptr(out.device)), "dfa_flatten_maps")
    return out


def keypoints_project(anchor, fix_scale, learnable_logits, projection_mat, image_wh,
                      want_key_points=False):
    _need(anchor, "anchor", torch.float32); _need(fix_scale, "fix_scale", torch.float32)
    _need(projection_mat, "projection_mat", torch.float32)
    bs, A = anchor.shape[:2]
    if anchor.shape[2] != 11:
        raise DfaError("anchor must be [bs, A, 11]")
    F_ = fix_scale.shape[0]
    n_learn = 0
    if learnable_logits is not None:
        _need(learnable_logits, "learnable_logits", torch.float32)
        n_learn = learnable_logits.numel() // (bs * A * 3)
    P, K = F_ + n_learn, projection_mat.shape[1]
    if image_wh is not None:
        _need(image_wh, "image_wh", torch.float32)
    kp = torch.empty(bs, A, P, 3, device=anchor.device) if want_key_points else None
    loc = torch.empty(bs, A, P, K, 2, device=anchor.device)
    with torch.cuda.device(anchor.device):
        check(lib.dfa_keypoints_project(
            anchor.data_ptr(), fix_scale.data_ptr(), F_,
            learnable_logits.data_ptr() if learnable_logits is not None else None,
            projection_mat.data_ptr(), image_wh.data_ptr() if image_wh is not None else None,
            kp.data_ptr() if kp is not None else None, loc.data_ptr(), bs, A, P, K,
            stream_ptr(anchor.device)), "dfa_keypoints_project")
    return (loc, kp) if want_key_points else loc


UNSUPPORTED = -5   # DFA_ERR_UNSUPPORTED


def forward_fused(feat, shape, start, anchor, fix_scale, learnable_logits, projection_mat, image_wh,
                  logits_anchor, logits_cam, dims6, want_locations=False):
    """The module's forward between its Linear layers in one launch (inference).  dims6 =
    (bs, A, K, L, P, G).  Returns the aggregated features [bs,A,C] (and the sampling locations), or
    None when the shape is outside the fused kernel's fast path (the caller then uses the separate
    kernels)."""
    _need(feat, "mc_ms_feat"); _need(shape, "spatial_shape", torch.int32)
    _need(start, "scale_start_index", torch.int32)
    for t, n in ((anchor, "anchor"), (fix_scale, "fix_scale"), (projection_mat, "projection_mat"),
                 (logits_anchor, "logits_anchor")):
        _need(t, n, torch.float32)
    for t, n in ((learnable_logits, "learnable_logits"), (image_wh, "image_wh"), (logits_cam, "logits_cam")):
        if t is not None:
            _need(t, n, torch.float32)
    bs, A, K, L, P, G = dims6
    lpg = L * P * G
    if feat.dim() != 3 or feat.shape[0] != bs or tuple(anchor.shape) != (bs, A, 11) \
            or tuple(shape.shape) != (K, L, 2) or projection_mat.numel() != bs * K * 16:
        raise DfaError("forward_fused: inconsistent shapes")
    if logits_anchor.numel() != (bs * A * lpg if logits_cam is not None else bs * A * K * lpg) \
            or (logits_cam is not None and logits_cam.numel() != bs * K * lpg):
        raise DfaError("forward_fused: logits must be [bs,A,L*P*G] + [bs,K,L*P*G] or [bs,A,K,L*P*G]")
    F_ = fix_scale.shape[0]
    n_learn = 0 if learnable_logits is None else learnable_logits.numel() // max(bs * A * 3, 1)
    if F_ + n_learn != P:
        raise DfaError("forward_fused: fixed + learnable key points != num_pts")
    d = Dims(bs, K, feat.shape[1], feat.shape[2], L, A, P, G)
    out = torch.empty(bs, A, feat.shape[2], device=feat.device, dtype=torch.float32)
    loc = torch.empty(bs, A, P, K, 2, device=feat.device) if want_locations else None
    if bs == 0 or A == 0:
        return (out, loc) if want_locations else out
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    with torch.cuda.device(feat.device):
        rc = lib.dfa_forward_fused(feat.data_ptr(), feat_dtype(feat), shape.data_ptr(), start.data_ptr(),
                                   anchor.data_ptr(), fix_scale.data_ptr(), F_, ptr(learnable_logits),
                                   projection_mat.data_ptr(), ptr(image_wh), logits_anchor.data_ptr(),
                                   ptr(logits_cam), out.data_ptr(), ptr(loc), ctypes.byref(d),
                                   stream_ptr(feat.device))
    if rc == UNSUPPORTED:
        return None
    check(rc, "dfa_forward_fused")
    return (out, loc) if want_locations else out


def _kp_dims(anchor, fix_scale, learnable_logits, projection_mat):
    bs, A = anchor.shape[:2]
    if anchor.dim() != 3 or anchor.shape[2] != 11:
        raise DfaError("anchor must be [bs, A, 11]")
    F_ = fix_scale.shape[0]
    n_learn = 0 if learnable_logits is None else learnable_logits.numel() // (bs * A * 3)
    return bs, A, F_, F_ + n_learn, projection_mat.shape[1]


def keypoints_project_backward(anchor, fix_scale, learnable_logits, projection_mat, image_wh, grad_loc):
    """Gradients of keypoints_project wrt the anchor and the learnable-offset logits."""
    _need(anchor, "anchor", torch.float32); _need(fix_scale, "fix_scale", torch.float32)
    _need(projection_mat, "projection_mat", torch.float32)
    _need(grad_loc, "grad_sampling_location", torch.float32)
    if learnable_logits is not None:
        _need(learnable_logits, "learnable_logits", torch.float32)
    if image_wh is not None:
        _need(image_wh, "image_wh", torch.float32)
    bs, A, F_, P, K = _kp_dims(anchor, fix_scale, learnable_logits, projection_mat)
    if tuple(grad_loc.shape) != (bs, A, P, K, 2):
        raise DfaError("grad_sampling_location must be [bs, A, P, K, 2]")
    g_anchor = torch.empty_like(anchor)
    g_logits = torch.empty_like(learnable_logits) if learnable_logits is not None else None
    with torch.cuda.device(anchor.device):
        check(lib.dfa_keypoints_project_backward(
            anchor.data_ptr(), fix_scale.data_ptr(), F_,
            learnable_logits.data_ptr() if learnable_logits is not None else None,
            projection_mat.data_ptr(), image_wh.data_ptr() if image_wh is not None else None,
            grad_loc.data_ptr(), g_anchor.data_ptr(),
            g_logits.data_ptr() if g_logits is not None else None, bs, A, P, K,
            stream_ptr(anchor.device)), "dfa_keypoints_project_backward")
    return g_anchor, g_logits


def _softmax_args(logits, dims, keep):
    _need(logits, "logits", torch.float32)
    bs, A, K, L, P, G = dims
    if logits.numel() != bs * A * K * L * P * G:
        raise DfaError("logits must hold bs*A*K*L*P*G elements")
    if keep is not None:
        _need(keep, "keep_mask", torch.uint8)
        if keep.numel() != bs * A * K * P:
            raise DfaError("keep_mask must be uint8 [bs, A, K, P]")


def softmax_weights(logits, dims, keep=None, scale=1.0):
    """logits in (bs, A, K, L, P, G) memory order → weights [bs, A, P, K, L, G]; dims = that tuple."""
    _softmax_args(logits, dims, keep)
    bs, A, K, L, P, G = dims
    w = torch.empty(bs, A, P, K, L, G, device=logits.device, dtype=torch.float32)
    with torch.cuda.device(logits.device):
        check(lib.dfa_softmax_weights(logits.data_ptr(), keep.data_ptr() if keep is not None else None,
                                      float(scale), w.data_ptr(), bs, A, K, L, P, G,
                                      stream_ptr(logits.device)), "dfa_softmax_weights")
    return w


def softmax_weights_backward(logits, dims, keep, scale, grad_w):
    _softmax_args(logits, dims, keep)
    _need(grad_w, "grad_weights", torch.float32)
    bs, A, K, L, P, G = dims
    if grad_w.numel() != logits.numel():
        raise DfaError("grad_weights must be [bs, A, P, K, L, G]")
    g = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        check(lib.dfa_softmax_weights_backward(
            logits.data_ptr(), keep.data_ptr() if keep is not None else None, float(scale),
            grad_w.data_ptr(), g.data_ptr(), bs, A, K, L, P, G, stream_ptr(logits.device)),
            "dfa_softmax_weights_backward")
    return g


def _split_args(logits_anchor, logits_cam, dims, keep):
    _need(logits_anchor, "logits_anchor", torch.float32); _need(logits_cam, "logits_cam", torch.float32)
    bs, A, K, L, P, G = dims
    if logits_anchor.numel() != bs * A * L * P * G or logits_cam.numel() != bs * K * L * P * G:
        raise DfaError("split logits must be [bs, A, L*P*G] and [bs, K, L*P*G]")
    if keep is not None:
        _need(keep, "keep_mask", torch.uint8)
        if keep.numel() != bs * A * K * P:
            raise DfaError("keep_mask must be uint8 [bs, A, K, P]")


def softmax_weights_split(logits_anchor, logits_cam, dims, keep=None, scale=1.0):
    """weights [bs,A,P,K,L,G] from logits_anchor [bs,A,L*P*G] + logits_cam [bs,K,L*P*G]."""
    _split_args(logits_anchor, logits_cam, dims, keep)
    bs, A, K, L, P, G = dims
    w = torch.empty(bs, A, P, K, L, G, device=logits_anchor.device, dtype=torch.float32)
    with torch.cuda.device(w.device):
        check(lib.dfa_softmax_weights_split(
            logits_anchor.data_ptr(), logits_cam.data_ptr(), keep.data_ptr() if keep is not None else None,
            float(scale), w.data_ptr(), bs, A, K, L, P, G, stream_ptr(w.device)), "dfa_softmax_weights_split")
    return w


def softmax_weights_split_backward(logits_anchor, logits_cam, dims, keep, scale, grad_w):
    """Returns (grad_logits_anchor [bs,A,L*P*G], grad_logits_cam [bs,K,L*P*G])."""
    _split_args(logits_anchor, logits_cam, dims, keep)
    _need(grad_w, "grad_weights", torch.float32)
    bs, A, K, L, P, G = dims
    g_full = torch.empty(bs, A, K, L * P * G, device=grad_w.device, dtype=torch.float32)
    g_anchor = torch.empty_like(logits_anchor)
    with torch.cuda.device(grad_w.device):
        check(lib.dfa_softmax_weights_split_backward(
            logits_anchor.data_ptr(), logits_cam.data_ptr(), keep.data_ptr() if keep is not None else None,
            float(scale), grad_w.data_ptr(), g_full.data_ptr(), g_anchor.data_ptr(), bs, A, K, L, P, G,
            stream_ptr(grad_w.device)), "dfa_softmax_weights_split_backward")
    return g_anchor, g_full.sum(dim=1).reshape(logits_cam.shape)


def _msda_dims(value, shapes, start, loc, w, query_table):
    _need(value, "value"); _need(shapes, "spatial_shapes", torch.int32)
    _need(start, "level_start_index", torch.int32)
    _need(loc, "sampling_locations", torch.float32); _need(w, "attention_weights", torch.float32)
    grouped = query_table is not None
    if value.dim() != (5 if grouped else 4) or loc.dim() != 6 or loc.shape[5] != 2 or w.dim() != 5 \
            or shapes.dim() != 2:
        raise DfaError("msda: value [bs,(K,)S,M,D], loc [bs,Q,M,L,P,2], weights [bs,Q,M,L,P] expected")
    bs = value.shape[0]
    K = value.shape[1] if grouped else 1
    S, M, D = value.shape[-3:]
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    if tuple(loc.shape) != (bs, Q, M, L, P, 2) or tuple(w.shape) != (bs, Q, M, L, P) \
            or tuple(shapes.shape) != (L, 2) or start.numel() != L:
        raise DfaError("msda: inconsistent shapes value=%s loc=%s weights=%s shapes=%s"
                       % (tuple(value.shape), tuple(loc.shape), tuple(w.shape), tuple(shapes.shape)))
    if grouped:
        _need(query_table, "query_table", torch.int32)
        if query_table.numel() != Q:
            raise DfaError("msda: query_table must hold one table index per query")
    return bs, S, M, D, Q, L, P, K


def msda_forward(value, shapes, start, loc, w, query_table=None):
    """value [bs,S,M,D], or [bs,K,S,M,D] with query_table int32 [Q] (table sampled by each query)."""
    bs, S, M, D, Q, L, P, K = _msda_dims(value, shapes, start, loc, w, query_table)
    out = torch.empty(bs, Q, M * D, device=value.device, dtype=torch.float32)
    if bs == 0 or Q == 0:
        return out
    with torch.cuda.device(value.device):
        check(lib.dfa_msda_forward(value.data_ptr(), feat_dtype(value), shapes.data_ptr(), start.data_ptr(),
                                   loc.data_ptr(), w.data_ptr(), out.data_ptr(), bs, S, M, D, Q, L, P, K,
                                   query_table.data_ptr() if query_table is not None else None,
                                   stream_ptr(value.device)), "dfa_msda_forward")
    return out


def msda_forward_raw(table, shapes, start, loc, w, query_table=None):
    """Gather-then-project variant (inference).  table [bs,S,C] or [bs,K,S,C] (UNPROJECTED rows of 512 or
    1024 bytes), loc [bs,Q,M,L,P,2], w [bs,Q,M,L,P]; returns (gathered [bs,Q,M,C], weight_sum [bs,Q,M]) —
    the caller applies value_proj per head afterwards (include/dfa_b200.h)."""
    _need(table, "table")
    _need(loc, "sampling_locations", torch.float32)
    _need(w, "attention_weights", torch.float32)
    if table.dim() not in (3, 4) or loc.dim() != 6 or loc.shape[-1] != 2 or w.shape != loc.shape[:-1]:
        raise DfaError("msda_forward_raw: table [bs,(K,)S,C], loc [bs,Q,M,L,P,2], w [bs,Q,M,L,P]")
    K = table.shape[1] if table.dim() == 4 else 1
    bs, S, C = table.shape[0], table.shape[-2], table.shape[-1]
    _, Q, M, L, P, _ = loc.shape
    if loc.shape[0] != bs or shapes.numel() != 2 * L or start.numel() != L:
        raise DfaError("msda_forward_raw: inconsistent batch size or level tables")
    if K > 1 and (query_table is None or query_table.numel() != Q or query_table.dtype != torch.int32):
        raise DfaError("msda_forward_raw: query_table int32 [Q] is required with several tables")
    g = torch.empty(bs, Q, M, C, device=table.device, dtype=torch.float32)
    ssum = torch.empty(bs, Q, M, device=table.device, dtype=torch.float32)
    if bs == 0 or Q == 0:
        return g, ssum
    with torch.cuda.device(table.device):
        check(lib.dfa_msda_forward_raw(table.data_ptr(), feat_dtype(table), shapes.data_ptr(), start.data_ptr(),
                                       loc.data_ptr(), w.data_ptr(), g.data_ptr(), ssum.data_ptr(), bs, S, C, M,
                                       Q, L, P, K, query_table.data_ptr() if query_table is not None else None,
                                       stream_ptr(table.device)), "dfa_msda_forward_raw")
    return g, ssum


def msda_backward(value, shapes, start, loc, w, grad_out, need_value=True, query_table=None):
    bs, S, M, D, Q, L, P, K = _msda_dims(value, shapes, start, loc, w, query_table)
    _need(grad_out, "grad_output", torch.float32)
    if grad_out.numel() != bs * Q * M * D:
        raise DfaError("msda: grad_output must be [bs, Q, M*D]")
    gv = torch.empty(value.shape, device=value.device, dtype=torch.float32) if need_value else None
    gl, gw = torch.empty_like(loc), torch.empty_like(w)
    if bs == 0 or Q == 0:
        if gv is not None:
            gv.zero_()
        return gv, gl, gw
    with torch.cuda.device(value.device):
        check(lib.dfa_msda_backward(value.data_ptr(), feat_dtype(value), shapes.data_ptr(), start.data_ptr(),
                                    loc.data_ptr(), w.data_ptr(), grad_out.data_ptr(),
                                    gv.data_ptr() if gv is not None else None, gl.data_ptr(), gw.data_ptr(),
                                    bs, S, M, D, Q, L, P, K,
                                    query_table.data_ptr() if query_table is not None else None, 1,
                                    stream_ptr(value.device)), "dfa_msda_backward")
    return gv, gl, gw


class HostForward:
    """End-to-end forward with HOST buffers through dfa_forward_host: the inputs' trip to the device,
    the kernel and the device→host copy of the result all happen inside the call.  Pinned (mapped)
    host tensors take the pull mode — only the rows and weight lines the forward reads cross the link;
    pageable ones are copied whole (include/dfa_b200.h).  stats() = what the last call moved."""

    def __init__(self, dims, dtype=torch.float32, device="cuda"):
        self.dims = dims
        self.dt = F32 if dtype == torch.float32 else BF16
        n = lib.dfa_forward_host_workspace_bytes(self.dt, ctypes.byref(dims))
        if n < 0:
            raise DfaError("bad dims for dfa_forward_host")
        self.nbytes = int(n)
        self.workspace = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.device = self.workspace.device

    def __call__(self, h_feat, h_shape, h_start, h_loc, h_w, h_out):
        for t, n in ((h_feat, "feat"), (h_shape, "shape"), (h_start, "start"), (h_loc, "loc"),
                     (h_w, "weights"), (h_out, "out")):
            _need(t, n, cuda=False)
            if t.is_cuda:
                raise DfaError("dfa_forward_host takes host tensors")
        with torch.cuda.device(self.device):
            check(lib.dfa_forward_host(h_feat.data_ptr(), self.dt, h_shape.data_ptr(),
                                       h_start.data_ptr(), h_loc.data_ptr(), h_w.data_ptr(),
                                       h_out.data_ptr(), ctypes.byref(self.dims),
                                       self.workspace.data_ptr(), self.nbytes,
                                       stream_ptr(self.device)), "dfa_forward_host")
        return h_out

    def stats(self):
        """(host→device bytes, feature rows, weight bytes) moved by the last call."""
        v = [ctypes.c_int64(0) for _ in range(3)]
        with torch.cuda.device(self.device):
            check(lib.dfa_forward_host_stats(self.workspace.data_ptr(), self.dt, ctypes.byref(self.dims),
                                             stream_ptr(self.device), ctypes.byref(v[0]), ctypes.byref(v[1]),
                                             ctypes.byref(v[2])), "dfa_forward_host_stats")
        return tuple(int(x.value) for x in v)
