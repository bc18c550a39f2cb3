"""autograd binding of the op — mirrors
/root/reference/projects/mmdet3d_plugin/ops/deformable_aggregation.py:7-75."""
import os

import torch
from torch.autograd.function import Function, once_differentiable

from .. import cabi

# Opt-in (DFA_PREFILL_GRAD_FEAT=1, read once at import): allocate the feature gradient in forward()
# and zero-fill it there on a side stream, so the fill overlaps the forward instead of running in front
# of the backward kernel.  It saves the fill's time on the backward's critical path (about a third of an
# op-level training step at bs=8) but keeps a feature-table-sized fp32 buffer alive per call from
# forward to backward (92 MB per sample and layer at R50) and competes with the forward for HBM —
# hence off by default: like the reference (ops/deformable_aggregation.py:55-57) the buffer then
# lives only inside backward().
PREFILL_GRAD_FEAT = os.environ.get("DFA_PREFILL_GRAD_FEAT", "0") == "1"

_SIDE_STREAMS = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _tables_i32(t):
    # the level tables arrive as int64 from feature_maps_format; cast once (tiny)
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.contiguous().int()


class DeformableAggregationFunction(Function):
    @staticmethod
    def forward(ctx, mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights):
        # Same coercions as the reference (:18-22), except that bfloat16 features are consumed
        # natively instead of being up-cast into a float32 copy.
        if mc_ms_feat.dtype != torch.bfloat16:
            mc_ms_feat = mc_ms_feat.float()
        mc_ms_feat = mc_ms_feat.contiguous()
        spatial_shape = _tables_i32(spatial_shape)
        scale_start_index = _tables_i32(scale_start_index)
        sampling_location = sampling_location.contiguous().float()
        weights = weights.contiguous().float()
        ctx.grad_feat = ctx.grad_feat_ready = None
        if ctx.needs_input_grad[0] and PREFILL_GRAD_FEAT and not torch.cuda.is_current_stream_capturing():
            # grad_mc_ms_feat is a scatter target and must start from zero: zero it NOW on a side
            # stream (see PREFILL_GRAD_FEAT above).  Never inside a stream capture: a forward captured
            # without its backward would leave the side stream un-joined.
            cur, side = torch.cuda.current_stream(mc_ms_feat.device), _side_stream(mc_ms_feat.device)
            ctx.grad_feat = torch.empty(mc_ms_feat.shape, device=mc_ms_feat.device, dtype=torch.float32)
            side.wait_stream(cur)              # the allocator may hand out a block still in use on `cur`
            with torch.cuda.stream(side):
                ctx.grad_feat.zero_()
                ctx.grad_feat_ready = torch.cuda.Event()
                ctx.grad_feat_ready.record(side)
            ctx.grad_feat.record_stream(side)
        output = cabi.forward(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                              weights)
        ctx.save_for_backward(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                              weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights = ctx.saved_tensors
        # one memset (grad_feat, issued by the library on this stream right before the kernel) instead
        # of the reference's three zeros_like (:55-57): the kernel writes the two small gradients in
        # full.  With PREFILL_GRAD_FEAT the fill was started in forward() instead.
        grad_feat, ctx.grad_feat = ctx.grad_feat, None
        if grad_feat is not None:
            torch.cuda.current_stream(mc_ms_feat.device).wait_event(ctx.grad_feat_ready)
            grad_feat, grad_loc, grad_w = cabi.backward(
                mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights,
                grad_output.contiguous().float(), grad_feat=grad_feat,
                grad_loc=torch.empty_like(sampling_location), grad_w=torch.empty_like(weights),
                flags=cabi.BWD_OVERWRITE_SMALL)
        else:   # the default; also frozen features (scatter skipped) and retained graphs
            grad_feat, grad_loc, grad_w = cabi.backward(
                mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights,
                grad_output.contiguous().float(), need_feat=ctx.needs_input_grad[0])
        if grad_feat is not None and mc_ms_feat.dtype != torch.float32:
            grad_feat = grad_feat.to(mc_ms_feat.dtype)
        return grad_feat, None, None, grad_loc, grad_w
