"""autograd binding of the op — mirrors
/root/reference/projects/mmdet3d_plugin/ops/deformable_aggregation.py:7-75."""
import torch
from torch.autograd.function import Function, once_differentiable

from .. import cabi


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
        output = cabi.forward(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                              weights)
        ctx.save_for_backward(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                              weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights = ctx.saved_tensors
        # one memset (grad_feat) instead of the reference's three zeros_like (:55-57): the kernel
        # writes the two small gradients in full
        grad_feat, grad_loc, grad_w = cabi.backward(
            mc_ms_feat, spatial_shape, scale_start_index, sampling_location, weights,
            grad_output.contiguous().float())
        if mc_ms_feat.dtype != torch.float32:
            grad_feat = grad_feat.to(mc_ms_feat.dtype)
        return grad_feat, None, None, grad_loc, grad_w
