"""Multi-scale deformable attention of SimPB's 2-D query branch on libdfa_b200.

`MultiScaleDeformableAttnFunction` keeps the call signature of mmcv-full 1.7.1's function of the
same name, which the reference applies once per camera group at
/root/reference/projects/mmdet3d_plugin/models/group_attn.py:229-233;
`QueryGroupMultiScaleDeformableAttention` mirrors the module around it (:136-256, built on mmcv's
`MultiScaleDeformableAttention`): same constructor arguments, parameter names and forward
signature.  The value tensor is consumed in place as `[bs, S, heads, head_dim]` — for SimPB that is a
view of one camera's rows of the same channel-last table the 3-D branch gathers from — in float32
or bfloat16.
"""
import math
import os

import torch
import torch.nn as nn
from torch.autograd.function import Function, once_differentiable

from . import cabi

__all__ = ["MultiScaleDeformableAttnFunction", "QueryGroupMultiScaleDeformableAttention"]


def _i32(t):
    return t if t.dtype == torch.int32 and t.is_contiguous() else t.contiguous().int()


class MultiScaleDeformableAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step=64):
        """value [bs,S,M,D], value_spatial_shapes [L,2] (H,W), value_level_start_index [L],
        sampling_locations [bs,Q,M,L,P,2] in [0,1] (x,y), attention_weights [bs,Q,M,L,P]
        → [bs, Q, M*D].  `im2col_step` is accepted for signature compatibility and unused (the
        kernel has no batch tiling)."""
        if value.dtype != torch.bfloat16:
            value = value.float()
        value = value.contiguous()
        shapes, start = _i32(value_spatial_shapes), _i32(value_level_start_index)
        loc = sampling_locations.contiguous().float()
        w = attention_weights.contiguous().float()
        ctx.save_for_backward(value, shapes, start, loc, w)
        return cabi.msda_forward(value, shapes, start, loc, w)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, loc, w = ctx.saved_tensors
        gv, gl, gw = cabi.msda_backward(value, shapes, start, loc, w, grad_output.contiguous().float(),
                                        need_value=ctx.needs_input_grad[0])
        if gv is not None and value.dtype != torch.float32:
            gv = gv.to(value.dtype)
        return gv, None, None, gl, gw, None


class GroupedMultiScaleDeformableAttnFunction(Function):
    """All camera groups in one launch: value [bs, K, S, M, D], query_table int32 [Q] names the
    table each query samples (the reference calls mmcv once per group and concatenates)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, query_table):
        if value.dtype != torch.bfloat16:
            value = value.float()
        value = value.contiguous()
        shapes, start = _i32(value_spatial_shapes), _i32(value_level_start_index)
        loc = sampling_locations.contiguous().float()
        w = attention_weights.contiguous().float()
        query_table = _i32(query_table)
        ctx.save_for_backward(value, shapes, start, loc, w, query_table)
        return cabi.msda_forward(value, shapes, start, loc, w, query_table)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, loc, w, query_table = ctx.saved_tensors
        gv, gl, gw = cabi.msda_backward(value, shapes, start, loc, w, grad_output.contiguous().float(),
                                        need_value=ctx.needs_input_grad[0], query_table=query_table)
        if gv is not None and value.dtype != torch.float32:
            gv = gv.to(value.dtype)
        return gv, None, None, gl, gw, None


class QueryGroupMultiScaleDeformableAttention(nn.Module):
    """group_attn.py:136-256.  Queries are partitioned into per-camera groups
    (`query_groups[i] = (q0, q1)` samples camera i); value is `[bs*num_cams, S, C]` when
    batch_first else `[S, bs*num_cams, C]`, exactly as the reference head passes it
    (models/simpb_head.py:282-292)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, num_cams=6,
                 query_groups=None, im2col_step=64, dropout=0.1, batch_first=False, norm_cfg=None,
                 init_cfg=None, residual_mode="add"):
        super().__init__()
        if embed_dims % num_heads != 0:
            raise ValueError("embed_dims must be divisible by num_heads, but got %d and %d"
                             % (embed_dims, num_heads))
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.num_levels, self.num_points, self.num_cams = num_levels, num_points, num_cams
        self.query_groups, self.residual_mode = query_groups, residual_mode
        self.im2col_step, self.batch_first, self.norm_cfg = im2col_step, batch_first, norm_cfg
        self.dropout = nn.Dropout(dropout)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        """mmcv's MultiScaleDeformableAttention.init_weights: zero offsets with a ring of biases."""
        nn.init.constant_(self.sampling_offsets.weight, 0.0)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2)
        grid = grid.repeat(1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.constant_(self.attention_weights.weight, 0.0)
        nn.init.constant_(self.attention_weights.bias, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.0)

    def _query_table(self, num_query, device):
        """int32 [num_query] camera index per query when the groups tile the queries in order
        (what the reference's allocation produces), else None.  Cached per (groups, device)."""
        groups = tuple((int(a), int(b)) for a, b in self.query_groups)
        key = (groups, num_query, str(device))
        if getattr(self, "_table_key", None) == key:
            return self._table
        pos, cams = 0, []
        for i, (a, b) in enumerate(groups):
            if b <= a:
                continue
            if a != pos or i >= self.num_cams:
                pos = -1
                break
            cams += [i] * (b - a)
            pos = b
        table = torch.tensor(cams, dtype=torch.int32, device=device) if pos == num_query else None
        self._table_key, self._table = key, table
        return table

    def sampling_locations(self, reference_points, sampling_offsets, spatial_shapes):
        """group_attn.py:191-217."""
        last = reference_points.shape[-1]
        if last in (2, 3):
            norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
            return (reference_points[:, :, None, :, None, :2]
                    + sampling_offsets / norm[None, None, None, :, None, :])
        if last in (4, 5):
            return (reference_points[:, :, None, :, None, :2]
                    + sampling_offsets / self.num_points * reference_points[:, :, None, :, None, 2:4] * 0.5)
        raise ValueError("Last dim of reference_points must be 2 or 4, but get %d instead." % last)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_padding_mask=None, reference_points=None, spatial_shapes=None,
                level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        bcs, num_value, _ = value.shape
        if bcs // self.num_cams != bs:
            raise ValueError("value must hold bs*num_cams = %d items, got %d" % (bs * self.num_cams, bcs))
        # Inference on a table whose rows are 512 or 1024 bytes, queries grouped by camera in order: gather
        # the UNPROJECTED rows per (query, head) and apply value_proj afterwards (sampling is linear) — the
        # [bs*K*S, C] x [C, C] GEMM over the whole table (0.39 ms per layer at SimPB's size) becomes num_heads
        # products over the queries only.  SIMPB_B200_MSDA_PROJECT_FIRST=1 keeps the reference's order.
        raw = None
        if (not torch.is_grad_enabled() and key_padding_mask is None and value.is_cuda
                and value.dtype in (torch.float32, torch.bfloat16)
                and value.shape[-1] * value.element_size() in (512, 1024)
                and os.environ.get("SIMPB_B200_MSDA_PROJECT_FIRST", "0") != "1"):
            raw = value.contiguous().view(bs, self.num_cams, num_value, -1)
        else:
            value = self.value_proj(value)
            if key_padding_mask is not None:
                value = value.masked_fill(key_padding_mask[..., None], 0.0)
            value = value.view(bs, self.num_cams, num_value, self.num_heads, -1)
        offsets = self.sampling_offsets(query).view(bs, num_query, self.num_heads, self.num_levels,
                                                    self.num_points, 2)
        weights = self.attention_weights(query).view(bs, num_query, self.num_heads,
                                                     self.num_levels * self.num_points).softmax(-1)
        weights = weights.view(bs, num_query, self.num_heads, self.num_levels, self.num_points)
        loc = self.sampling_locations(reference_points, offsets, spatial_shapes)
        ref_depth = kwargs.get("ref_depth2d", None)
        if ref_depth is not None:                       # :219-222
            xs, ys, _ = torch.where(ref_depth == 0)
            loc[xs, ys] = 0
        if kwargs.get("query_groups", None) is not None:
            self.query_groups = kwargs["query_groups"]
        table = self._query_table(num_query, value.device)
        if raw is not None and table is None:          # groups out of order: the reference's order after all
            value = self.value_proj(value).view(bs, self.num_cams, num_value, self.num_heads, -1)
            raw = None
        if raw is not None:
            g, ssum = cabi.msda_forward_raw(raw, _i32(spatial_shapes), _i32(level_start_index),
                                            loc.contiguous().float(), weights.contiguous().float(), table)
            M, D = self.num_heads, self.embed_dims // self.num_heads
            wv = self.value_proj.weight.view(M, D, -1)                                  # [m, d, c]
            output = torch.bmm(g.view(bs * num_query, M, -1).transpose(0, 1).to(wv.dtype), wv.transpose(1, 2))
            output = output.transpose(0, 1).reshape(bs, num_query, M, D)
            if self.value_proj.bias is not None:
                output = output + ssum[..., None] * self.value_proj.bias.view(M, D)
            output = output.reshape(bs, num_query, M * D)
        elif table is not None:     # the groups tile [0, num_query) in order: one launch for all cameras
            output = GroupedMultiScaleDeformableAttnFunction.apply(value, spatial_shapes, level_start_index,
                                                                   loc, weights, table)
        else:                       # anything else: group by group, as the reference (:226-236)
            outs = []
            for i, qg in enumerate(self.query_groups):
                if qg[1] - qg[0] > 0:
                    outs.append(MultiScaleDeformableAttnFunction.apply(
                        value[:, i], spatial_shapes, level_start_index, loc[:, qg[0]:qg[1]],
                        weights[:, qg[0]:qg[1]], self.im2col_step))
            output = torch.cat(outs, dim=1)
        output = self.output_proj(output)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        output = self.dropout(output)
        if self.residual_mode == "add":
            output = output + identity
        elif self.residual_mode == "cat":
            output = torch.cat([output, identity], dim=-1)
        return output
