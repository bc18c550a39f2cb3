"""Drop-in for the reference's `projects/mmdet3d_plugin/ops` package: same three public names,
same argument order and meaning (/root/reference/projects/mmdet3d_plugin/ops/__init__.py)."""
import torch

from .. import cabi
from .deformable_aggregation import DeformableAggregationFunction

__all__ = ["DeformableAggregationFunction", "deformable_aggregation_function",
           "feature_maps_format"]


def deformable_aggregation_function(feature_maps, spatial_shape, scale_start_index,
                                    sampling_location, weights):
    """ops/__init__.py:6-19 — (mc_ms_feat [bs,num_feat,C], spatial_shape [K,L,2],
    scale_start_index [K,L], sampling_location [bs,A,P,K,2], weights [bs,A,P,K,L,G])
    → [bs, A, C] float32, differentiable wrt features, locations and weights."""
    return DeformableAggregationFunction.apply(feature_maps, spatial_shape, scale_start_index,
                                               sampling_location, weights)


class _FlattenMaps(torch.autograd.Function):
    """col_feats = flatten(maps) with the transposing kernel; the backward hands every level its
    slice of grad_col_feats back in NCHW (the reference gets this from autograd through its
    reshape / cat / permute chain)."""

    @staticmethod
    def forward(ctx, dtype, *maps):
        ctx.shapes = [tuple(m.shape) for m in maps]
        return cabi.flatten_maps(list(maps), out_dtype=dtype)

    @staticmethod
    def backward(ctx, grad_col):
        bs, K, C = ctx.shapes[0][:3]
        per_cam = sum(s[3] * s[4] for s in ctx.shapes)
        g = grad_col.reshape(bs, K, per_cam, C)
        out, o = [], 0
        for s in ctx.shapes:
            h, w = s[3], s[4]
            out.append(g[:, :, o:o + h * w].reshape(bs, K, h, w, C).permute(0, 1, 4, 2, 3)
                       .contiguous().float())
            o += h * w
        return (None,) + tuple(out)


_TABLES = {}


def _tables(sizes, num_cams, device):
    """The two small integer tables, built on the host once per (sizes, cameras, device) and cloned
    per call: no host-to-device copy on later calls, so a warmed-up call can be graph-captured."""
    key = (tuple(sizes), num_cams, str(device))
    if key not in _TABLES:
        shape = torch.tensor([list(sizes)] * num_cams, dtype=torch.int64)
        counts = (shape[..., 0] * shape[..., 1]).flatten()
        start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]]).reshape(num_cams, -1)
        _TABLES[key] = (shape.to(device), start.to(device))
    shape, start = _TABLES[key]
    return shape.clone(), start.clone()


def feature_maps_format(feature_maps, inverse=False, dtype=None, reference_start_index=False):
    """ops/__init__.py:22-92.  Forward direction: a list over levels of [bs, K, C, H_l, W_l] maps
    becomes [col_feats [bs, K*sum(H_l*W_l), C], spatial_shape [K,L,2] int64, scale_start_index
    [K,L] int64] with col_feats[b, start[k,l] + y*W_l + x, c] == maps[l][b,k,c,y,x] — done here by
    ONE transposing kernel per level writing straight into the final buffer (the reference does
    reshape + cat + permute + flatten).  `dtype=torch.bfloat16` emits a bf16 table (extension).
    A list of lists (camera groups with different resolutions) is formatted group by group and
    concatenated, as the reference does (:56-61).  DOCUMENTED DEVIATION: the reference concatenates
    the groups' own scale_start_index tables, each of which restarts at 0 — rows of every group after
    the first would then be looked up inside the first group's rows (the branch is never called in the
    reference repository).  Here the start table of a nested list counts rows from the beginning of
    col_feats, which is what the op needs; `reference_start_index=True` reproduces the reference's
    integers bit for bit (tests/golden/flatten_nested.npz).  col_feats and spatial_shape are identical
    either way.  inverse=True returns the nested list `[[level maps of camera group 0], ...]` of
    [bs, n_cam, C, H, W] views (:23-54); like the reference's it splits by the sizes in spatial_shape
    only and never reads scale_start_index, so it inverts both kinds of table."""
    if inverse:
        col, shape, start = feature_maps
        K, L = shape.shape[:2]
        sizes = shape.tolist()
        bs, _, C = col.shape
        groups, k, row0 = [], 0, 0
        while k < K:   # consecutive cameras with identical level sizes form one group
            k1 = k + 1
            while k1 < K and sizes[k1] == sizes[k]:
                k1 += 1
            per_cam = sum(h * w for h, w in sizes[k])
            block = col[:, row0:row0 + (k1 - k) * per_cam]
            row0 += (k1 - k) * per_cam
            block = block.reshape(bs, k1 - k, per_cam, C)
            levels, o = [], 0
            for h, w in sizes[k]:
                levels.append(block[:, :, o:o + h * w].reshape(bs, k1 - k, h, w, C)
                              .permute(0, 1, 4, 2, 3))
                o += h * w
            groups.append(levels)
            k = k1
        return groups

    if isinstance(feature_maps[0], (list, tuple)):
        parts = [feature_maps_format(x, dtype=dtype) for x in feature_maps]
        col = torch.cat([p[0] for p in parts], dim=1)
        shape = torch.cat([p[1] for p in parts], dim=0)
        if reference_start_index:     # ops/__init__.py:60: every group's table restarts at 0
            return [col, shape, torch.cat([p[2] for p in parts], dim=0)]
        counts = (shape[..., 0] * shape[..., 1]).flatten()
        start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]]).reshape(shape.shape[0], -1)
        return [col, shape, start]

    maps = [m.contiguous().float() for m in feature_maps]
    K = maps[0].shape[1]
    col = _FlattenMaps.apply(dtype or torch.float32, *maps)
    shape, start = _tables([tuple(m.shape[-2:]) for m in maps], K, col.device)
    return [col, shape, start]
