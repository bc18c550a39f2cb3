"""Python-level oracle for the pieces of SimPB's DFA module around the op.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain torch CPU ops, written for
clarity, each citing the reference lines it follows (paths relative to
/root/reference/projects/mmdet3d_plugin).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# anchor vector layout, core/box3d.py:1
X, Y, Z, W, L, H, SIN_YAW, COS_YAW, VX, VY, VZ = range(11)


# ------------------------------------------------------------------ feature layout
def flatten_feature_maps(maps):
    """ops/__init__.py:63-92.  `maps`: list over levels of [bs, K, C, H_l, W_l].
    Returns (col_feats [bs, K*sum(H_l*W_l), C], spatial_shape [K, L, 2] int64 (H, W),
    scale_start_index [K, L] int64) with
        col_feats[b, start[k,l] + y*W_l + x, c] == maps[l][b, k, c, y, x].
    """
    bs, K, C = maps[0].shape[:3]
    sizes = [(int(m.shape[-2]), int(m.shape[-1])) for m in maps]
    per_cam = sum(h * w for h, w in sizes)
    col = maps[0].new_empty(bs, K * per_cam, C)
    shape = torch.tensor([sizes] * K, dtype=torch.int64)
    start = torch.zeros(K, len(maps), dtype=torch.int64)
    row = 0
    for k in range(K):
        for l, m in enumerate(maps):
            h, w = sizes[l]
            start[k, l] = row
            col[:, row:row + h * w] = m[:, k].reshape(bs, C, h * w).transpose(1, 2)
            row += h * w
    return col, shape, start


def unflatten_feature_maps(col, shape, start):
    """Inverse of the above for the uniform-camera case (ops/__init__.py:23-54 returns a
    nested list `[[level maps of a camera group]]` of [bs, n_cam, C, H, W]); here: a flat
    list over levels of [bs, K, C, H_l, W_l]."""
    K, Lv = shape.shape[:2]
    bs, _, C = col.shape
    out = []
    for l in range(Lv):
        h, w = int(shape[0, l, 0]), int(shape[0, l, 1])
        lev = col.new_empty(bs, K, C, h, w)
        for k in range(K):
            s = int(start[k, l])
            lev[:, k] = col[:, s:s + h * w].transpose(1, 2).reshape(bs, C, h, w)
        out.append(lev)
    return out


# ------------------------------------------------------------------ geometry
def key_points(anchor, fix_scale, learnable_logits=None):
    """models/detection3d/blocks.py:181-222.  anchor [bs,A,11]; fix_scale [F,3];
    learnable_logits [bs,A,n*3] = learnable_fc(instance_feature) (pre-sigmoid) or None.
    Returns [bs, A, F+n, 3]."""
    bs, A = anchor.shape[:2]
    size = anchor[..., [W, L, H]].exp()[:, :, None]                      # :183
    pts = fix_scale[None, None] * size                                   # :184
    if learnable_logits is not None:
        off = learnable_logits.reshape(bs, A, -1, 3).sigmoid() - 0.5     # :186-191
        pts = torch.cat([pts, off * size], dim=2)                        # :192-194
    c, s = anchor[..., COS_YAW][..., None], anchor[..., SIN_YAW][..., None]
    x, y, z = pts.unbind(-1)
    rot = torch.stack([c * x - s * y, s * x + c * y, z], dim=-1)         # :196-206
    return rot + anchor[..., [X, Y, Z]][:, :, None]                      # :207


def project_points(pts, projection_mat, image_wh=None):
    """models/blocks.py:198-213.  pts [bs,A,P,3]; projection_mat [bs,K,4,4];
    image_wh [bs,K,2].  Returns [bs, K, A, P, 2] (x, y) normalised by the image size."""
    homo = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)       # :202-204
    cam = torch.einsum("bkij,bapj->bkapi", projection_mat, homo)         # :205-207
    uv = cam[..., :2] / cam[..., 2:3].clamp(min=1e-5)                    # :208-210
    if image_wh is not None:
        uv = uv / image_wh[:, :, None, None]                             # :211-212
    return uv


def op_valid_mask(loc):
    """…/ops/src/deformable_aggregation_cuda.cu:168-171 — exclusive (0,1) test on x and y."""
    return ~((loc[..., 0] <= 0) | (loc[..., 0] >= 1)) & ~((loc[..., 1] <= 0) | (loc[..., 1] >= 1))


# ------------------------------------------------------------------ grid_sample CPU path
def grid_sample_features(maps, uv):
    """models/blocks.py:215-246.  maps: list of [bs,K,C,H_l,W_l]; uv [bs,K,A,P,2] in [0,1]
    image-normalised.  Returns [bs, A, K, L, P, C]."""
    bs, K, A, P = uv.shape[:4]
    grid = (uv * 2 - 1).reshape(bs * K, A, P, 2)                         # :229-230
    per_level = [F.grid_sample(m.flatten(0, 1), grid, mode="bilinear", padding_mode="zeros",
                               align_corners=False) for m in maps]       # :233-238
    f = torch.stack(per_level, dim=1)                                    # [bs*K, L, C, A, P]
    f = f.reshape(bs, K, len(maps), -1, A, P)
    return f.permute(0, 4, 1, 2, 5, 3)                                   # :240-244


def fuse_views_levels(features, weights, num_groups):
    """models/blocks.py:248-261 followed by the point sum of :156.
    features [bs,A,K,L,P,C], weights [bs,A,K,L,P,G] → [bs, A, C]."""
    bs, A, K, Lv, P, C = features.shape
    f = features.reshape(bs, A, K, Lv, P, num_groups, C // num_groups)
    f = (weights[..., None] * f).sum(dim=2).sum(dim=2)                   # cams, then levels
    return f.reshape(bs, A, P, C).sum(dim=2)


def aggregate_grid_sample(maps, uv, weights_kl, num_groups, apply_op_mask):
    """The reference CPU path end to end for given projected points.
    uv [bs,K,A,P,2]; weights_kl [bs,A,K,L,P,G].  With apply_op_mask the op's (0,1) mask is
    multiplied in — the reference CUDA op and grid_sample differ in the half-pixel border
    band."""
    f = grid_sample_features(maps, uv)
    if apply_op_mask:
        m = op_valid_mask(uv).permute(0, 2, 1, 3)                        # [bs,A,K,P]
        f = f * m[:, :, :, None, :, None].to(f.dtype)
    return fuse_views_levels(f, weights_kl, num_groups)


# ------------------------------------------------------------------ the module
def _linear_relu_ln(dims, in_loops, out_loops, input_dims):
    layers = []                                                           # models/blocks.py:32-42
    for _ in range(out_loops):
        for _ in range(in_loops):
            layers += [nn.Linear(input_dims, dims), nn.ReLU(inplace=True)]
            input_dims = dims
        layers.append(nn.LayerNorm(dims))
    return layers


class DFAModuleRef(nn.Module):
    """CPU restatement of DeformableFeatureAggregation (models/blocks.py:45-261) with the
    released-config sub-modules; same state-dict keys as the reference.  `op` selects how the
    aggregation itself is evaluated: "grid_sample" (reference fallback, unmasked),
    "grid_sample_masked" (fallback × op mask) or a callable with the op signature."""

    def __init__(self, embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.0,
                 fix_scale=((0.0, 0.0, 0.0),), num_learnable_pts=0, use_camera_embed=False,
                 residual_mode="add", op="grid_sample"):
        super().__init__()
        self.embed_dims, self.num_groups = embed_dims, num_groups
        self.num_levels, self.num_cams = num_levels, num_cams
        self.attn_drop, self.residual_mode, self.op = attn_drop, residual_mode, op
        self.kps_generator = nn.Module()
        self.kps_generator.fix_scale = nn.Parameter(torch.tensor(fix_scale, dtype=torch.float32),
                                                    requires_grad=False)
        self.num_learnable_pts = num_learnable_pts
        if num_learnable_pts > 0:
            self.kps_generator.learnable_fc = nn.Linear(embed_dims, num_learnable_pts * 3)
        self.num_pts = len(fix_scale) + num_learnable_pts
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        if use_camera_embed:
            self.camera_encoder = nn.Sequential(*_linear_relu_ln(embed_dims, 1, 2, 12))
            self.weights_fc = nn.Linear(embed_dims, num_groups * num_levels * self.num_pts)
        else:
            self.camera_encoder = None
            self.weights_fc = nn.Linear(embed_dims,
                                        num_groups * num_cams * num_levels * self.num_pts)

    def attention_weights(self, instance_feature, anchor_embed, projection_mat, drop_mask=None):
        """models/blocks.py:164-196 → [bs, A, K, L, P, G]."""
        bs, A = instance_feature.shape[:2]
        f = instance_feature + anchor_embed
        if self.camera_encoder is not None:
            cam = self.camera_encoder(projection_mat[:, :, :3].reshape(bs, self.num_cams, -1))
            f = f[:, :, None] + cam[:, None]
        w = self.weights_fc(f).reshape(bs, A, -1, self.num_groups).softmax(dim=-2)
        w = w.reshape(bs, A, self.num_cams, self.num_levels, self.num_pts, self.num_groups)
        if drop_mask is not None:                                         # :188-195
            w = (drop_mask.to(w.dtype) * w) / (1 - self.attn_drop)
        return w

    def forward(self, instance_feature, anchor, anchor_embed, maps, projection_mat, image_wh,
                drop_mask=None):
        logits = (self.kps_generator.learnable_fc(instance_feature)
                  if self.num_learnable_pts > 0 else None)
        pts = key_points(anchor, self.kps_generator.fix_scale, logits)
        w = self.attention_weights(instance_feature, anchor_embed, projection_mat, drop_mask)
        uv = project_points(pts, projection_mat, image_wh)
        if callable(self.op):                                             # models/blocks.py:123-147
            col, shape, start = flatten_feature_maps(maps)
            feats = self.op(col, shape, start, uv.permute(0, 2, 3, 1, 4).contiguous(),
                            w.permute(0, 1, 4, 2, 3, 5).contiguous())
        else:
            feats = aggregate_grid_sample(maps, uv, w, self.num_groups,
                                          apply_op_mask=(self.op == "grid_sample_masked"))
        out = self.output_proj(feats)
        if self.residual_mode == "add":
            return out + instance_feature
        if self.residual_mode == "cat":
            return torch.cat([out, instance_feature], dim=-1)
        return out
