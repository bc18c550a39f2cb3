"""Host-side mirror of the reference's DFA module for the fused CUDA path.

`DeformableFeatureAggregation` and `SparseBox3DKeyPointsGenerator` keep the constructor
arguments, parameter names (state-dict keys) and forward signatures of
/root/reference/projects/mmdet3d_plugin/models/blocks.py:45-162 and
/root/reference/projects/mmdet3d_plugin/models/detection3d/blocks.py:157-222, so released SimPB
checkpoints load and the modules can be built from the reference's config dict
(projects/configs/simpb_nus_r50_img_704x256.py:216-239).

What runs where: the Linear layers, the two LayerNorms and the residual stay in PyTorch
(cuBLAS / ATen); everything between them is libdfa_b200 —
  * inference (no gradient needed): ONE launch, dfa_forward_fused — key points, camera projection,
    softmax of the attention logits and the aggregation; neither the sampling locations nor the
    9 MB weights tensor are materialised;
  * training: three launches, each with its backward —
      dfa_keypoints_project   anchors (+ learnable offsets) -> sampling locations [bs,A,P,K,2]
      dfa_softmax_weights     weights_fc logits -> softmax over cams x levels x points, attn-drop,
                              laid out [bs,A,P,K,L,G]
      dfa_forward             the aggregation itself
instead of the ~40 small ATen kernels and the 9 MB permute copy of the reference's forward.  With
camera embedding the weights_fc GEMM runs on bs*(A+K) rows instead of bs*A*K (weights_fc is linear;
the broadcast add happens inside the kernels).  There is no grid_sample / CPU path:
`use_deformable_func=False` raises.
"""
import os

import torch
import torch.nn as nn
from torch.autograd.function import Function, once_differentiable

from . import cabi
from .ops import deformable_aggregation_function as DAF

__all__ = ["DeformableFeatureAggregation", "SparseBox3DKeyPointsGenerator", "build_kps_generator"]

# anchor vector layout (core/box3d.py:1)
X, Y, Z, W, L, H, SIN_YAW, COS_YAW, VX, VY, VZ = range(11)


class _KeyPointsProject(Function):
    """sampling_location = project(key_points(anchor, offsets)) — models/detection3d/blocks.py:181-207
    followed by models/blocks.py:198-213 and the permute of :124-132, in one kernel."""

    @staticmethod
    def forward(ctx, anchor, fix_scale, logits, projection_mat, image_wh):
        anchor = anchor.contiguous().float()
        fix_scale = fix_scale.contiguous().float()
        logits = None if logits is None else logits.contiguous().float()
        projection_mat = projection_mat.contiguous().float()
        image_wh = None if image_wh is None else image_wh.contiguous().float()
        ctx.save_for_backward(anchor, fix_scale, logits, projection_mat, image_wh)
        return cabi.keypoints_project(anchor, fix_scale, logits, projection_mat, image_wh)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc):
        anchor, fix_scale, logits, projection_mat, image_wh = ctx.saved_tensors
        g_anchor, g_logits = cabi.keypoints_project_backward(
            anchor, fix_scale, logits, projection_mat, image_wh, grad_loc.contiguous().float())
        return g_anchor, None, g_logits, None, None


class _SoftmaxWeights(Function):
    """models/blocks.py:175-195 + the permute of :133-144."""

    @staticmethod
    def forward(ctx, logits, dims, keep, scale):
        logits = logits.contiguous().float()
        ctx.dims, ctx.scale = dims, scale
        ctx.save_for_backward(logits, keep)
        return cabi.softmax_weights(logits, dims, keep, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_w):
        logits, keep = ctx.saved_tensors
        g = cabi.softmax_weights_backward(logits, ctx.dims, keep, ctx.scale, grad_w.contiguous().float())
        return g, None, None, None


class _SoftmaxWeightsSplit(Function):
    """The same for the camera-embedding branch (:166-174), using that weights_fc is linear:
    weights_fc(feature[b,a] + cam[b,k]) = weights_fc(feature)[b,a] + (cam @ W^T)[b,k].  The GEMM then
    runs on bs*(A+K) rows instead of bs*A*K, and neither the [bs,A,K,C] sum nor the 9 MB logits
    tensor is ever materialised: the kernel adds the two parts while it reads them."""

    @staticmethod
    def forward(ctx, logits_anchor, logits_cam, dims, keep, scale):
        logits_anchor = logits_anchor.contiguous().float()
        logits_cam = logits_cam.contiguous().float()
        ctx.dims, ctx.scale = dims, scale
        ctx.save_for_backward(logits_anchor, logits_cam, keep)
        return cabi.softmax_weights_split(logits_anchor, logits_cam, dims, keep, scale)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_w):
        la, lk, keep = ctx.saved_tensors
        ga, gk = cabi.softmax_weights_split_backward(la, lk, ctx.dims, keep, ctx.scale,
                                                     grad_w.contiguous().float())
        return ga, gk, None, None, None


class SparseBox3DKeyPointsGenerator(nn.Module):
    """models/detection3d/blocks.py:157-222.  Inside DFA the fused kernel is used; calling the
    module directly returns the 3-D key points [bs, A, P, 3] like the reference (plain torch —
    this entry point is not on the hot path)."""

    def __init__(self, embed_dims=256, num_learnable_pts=0, fix_scale=None):
        super().__init__()
        self.embed_dims = embed_dims
        self.num_learnable_pts = num_learnable_pts
        if fix_scale is None:
            fix_scale = ((0.0, 0.0, 0.0),)
        self.fix_scale = nn.Parameter(torch.tensor(fix_scale, dtype=torch.float32), requires_grad=False)
        self.num_pts = len(self.fix_scale) + num_learnable_pts
        if num_learnable_pts > 0:
            self.learnable_fc = nn.Linear(self.embed_dims, num_learnable_pts * 3)

    def init_weight(self):
        if self.num_learnable_pts > 0:
            nn.init.xavier_uniform_(self.learnable_fc.weight)
            nn.init.constant_(self.learnable_fc.bias, 0.0)

    def offset_logits(self, instance_feature):
        """Pre-sigmoid learnable offsets [bs, A, num_learnable_pts*3], or None."""
        if self.num_learnable_pts > 0 and instance_feature is not None:
            return self.learnable_fc(instance_feature)
        return None

    def forward(self, anchor, instance_feature=None, T_cur2temp_list=None, cur_timestamp=None,
                temp_timestamps=None):
        bs, A = anchor.shape[:2]
        size = anchor[..., None, [W, L, H]].exp()
        pts = self.fix_scale * size
        logits = self.offset_logits(instance_feature)
        if logits is not None:
            pts = torch.cat([pts, (logits.reshape(bs, A, -1, 3).sigmoid() - 0.5) * size], dim=-2)
        cos, sin = anchor[..., None, COS_YAW], anchor[..., None, SIN_YAW]
        pts = torch.stack([cos * pts[..., 0] - sin * pts[..., 1],
                           sin * pts[..., 0] + cos * pts[..., 1], pts[..., 2]], dim=-1)
        pts = pts + anchor[..., None, X:Z + 1]
        if (cur_timestamp is None or temp_timestamps is None or T_cur2temp_list is None
                or len(temp_timestamps) == 0):
            return pts
        # temporal key points (:209-222): move by velocity x dt, then into the past frame
        past = []
        vel = anchor[..., VX:]
        for T, t in zip(T_cur2temp_list, temp_timestamps):
            dt = (cur_timestamp - t).to(vel.dtype)
            moved = pts - (vel * dt[:, None, None])[:, :, None]
            T = T.to(pts.dtype)
            past.append(moved @ T[:, None, :3, :3].transpose(-1, -2) + T[:, None, None, :3, 3])
        return pts, past

    @staticmethod
    def anchor_projection(anchor, T_src2dst_list, src_timestamp=None, dst_timestamps=None,
                          time_intervals=None):
        """:224-258 — anchors carried into other frames (used by the instance bank, not by DFA)."""
        out = []
        for i, T in enumerate(T_src2dst_list):
            T = T.to(anchor.dtype)[:, None]
            vel = anchor[..., VX:]
            nv = vel.shape[-1]
            center = anchor[..., X:Z + 1]
            if time_intervals is not None:
                dt = time_intervals[i]
            elif src_timestamp is not None and dst_timestamps is not None:
                dt = (src_timestamp - dst_timestamps[i]).to(vel.dtype)
            else:
                dt = None
            if dt is not None:
                center = center - (vel.transpose(0, -1) * dt).transpose(0, -1)
            center = (T[..., :3, :3] @ center[..., None]).squeeze(-1) + T[..., :3, 3]
            yaw = (T[..., :2, :2] @ anchor[..., SIN_YAW:COS_YAW + 1].flip(-1)[..., None]).squeeze(-1)
            vel = (T[..., :nv, :nv] @ vel[..., None]).squeeze(-1)
            out.append(torch.cat([center, anchor[..., W:H + 1], yaw, vel], dim=-1))
        return out

    @staticmethod
    def distance(anchor):
        return torch.norm(anchor[..., :2], p=2, dim=-1)


_KPS_TYPES = {"SparseBox3DKeyPointsGenerator": SparseBox3DKeyPointsGenerator}


def build_kps_generator(cfg):
    """The one `build_from_cfg` call the module makes (models/blocks.py:80-81), without mmcv."""
    if isinstance(cfg, nn.Module):
        return cfg
    cfg = dict(cfg)
    typ = cfg.pop("type")
    if not isinstance(typ, str):
        return typ(**cfg)
    if typ not in _KPS_TYPES:
        raise KeyError("unknown key-point generator %r (known: %s)" % (typ, sorted(_KPS_TYPES)))
    return _KPS_TYPES[typ](**cfg)


def _linear_relu_ln(embed_dims, in_loops, out_loops, input_dims):
    layers = []                                    # models/blocks.py:32-42
    for _ in range(out_loops):
        for _ in range(in_loops):
            layers += [nn.Linear(input_dims, embed_dims), nn.ReLU(inplace=True)]
            input_dims = embed_dims
        layers.append(nn.LayerNorm(embed_dims))
    return layers


class DeformableFeatureAggregation(nn.Module):
    """models/blocks.py:45-162 on the fused CUDA path.  `feature_maps` is the triple produced by
    `simpb_b200.ops.feature_maps_format` (col_feats fp32 or bf16, spatial_shape, scale_start_index);
    `metas` needs "projection_mat" [bs,K,4,4] and optionally "image_wh" [bs,K,2]."""

    def __init__(self, embed_dims=256, num_groups=8, num_levels=4, num_cams=6, proj_drop=0.0,
                 attn_drop=0.0, kps_generator=None, temporal_fusion_module=None,
                 use_temporal_anchor_embed=True, use_deformable_func=True, use_camera_embed=False,
                 residual_mode="add"):
        super().__init__()
        if embed_dims % num_groups != 0:
            raise ValueError("embed_dims must be divisible by num_groups, but got %d and %d"
                             % (embed_dims, num_groups))
        if not use_deformable_func:
            raise ValueError("simpb_b200 has no grid_sample path: use_deformable_func must be True")
        self.group_dims = embed_dims // num_groups
        self.embed_dims, self.num_levels = embed_dims, num_levels
        self.num_groups, self.num_cams = num_groups, num_cams
        self.use_temporal_anchor_embed = use_temporal_anchor_embed   # stored only, as upstream (:73)
        self.use_deformable_func = True
        self.attn_drop, self.residual_mode = attn_drop, residual_mode
        self.proj_drop = nn.Dropout(proj_drop)
        kps_generator = dict(kps_generator or dict(type="SparseBox3DKeyPointsGenerator"))
        kps_generator["embed_dims"] = embed_dims
        self.kps_generator = build_kps_generator(kps_generator)
        self.num_pts = self.kps_generator.num_pts
        # The attention-weight kernels (csrc/dfa_frontend.cu) stage one anchor's K*L*P*G logits in shared
        # memory and reduce groups with 256-thread blocks; say so here instead of failing inside forward()
        # (the reference accepts any embed_dims % num_groups == 0; there is no eager fallback by design).
        n_logits = num_cams * num_levels * self.num_pts * num_groups
        if 256 % num_groups != 0 or num_cams * num_levels * self.num_pts >= 65536 or 4 * n_logits > 190 * 1024:
            raise ValueError("DeformableFeatureAggregation: num_groups must divide 256 and one anchor's "
                             "K*L*P*G = %d attention logits must fit 190 KB of shared memory" % n_logits)
        # upstream builds temporal_fusion_module and never calls it in forward (:83-90); only an
        # already-built module is accepted here so that checkpoints with such keys still load
        self.temp_module = temporal_fusion_module if isinstance(temporal_fusion_module, nn.Module) else None
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        if use_camera_embed:
            self.camera_encoder = nn.Sequential(*_linear_relu_ln(embed_dims, 1, 2, 12))
            self.weights_fc = nn.Linear(embed_dims, num_groups * num_levels * self.num_pts)
        else:
            self.camera_encoder = None
            self.weights_fc = nn.Linear(embed_dims, num_groups * num_cams * num_levels * self.num_pts)

    def init_weight(self):
        nn.init.constant_(self.weights_fc.weight, 0.0)   # :106-108
        nn.init.constant_(self.weights_fc.bias, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.0)

    def weight_logits(self, instance_feature, anchor_embed, metas):
        """:164-174 — the input of the softmax, memory order (bs, A, K, L, P, G)."""
        bs = instance_feature.shape[0]
        feature = instance_feature + anchor_embed
        if self.camera_encoder is not None:
            cam = self.camera_encoder(metas["projection_mat"][:, :, :3].reshape(bs, self.num_cams, -1))
            feature = feature[:, :, None] + cam[:, None]
        return self.weights_fc(feature)

    def sampling_and_weights(self, instance_feature, anchor, anchor_embed, metas, keep=None):
        """The two op inputs: sampling_location [bs,A,P,K,2] and weights [bs,A,P,K,L,G].
        `keep` (bool/uint8 [bs,A,K,P]) overrides the random attn-drop mask (tests)."""
        bs, A = instance_feature.shape[:2]
        gen = self.kps_generator
        loc = _KeyPointsProject.apply(anchor, gen.fix_scale, gen.offset_logits(instance_feature),
                                      metas["projection_mat"], metas.get("image_wh"))
        scale = 1.0
        if keep is None and self.training and self.attn_drop > 0:
            keep = torch.rand(bs, A, self.num_cams, self.num_pts, device=anchor.device) > self.attn_drop
        if keep is not None:
            keep = keep.to(torch.uint8).contiguous()
            scale = 1.0 / (1.0 - self.attn_drop)
        dims = (bs, A, self.num_cams, self.num_levels, self.num_pts, self.num_groups)
        if self.camera_encoder is not None:
            cam = self.camera_encoder(metas["projection_mat"][:, :, :3].reshape(bs, self.num_cams, -1))
            w = _SoftmaxWeightsSplit.apply(self.weights_fc(instance_feature + anchor_embed),
                                           nn.functional.linear(cam, self.weights_fc.weight),
                                           dims, keep, scale)
        else:
            w = _SoftmaxWeights.apply(self.weight_logits(instance_feature, anchor_embed, metas), dims,
                                      keep, scale)
        return loc, w

    def fused_features(self, instance_feature, anchor, anchor_embed, feature_maps, metas):
        """Inference path: key points, projection, softmax and aggregation in ONE launch
        (dfa_forward_fused); None when the shape is outside that kernel's fast path."""
        bs, A = instance_feature.shape[:2]
        col, shape, start = feature_maps
        if col.dtype != torch.bfloat16:
            col = col.float()
        gen = self.kps_generator
        logits_a = self.weights_fc(instance_feature + anchor_embed)
        logits_k = None
        if self.camera_encoder is not None:
            cam = self.camera_encoder(metas["projection_mat"][:, :, :3].reshape(bs, self.num_cams, -1))
            logits_k = nn.functional.linear(cam, self.weights_fc.weight).contiguous().float()
        off = gen.offset_logits(instance_feature)
        wh = metas.get("image_wh")
        i32 = lambda t: t if t.dtype == torch.int32 and t.is_contiguous() else t.contiguous().int()  # noqa: E731
        f32 = lambda t: None if t is None else t.contiguous().float()  # noqa: E731
        return cabi.forward_fused(col.contiguous(), i32(shape), i32(start), f32(anchor), f32(gen.fix_scale),
                                  f32(off), f32(metas["projection_mat"]), f32(wh), f32(logits_a), logits_k,
                                  (bs, A, self.num_cams, self.num_levels, self.num_pts, self.num_groups))

    def forward(self, instance_feature, anchor, anchor_embed, feature_maps, metas, **kwargs):
        features = None
        keep = kwargs.get("attn_keep_mask")
        no_grad = not torch.is_grad_enabled() or not (
            instance_feature.requires_grad or anchor.requires_grad or anchor_embed.requires_grad
            or feature_maps[0].requires_grad or any(p.requires_grad for p in self.parameters()))
        if keep is None and not (self.training and self.attn_drop > 0) and no_grad:
            features = self.fused_features(instance_feature, anchor, anchor_embed, feature_maps, metas)
        if features is None:
            loc, w = self.sampling_and_weights(instance_feature, anchor, anchor_embed, metas, keep=keep)
            features = DAF(*feature_maps, loc, w)
        output = self.proj_drop(self.output_proj(features))
        if self.residual_mode == "add":
            output = output + instance_feature
        elif self.residual_mode == "cat":
            output = torch.cat([output, instance_feature], dim=-1)
        return output


def _register_with_mmcv():
    """Optional: register under the reference's names in mmcv's registries (mmcv 1.x)."""
    try:
        from mmcv.cnn.bricks.registry import ATTENTION, PLUGIN_LAYERS
    except Exception:
        return False
    ATTENTION.register_module(name="DeformableFeatureAggregation", force=True,
                              module=DeformableFeatureAggregation)
    PLUGIN_LAYERS.register_module(name="SparseBox3DKeyPointsGenerator", force=True,
                                  module=SparseBox3DKeyPointsGenerator)
    return True


if os.environ.get("SIMPB_B200_REGISTER") == "1":
    _register_with_mmcv()
