"""Seeded synthetic inputs of SimPB's shapes (S0 "uniform", S1 "rig").

Everything is generated on the CPU from a `torch.Generator` so the same tensors can be fed
to the CUDA path, the oracle and the reference.  No dataset or checkpoint is involved.
"""
import math

import torch

# (H, W) of the 4 FPN levels, strides 4/8/16/32
# (/root/reference/projects/configs/simpb_nus_r50_img_704x256.py:27,56)
R50_LEVELS = ((64, 176), (32, 88), (16, 44), (8, 22))       # input 704x256
R101_LEVELS = ((128, 352), (64, 176), (32, 88), (16, 44))   # input 1408x512
FIX_SCALE = ((0.0, 0.0, 0.0), (0.45, 0.0, 0.0), (-0.45, 0.0, 0.0), (0.0, 0.45, 0.0),
             (0.0, -0.45, 0.0), (0.0, 0.0, 0.45), (0.0, 0.0, -0.45))   # config :229-237


def level_tables(levels=R50_LEVELS, num_cams=6):
    """spatial_shape [K,L,2] (H,W) and scale_start_index [K,L], int64, exactly what
    feature_maps_format produces (/root/reference/projects/mmdet3d_plugin/ops/__init__.py:74-84)."""
    shape = torch.tensor([list(levels)] * num_cams, dtype=torch.int64)
    cnt = (shape[..., 0] * shape[..., 1]).flatten()
    start = torch.cat([cnt.new_zeros(1), cnt.cumsum(0)[:-1]]).reshape(num_cams, len(levels))
    return shape, start, int(cnt.sum())


def softmax_weights(gen, bs, A, P, K, L, G):
    """softmax over (K,L,P) of N(0,1) logits per (b,a,g), laid out [bs,A,P,K,L,G]
    (models/blocks.py:175-187 then the permute of :133-144)."""
    logits = torch.randn(bs, A, K * L * P, G, generator=gen)
    w = logits.softmax(dim=2).reshape(bs, A, K, L, P, G)
    return w.permute(0, 1, 4, 2, 3, 5).contiguous()


def op_inputs_uniform(bs=1, A=900, P=13, K=6, levels=R50_LEVELS, C=256, G=8, seed=0,
                      lo=-0.1, hi=1.1, feat=True):
    """S0: loc ~ U(lo,hi) (≈69 % valid for (-0.1,1.1)), feat ~ N(0,1), softmaxed weights."""
    gen = torch.Generator().manual_seed(seed)
    shape, start, num_feat = level_tables(levels, K)
    d = dict(spatial_shape=shape, scale_start_index=start, num_feat=num_feat)
    d["mc_ms_feat"] = torch.randn(bs, num_feat, C, generator=gen) if feat else None
    d["sampling_location"] = torch.rand(bs, A, P, K, 2, generator=gen) * (hi - lo) + lo
    d["weights"] = softmax_weights(gen, bs, A, P, K, len(levels), G)
    d["grad_output"] = torch.randn(bs, A, C, generator=gen)
    return d


def camera_rig(bs=1, scale=0.44, crop_h=140.0, image_wh=(704.0, 256.0)):
    """nuScenes-like 6-camera rig: returns projection_mat [bs,6,4,4]
    (lidar → pixel) and image_wh [bs,6,2]."""
    yaws = (0.0, -55.0, 55.0, 180.0, 110.0, -110.0)
    focal = (1266.4, 1266.4, 1266.4, 809.2, 1256.7, 1256.7)
    mats = []
    for yaw, f in zip(yaws, focal):
        a = math.radians(yaw)
        fwd = torch.tensor([-math.sin(a), math.cos(a), 0.0], dtype=torch.float64)
        right = torch.tensor([math.cos(a), math.sin(a), 0.0], dtype=torch.float64)
        down = torch.tensor([0.0, 0.0, -1.0], dtype=torch.float64)
        R = torch.stack([right, down, fwd])
        centre = 0.5 * fwd + torch.tensor([0.0, 0.0, -0.3], dtype=torch.float64)
        E = torch.eye(4, dtype=torch.float64)
        E[:3, :3] = R
        E[:3, 3] = -R @ centre
        Kmat = torch.eye(4, dtype=torch.float64)
        Kmat[0, 0] = Kmat[1, 1] = f * scale
        Kmat[0, 2] = 816.3 * scale
        Kmat[1, 2] = 491.5 * scale - crop_h
        mats.append(Kmat @ E)
    proj = torch.stack(mats).float()[None].repeat(bs, 1, 1, 1)
    wh = torch.tensor(image_wh, dtype=torch.float32)[None, None].repeat(bs, 6, 1)
    return proj, wh


def rig_anchors(gen, bs, A):
    """anchors [bs,A,11] in the reference encoding
    (/root/reference/tools/anchor_generator.py:23-27, core/box3d.py:1)."""
    r = 55.0 * torch.rand(bs, A, generator=gen).sqrt()
    th = 2 * math.pi * torch.rand(bs, A, generator=gen)
    z = -1.0 + 0.5 * torch.randn(bs, A, generator=gen)
    logsize = torch.tensor([0.65, 1.5, 0.55]) + 0.3 * torch.randn(bs, A, 3, generator=gen)
    yaw = 2 * math.pi * torch.rand(bs, A, generator=gen)
    vel = torch.randn(bs, A, 3, generator=gen)
    return torch.cat([(r * th.cos())[..., None], (r * th.sin())[..., None], z[..., None], logsize,
                      yaw.sin()[..., None], yaw.cos()[..., None], vel], dim=-1)


def module_inputs_rig(bs=1, A=900, levels=R50_LEVELS, C=256, seed=0, feat=True,
                      as_list=True):
    """S1: realistic module-level inputs.  `feature_maps` is the un-flattened list of
    [bs,6,C,H_l,W_l] maps (what the reference detector hands to DFA when
    use_deformable_func=False, models/simpb.py:79-88)."""
    gen = torch.Generator().manual_seed(seed)
    big = levels[0][0] > 64
    proj, wh = camera_rig(bs, scale=0.88 if big else 0.44, crop_h=280.0 if big else 140.0,
                          image_wh=(1408.0, 512.0) if big else (704.0, 256.0))
    d = dict(projection_mat=proj, image_wh=wh)
    d["anchor"] = rig_anchors(gen, bs, A)
    d["instance_feature"] = torch.randn(bs, A, C, generator=gen)
    d["anchor_embed"] = torch.randn(bs, A, C, generator=gen)
    d["feature_maps"] = ([torch.randn(bs, 6, C, h, w, generator=gen) for h, w in levels]
                         if feat else None)
    return d


def rig_op_inputs(bs=1, A=900, P=13, levels=R50_LEVELS, C=256, G=8, seed=0, feat=True):
    """S1 at op level: sampling locations from the rig geometry (fixed + U(-0.5,0.5)·size
    learnable key points projected through the 6 cameras), softmaxed N(0,1) weights."""
    gen = torch.Generator().manual_seed(seed)
    K = 6
    big = levels[0][0] > 64
    proj, wh = camera_rig(bs, scale=0.88 if big else 0.44, crop_h=280.0 if big else 140.0,
                          image_wh=(1408.0, 512.0) if big else (704.0, 256.0))
    anchor = rig_anchors(gen, bs, A)
    size = anchor[..., 3:6].exp()[:, :, None]
    fix = torch.tensor(FIX_SCALE)
    n_learn = P - fix.shape[0]
    pts = fix[None, None] * size
    if n_learn > 0:
        pts = torch.cat([pts, (torch.rand(bs, A, n_learn, 3, generator=gen) - 0.5) * size], 2)
    else:
        pts = pts[:, :, :P]
    c, s = anchor[..., 7, None], anchor[..., 6, None]
    x, y, z = pts.unbind(-1)
    pts = torch.stack([c * x - s * y, s * x + c * y, z], -1) + anchor[..., None, :3]
    key_points = pts
    homo = torch.cat([pts, torch.ones_like(pts[..., :1])], -1)
    cam = torch.einsum("bkij,bapj->bapki", proj, homo)
    loc = cam[..., :2] / cam[..., 2:3].clamp(min=1e-5) / wh[:, None, None]
    shape, start, num_feat = level_tables(levels, K)
    d = dict(spatial_shape=shape, scale_start_index=start, num_feat=num_feat)
    # the 3-D key points and camera matrices the locations come from (the reference's grid_sample path
    # projects them itself, models/blocks.py:215-230)
    d["key_points"], d["projection_mat"], d["image_wh"] = key_points.contiguous(), proj, wh
    d["mc_ms_feat"] = torch.randn(bs, num_feat, C, generator=gen) if feat else None
    d["sampling_location"] = loc.contiguous()
    d["weights"] = softmax_weights(gen, bs, A, P, K, len(levels), G)
    d["grad_output"] = torch.randn(bs, A, C, generator=gen)
    return d
