"""Generates tests/golden/*.npz by running the UNMODIFIED reference Python code.

Run in the build container only (it needs /root/reference, which does not exist on the GPU
box):  python tests/golden/make_golden.py

The reference plugin cannot be imported as a package without mmcv/mmdet (its __init__ star-
imports them), so the two files on the hot path are imported with (a) empty package objects
pre-seeded in sys.modules and (b) a minimal stand-in for the handful of mmcv symbols they use
(Linear = nn.Linear, registries, init helpers).  No reference arithmetic is replaced: every
number stored below is produced by
  projects/mmdet3d_plugin/ops/__init__.py            (feature_maps_format)
  projects/mmdet3d_plugin/models/blocks.py           (DeformableFeatureAggregation)
  projects/mmdet3d_plugin/models/detection3d/blocks.py (SparseBox3DKeyPointsGenerator)
and torch's grid_sample.  Fixtures are kept small (a few hundred kB in total).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    """The stub recipe lives in oracle/ref_import.py (bench.py's reference arm uses it as well)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import ref_import
    return ref_import.import_reference(REF)


FIX7 = [[0, 0, 0], [0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0], [0, -0.45, 0], [0, 0, 0.45],
        [0, 0, -0.45]]


def make_dfa(blocks, embed, groups, levels, cams, n_learn, fix_scale, camera_embed, residual):
    return blocks.DeformableFeatureAggregation(
        embed_dims=embed, num_groups=groups, num_levels=levels, num_cams=cams, attn_drop=0.15,
        use_deformable_func=False, use_camera_embed=camera_embed, residual_mode=residual,
        kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=n_learn,
                           fix_scale=fix_scale))


def op_case(blocks, name, seed, bs, A, P, K, sizes, C, G, lo, hi, masked, dtype):
    """Reference fallback (`feature_sampling` + `multi_view_level_fusion` + point sum) on op-level
    inputs.  The sampling locations are injected through an identity projection:
    key_points = (x, y, 1), projection_mat = I, image_wh = None  ⇒  points_2d == (x, y)."""
    g = torch.Generator().manual_seed(seed)
    maps = [torch.randn(bs, K, C, h, w, generator=g).to(dtype).requires_grad_() for h, w in sizes]
    loc = (torch.rand(bs, A, P, K, 2, generator=g) * (hi - lo) + lo)
    # store/compute from the float32 value the op will see
    loc = loc.float().to(dtype).requires_grad_()
    logits = torch.randn(bs, A, K * len(sizes) * P, G, generator=g)
    w_klp = logits.softmax(2).reshape(bs, A, K, len(sizes), P, G).float().to(dtype).requires_grad_()
    grad_out = torch.randn(bs, A, C, generator=g).float().to(dtype)
    dfa = make_dfa(blocks, C, G, len(sizes), K, 0, [[0, 0, 0]] * P, False, "add")
    outs, grads_loc = 0, None
    # feature_sampling projects ONE set of key points into all cameras; the op takes a separate
    # location per camera, so evaluate camera by camera (other cameras' weights zeroed).
    total = 0
    for k in range(K):
        kp = torch.cat([loc[:, :, :, k], torch.ones_like(loc[:, :, :, k, :1])], -1)
        proj = torch.eye(4, dtype=dtype)[None, None].repeat(bs, K, 1, 1)
        f = blocks.DeformableFeatureAggregation.feature_sampling(maps, kp, proj, None)
        if masked:
            x, y = loc[:, :, :, k, 0], loc[:, :, :, k, 1]
            m = ((x > 0) & (x < 1) & (y > 0) & (y < 1)).to(dtype)          # op mask, .cu:168-171
            f = f * m[:, :, None, None, :, None]
        sel = torch.zeros(K, dtype=dtype)
        sel[k] = 1
        wk = w_klp * sel[None, None, :, None, None, None]
        total = total + dfa.multi_view_level_fusion(f, wk).sum(dim=2)
    total.backward(grad_out)
    col, shape, start = None, None, None
    d = dict(loc=loc.detach().float().numpy(),
             weights=w_klp.detach().permute(0, 1, 4, 2, 3, 5).contiguous().float().numpy(),
             grad_out=grad_out.float().numpy(), out=total.detach().numpy(),
             grad_loc=loc.grad.numpy(),
             grad_weights=w_klp.grad.permute(0, 1, 4, 2, 3, 5).contiguous().numpy(),
             sizes=np.array(sizes, np.int64), G=np.int64(G), masked=np.int64(masked))
    for l, m in enumerate(maps):
        d["map%d" % l] = m.detach().float().numpy()
        d["grad_map%d" % l] = m.grad.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "out", total.shape, "abs max", float(total.abs().max()))


def flatten_case(ops, name, seed, bs, K, C, sizes):
    g = torch.Generator().manual_seed(seed)
    maps = [torch.randn(bs, K, C, h, w, generator=g) for h, w in sizes]
    col, shape, start = ops.feature_maps_format(maps)
    back = ops.feature_maps_format([col, shape, start], inverse=True)
    d = dict(col=col.numpy(), shape=shape.numpy(), start=start.numpy(),
             sizes=np.array(sizes, np.int64),
             inverse_n_groups=np.int64(len(back)), inverse_n_levels=np.int64(len(back[0])))
    for l, m in enumerate(maps):
        d["map%d" % l] = m.numpy()
        d["inv%d" % l] = back[0][l].contiguous().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, col.shape, shape.tolist()[0], start.tolist()[0])


def flatten_nested_case(ops, name, seed, bs, C, groups):
    """A list of camera GROUPS with different resolutions (ops/__init__.py:56-61): the reference formats
    every group on its own and concatenates — its scale_start_index therefore RESTARTS AT 0 in every
    group.  groups = [(n_cams, [(H, W), ...]), ...]."""
    g = torch.Generator().manual_seed(seed)
    nested = [[torch.randn(bs, k, C, h, w, generator=g) for h, w in sizes] for k, sizes in groups]
    col, shape, start = ops.feature_maps_format(nested)
    back = ops.feature_maps_format([col, shape, start], inverse=True)
    d = dict(col=col.numpy(), shape=shape.numpy(), start=start.numpy(), n_groups=np.int64(len(groups)))
    for gi, (k, sizes) in enumerate(groups):
        d["g%d_cams" % gi] = np.int64(k)
        d["g%d_sizes" % gi] = np.array(sizes, np.int64)
        for l, m in enumerate(nested[gi]):
            d["g%d_map%d" % (gi, l)] = m.numpy()
            assert torch.equal(back[gi][l], m)      # the reference's inverse ignores the start table
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, col.shape, shape.tolist(), start.tolist())


def module_case(blocks, name, seed, bs, A, embed, groups, sizes, cams, n_learn, camera_embed,
                residual):
    torch.manual_seed(seed)
    dfa = make_dfa(blocks, embed, groups, len(sizes), cams, n_learn, FIX7, camera_embed, residual)
    dfa.eval()
    g = torch.Generator().manual_seed(seed + 1)
    inst = torch.randn(bs, A, embed, generator=g)
    emb = torch.randn(bs, A, embed, generator=g)
    anchor = torch.randn(bs, A, 11, generator=g)
    anchor[..., :2] *= 8.0
    anchor[..., 3:6] *= 0.3
    maps = [torch.randn(bs, cams, embed, h, w, generator=g) for h, w in sizes]
    # toy pinhole cameras looking along +y / -y / +x, 64x32 image
    proj = torch.zeros(bs, cams, 4, 4)
    for k in range(cams):
        a = 2 * np.pi * k / cams
        fwd = np.array([-np.sin(a), np.cos(a), 0.0])
        right = np.array([np.cos(a), np.sin(a), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        E = np.eye(4)
        E[:3, :3] = np.stack([right, down, fwd])
        Km = np.eye(4)
        Km[0, 0] = Km[1, 1] = 30.0
        Km[0, 2], Km[1, 2] = 32.0, 16.0
        proj[:, k] = torch.tensor(Km @ E, dtype=torch.float32)
    wh = torch.tensor([64.0, 32.0])[None, None].repeat(bs, cams, 1)
    metas = dict(projection_mat=proj, image_wh=wh)
    with torch.no_grad():
        out = dfa(inst, anchor, emb, maps, metas)
        kp = dfa.kps_generator(anchor, inst)
        w = dfa._get_weights(inst, emb, metas)
        uv = dfa.project_points(kp, proj, wh)
    # training mode: the reference draws its attn-drop mask with torch.rand on the CPU generator
    # (models/blocks.py:188-195); the same draw is repeated here and stored so that tests can inject it
    dfa.train()
    torch.manual_seed(seed + 2)
    with torch.no_grad():
        w_train = dfa._get_weights(inst, emb, metas)
    dfa.eval()
    torch.manual_seed(seed + 2)
    keep = torch.rand(bs, A, cams, 1, w.shape[4], 1) > 0.15
    assert torch.equal(w_train, (keep * w) / (1 - 0.15)), "mask replay does not match the reference draw"
    d = dict(instance_feature=inst.numpy(), anchor_embed=emb.numpy(), anchor=anchor.numpy(),
             projection_mat=proj.numpy(), image_wh=wh.numpy(), out=out.numpy(),
             key_points=kp.numpy(), weights=w.numpy(), points_2d=uv.numpy(),
             train_weights=w_train.numpy(), train_keep=keep[:, :, :, 0, :, 0].numpy().astype(np.uint8),
             sizes=np.array(sizes, np.int64),
             cfg=np.array([embed, groups, len(sizes), cams, n_learn, int(camera_embed),
                           int(residual == "cat")], np.int64))
    for l, m in enumerate(maps):
        d["map%d" % l] = m.numpy()
    for k, v in dfa.state_dict().items():
        d["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, out.shape, sorted(dfa.state_dict().keys()))


def op_full_size_case(blocks, ops, name, seed):
    """BASELINE.json config #1 at full size: the reference's grid_sample path (`feature_sampling` +
    `multi_view_level_fusion`, with the op's mask multiplied in) at the SimPB R50 704x256 shape —
    bs=1, 6 cameras, 4 levels, 900 anchors, 13 key points, 256 channels / 8 groups — on the seeded
    camera-rig inputs the benchmark uses (simpb_b200.synthetic.rig_op_inputs).  Only the output is
    stored (0.9 MB); the inputs are regenerated from the seed by the test."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from simpb_b200 import synthetic
    d = synthetic.rig_op_inputs(bs=1, seed=seed)
    maps = ops.feature_maps_format([d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"]],
                                   inverse=True)[0]
    loc, w = d["sampling_location"], d["weights"]                       # [1,A,P,K,2], [1,A,P,K,L,G]
    bs, A, P, K, _ = loc.shape
    L, G = w.shape[4:6]
    C = d["mc_ms_feat"].shape[-1]
    w_klp = w.permute(0, 1, 3, 4, 2, 5).contiguous()                    # [1,A,K,L,P,G] (blocks.py:133-144 inverse)
    dfa = make_dfa(blocks, C, G, L, K, 0, [[0, 0, 0]] * P, False, "add")
    # gradients too (autograd through the reference path, one camera at a time to bound memory);
    # the two large ones are stored as reductions: grad_weights summed over (p,k,l) per (anchor,
    # group), grad_feat summed over the pixels of every (camera, level) per channel
    maps = [m.requires_grad_() for m in maps]
    loc = loc.clone().requires_grad_()
    w_klp = w_klp.requires_grad_()
    go = d["grad_output"]
    total = 0
    for k in range(K):
        kp = torch.cat([loc[:, :, :, k], torch.ones_like(loc[:, :, :, k, :1])], -1)
        proj = torch.eye(4)[None, None].repeat(bs, K, 1, 1)
        f = blocks.DeformableFeatureAggregation.feature_sampling(maps, kp, proj, None)
        x, y = loc[:, :, :, k, 0], loc[:, :, :, k, 1]
        m = ((x > 0) & (x < 1) & (y > 0) & (y < 1)).float()          # op mask, .cu:168-171
        f = f * m[:, :, None, None, :, None]
        sel = torch.zeros(K)
        sel[k] = 1
        part = dfa.multi_view_level_fusion(f, w_klp * sel[None, None, :, None, None, None]).sum(dim=2)
        part.backward(go)                                            # linear in `part`: per-camera backward
        total = total + part.detach()
    gw = w_klp.grad.permute(0, 1, 4, 2, 3, 5)                            # [1,A,P,K,L,G]
    gfeat = torch.stack([torch.stack([mp.grad[0, k].sum(dim=(1, 2)) for mp in maps]) for k in range(K)])  # [K,L,C]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), out=total.numpy(), seed=np.int64(seed),
                        grad_loc=loc.grad.numpy(), grad_weights_sum=gw.sum(dim=(2, 3, 4)).numpy(),
                        grad_feat_level_sum=gfeat.numpy())
    print(name, "out", total.shape, "abs max", float(total.abs().max()), "grad_loc abs max",
          float(loc.grad.abs().max()))


def module_released_case(blocks, name, seed, n_keep=48):
    """The released SimPB+ R50 configuration (projects/configs/simpb_nus_r50_img_704x256.py:216-239:
    256 channels, 8 groups, 4 levels, 6 cameras, 7 fixed + 6 learnable key points, camera embedding,
    residual "cat") at full size — bs=1, 900 anchors, R50 704x256 maps, camera-rig inputs from
    simpb_b200.synthetic.module_inputs_rig(seed).  Parameters come from tests/helpers.seeded_state_dict
    (a 1 MB checkpoint would not be a small fixture), inputs are regenerated from the seed; stored are the
    reference module's key points and, for the first n_keep anchors, its projected points, attention
    weights and output (grid_sample path)."""
    root = os.path.dirname(os.path.dirname(OUT))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from simpb_b200 import synthetic
    from helpers import seeded_state_dict
    dfa = make_dfa(blocks, 256, 8, 4, 6, 6, FIX7, True, "cat")
    dfa.load_state_dict(seeded_state_dict(dfa, seed))
    dfa.eval()
    d = synthetic.module_inputs_rig(bs=1, seed=seed)
    metas = dict(projection_mat=d["projection_mat"], image_wh=d["image_wh"])
    with torch.no_grad():
        out = dfa(d["instance_feature"], d["anchor"], d["anchor_embed"], d["feature_maps"], metas)
        kp = dfa.kps_generator(d["anchor"], d["instance_feature"])
        w = dfa._get_weights(d["instance_feature"], d["anchor_embed"], metas)
        uv = dfa.project_points(kp, d["projection_mat"], d["image_wh"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=np.int64(seed), n_keep=np.int64(n_keep),
                        key_points=kp.numpy(), points_2d=uv[:, :, :n_keep].numpy(),
                        weights=w[:, :n_keep].numpy(), out=out[:, :n_keep].numpy(),
                        out_abs_max=np.float64(out.abs().max()))
    # every anchor's aggregated + projected features (the first 256 channels; the rest of the "cat"
    # output is the instance feature itself), for the full-size check of the fused forward
    np.savez_compressed(os.path.join(OUT, name + "_out.npz"), seed=np.int64(seed),
                        out=out[..., :256].numpy())
    print(name, "out", out.shape, "abs max", float(out.abs().max()), "weights", w.shape)


def msda_case(name, seed, bs, Q, M, D, sizes, P):
    """Multi-scale deformable attention (the reference's 2-D branch calls mmcv-full 1.7.1's
    MultiScaleDeformableAttnFunction, projects/mmdet3d_plugin/models/group_attn.py:229-233; mmcv is
    not vendored and not installable here).  The fixture comes from an INDEPENDENT third-party
    implementation of the same published function that this image does ship: HuggingFace transformers'
    `MultiScaleDeformableAttention.forward` (Deformable-DETR's pure-PyTorch path, line for line the
    algorithm of mmcv's `multi_scale_deformable_attn_pytorch`), forward and autograd gradients in fp64."""
    from transformers.models.deformable_detr.modeling_deformable_detr import MultiScaleDeformableAttention
    import transformers
    g = torch.Generator().manual_seed(seed)
    shapes = torch.tensor(sizes, dtype=torch.int64)
    counts = shapes[:, 0] * shapes[:, 1]
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])
    S, L = int(counts.sum()), len(sizes)
    value = torch.randn(bs, S, M, D, generator=g)
    loc = torch.rand(bs, Q, M, L, P, 2, generator=g) * 1.3 - 0.15
    w = torch.rand(bs, Q, M, L, P, generator=g).reshape(bs, Q, M, -1).softmax(-1).reshape(bs, Q, M, L, P)
    go = torch.randn(bs, Q, M * D, generator=g)
    v64, l64, w64 = (t.double().requires_grad_() for t in (value, loc, w))
    out = MultiScaleDeformableAttention()(v64, shapes, [tuple(x) for x in sizes], start, l64, w64, 64)
    out.backward(go.double())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), value=value.numpy(), loc=loc.numpy(), w=w.numpy(),
                        go=go.numpy(), sizes=np.array(sizes, np.int64), out=out.detach().numpy(),
                        grad_value=v64.grad.numpy(), grad_loc=l64.grad.numpy(), grad_w=w64.grad.numpy(),
                        transformers_version=np.array(transformers.__version__))
    print(name, "out", out.shape, "abs max", float(out.abs().max()), "transformers", transformers.__version__)


if __name__ == "__main__":
    blocks, ops = import_reference()
    sizes = [(8, 12), (4, 6), (2, 3)]
    if "--round2" in sys.argv:      # only the fixtures added in round 2 (the others are unchanged)
        flatten_nested_case(ops, "flatten_nested", 11, bs=2, C=8,
                            groups=[(2, [(6, 10), (3, 5)]), (1, [(4, 6), (2, 3)]), (3, [(6, 10), (3, 5)])])
        module_released_case(blocks, "module_released_r50", 88)
        sys.exit(0)
    flatten_nested_case(ops, "flatten_nested", 11, bs=2, C=8,
                        groups=[(2, [(6, 10), (3, 5)]), (1, [(4, 6), (2, 3)]), (3, [(6, 10), (3, 5)])])
    flatten_case(ops, "flatten_small", 1, bs=2, K=3, C=16, sizes=sizes)
    # op vs reference fallback with the op's mask multiplied in, fp64 + fp32
    op_case(blocks, "op_masked_f64", 2, 2, 7, 5, 3, sizes, 32, 4, -0.15, 1.15, True, torch.float64)
    op_case(blocks, "op_masked_f32", 3, 2, 7, 5, 3, sizes, 32, 4, -0.15, 1.15, True, torch.float32)
    # unmasked fallback on inputs that stay clear of the half-pixel border band
    op_case(blocks, "op_inner_f64", 4, 1, 9, 4, 2, sizes, 64, 8, 0.26, 0.74, False, torch.float64)
    module_case(blocks, "module_cat_cam", 5, bs=2, A=11, embed=64, groups=4, sizes=sizes, cams=3,
                n_learn=2, camera_embed=True, residual="cat")
    op_full_size_case(blocks, ops, "op_r50_rig_full", 77)
    module_released_case(blocks, "module_released_r50", 88)
    msda_case("msda_hf", 9, bs=2, Q=17, M=8, D=32, sizes=[(6, 10), (3, 5), (2, 3)], P=4)
    module_case(blocks, "module_add_nocam", 6, bs=1, A=6, embed=32, groups=2, sizes=sizes, cams=2,
                n_learn=0, camera_embed=False, residual="add")
