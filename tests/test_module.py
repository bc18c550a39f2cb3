"""The module-level mirror (simpb_b200/blocks.py): interface and state-dict compatibility on CPU,
parity of the fused front end and of the whole module on the GPU."""
import numpy as np
import pytest
import torch

from oracle import module_ref
from helpers import RTOL_F32, assert_close, load_golden, seeded_state_dict

GOLDEN = ["module_cat_cam", "module_add_nocam"]


def build_from_golden(g, **over):
    from simpb_b200 import blocks
    embed, groups, n_levels, cams, n_learn, cam_embed, cat = [int(v) for v in g["cfg"]]
    kw = dict(embed_dims=embed, num_groups=groups, num_levels=n_levels, num_cams=cams, attn_drop=0.15,
              use_deformable_func=True, use_camera_embed=bool(cam_embed),
              residual_mode="cat" if cat else "add",
              kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=n_learn,
                                 fix_scale=g["sd.kps_generator.fix_scale"].tolist()))
    kw.update(over)
    m = blocks.DeformableFeatureAggregation(**kw)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    return m, sd


def ref_from_golden(g, op="grid_sample_masked"):
    embed, groups, n_levels, cams, n_learn, cam_embed, cat = [int(v) for v in g["cfg"]]
    m = module_ref.DFAModuleRef(embed, groups, n_levels, cams, attn_drop=0.15,
                                fix_scale=g["sd.kps_generator.fix_scale"].tolist(),
                                num_learnable_pts=n_learn, use_camera_embed=bool(cam_embed),
                                residual_mode="cat" if cat else "add", op=op)
    m.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")})
    return m


def inputs(g, device="cpu", dtype=torch.float32):
    t = {k: torch.from_numpy(g[k]).to(device=device, dtype=dtype)
         for k in ("instance_feature", "anchor", "anchor_embed", "projection_mat", "image_wh")}
    n_levels = int(g["cfg"][2])
    maps = [torch.from_numpy(g["map%d" % l]).to(device=device, dtype=dtype) for l in range(n_levels)]
    return t, maps


# ------------------------------------------------------------------ CPU: interface
@pytest.mark.parametrize("name", GOLDEN)
def test_state_dict_keys_match_the_reference_module(name):
    g = load_golden(name)
    m, sd = build_from_golden(g)
    mine = m.state_dict()
    assert set(mine) == set(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd)      # a reference checkpoint loads as is


def test_released_config_builds_with_reference_parameter_count():
    from simpb_b200 import blocks, synthetic
    m = blocks.DeformableFeatureAggregation(
        embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.15,
        use_deformable_func=True, use_camera_embed=True, residual_mode="cat",
        kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                           fix_scale=synthetic.FIX_SCALE))
    assert m.num_pts == 13
    assert sum(p.numel() for p in m.parameters()) == 247495
    m.init_weight()
    assert float(m.weights_fc.weight.abs().max()) == 0.0          # blocks.py:106-108


def test_there_is_no_grid_sample_path():
    from simpb_b200 import blocks
    with pytest.raises(ValueError):
        blocks.DeformableFeatureAggregation(use_deformable_func=False,
                                            kps_generator=dict(type="SparseBox3DKeyPointsGenerator"))
    m = blocks.DeformableFeatureAggregation(kps_generator=dict(type="SparseBox3DKeyPointsGenerator"))
    x = torch.zeros(1, 2, 256)
    with pytest.raises(Exception):          # CPU tensors: the C ABI wrapper refuses them
        m(x, torch.zeros(1, 2, 11), x, [torch.zeros(1, 4, 256), torch.zeros(6, 4, 2), torch.zeros(6, 4)],
          dict(projection_mat=torch.zeros(1, 6, 4, 4)))


@pytest.mark.parametrize("name", GOLDEN)
def test_key_point_generator_standalone_matches_reference(name):
    """Calling the generator module directly returns the reference's 3-D key points."""
    g = load_golden(name)
    m, sd = build_from_golden(g)
    m.load_state_dict(sd)
    t, _ = inputs(g)
    with torch.no_grad():
        kp = m.kps_generator(t["anchor"], t["instance_feature"])
    assert_close(kp, g["key_points"], 1e-6, "key points")


def test_temporal_key_points_and_anchor_projection_follow_the_reference_formulas():
    from simpb_b200 import blocks
    gen = blocks.SparseBox3DKeyPointsGenerator(embed_dims=16, num_learnable_pts=0,
                                               fix_scale=[[0, 0, 0], [0.45, 0, 0]])
    g = torch.Generator().manual_seed(0)
    anchor = torch.randn(2, 5, 11, generator=g)
    T = torch.eye(4)[None].repeat(2, 1, 1)
    T[:, :3, 3] = torch.randn(2, 3, generator=g)
    now, then = torch.tensor([1.0, 2.0]), torch.tensor([0.5, 1.0])
    pts, past = gen(anchor, None, [T], now, [then])
    dt = (now - then)[:, None, None]
    want = pts - (anchor[..., 8:] * dt)[:, :, None] + T[:, None, None, :3, 3]
    assert_close(past[0], want, 1e-6, "temporal key points (identity rotation)")
    out = gen.anchor_projection(anchor, [T], time_intervals=[now - then])[0]
    assert out.shape == anchor.shape
    center = anchor[..., :3] - anchor[..., 8:] * dt + T[:, None, :3, 3]
    assert_close(out[..., :3], center, 1e-6, "projected centre")
    assert_close(out[..., 3:6], anchor[..., 3:6], 0, "sizes unchanged")
    assert_close(gen.distance(anchor), anchor[..., :2].norm(dim=-1), 1e-7, "distance")


# ------------------------------------------------------------------ GPU: parity
@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_fused_front_end_matches_reference_fixtures(name):
    """sampling locations and attention weights against the tensors the unmodified reference
    module produced (tests/golden/make_golden.py)."""
    g = load_golden(name)
    m, sd = build_from_golden(g)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    t, _ = inputs(g, "cuda")
    with torch.no_grad():
        loc, w = m.sampling_and_weights(t["instance_feature"], t["anchor"], t["anchor_embed"],
                                        dict(projection_mat=t["projection_mat"], image_wh=t["image_wh"]))
    ref_w = torch.from_numpy(g["weights"]).permute(0, 1, 4, 2, 3, 5)          # [bs,A,P,K,L,G]
    assert_close(w, ref_w, RTOL_F32, "weights")
    ref_uv = torch.from_numpy(g["points_2d"]).permute(0, 2, 3, 1, 4)          # [bs,A,P,K,2]
    sel = (ref_uv.abs() < 4).all(-1)
    assert (loc.cpu() - ref_uv)[sel].abs().max() < 1e-4
    mine, ref = module_ref.op_valid_mask(loc.cpu()), module_ref.op_valid_mask(ref_uv)
    flips = mine != ref
    if flips.any():       # only points within rounding distance of a border may differ
        dist = torch.minimum(ref_uv.abs(), (ref_uv - 1).abs()).min(-1).values
        assert (dist[flips] < 1e-5).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_module_forward_matches_masked_reference_path(name):
    from simpb_b200 import feature_maps_format
    g = load_golden(name)
    m, sd = build_from_golden(g)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    ref = ref_from_golden(g).eval()
    t, maps = inputs(g, "cuda")
    tc, maps_c = inputs(g)
    with torch.no_grad():
        out = m(t["instance_feature"], t["anchor"], t["anchor_embed"],
                feature_maps_format(maps), dict(projection_mat=t["projection_mat"], image_wh=t["image_wh"]))
        want = ref(tc["instance_feature"], tc["anchor"], tc["anchor_embed"], maps_c,
                   tc["projection_mat"], tc["image_wh"])
    assert_close(out, want, RTOL_F32, "module output")


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_training_mode_attn_drop_matches_reference_fixtures(name):
    """Softmax + attn-drop + permute kernel in training mode against the weights the unmodified
    reference module produced with the keep mask it drew (stored in the fixture)."""
    g = load_golden(name)
    m, sd = build_from_golden(g)
    m.load_state_dict(sd)
    m = m.cuda().train()
    t, _ = inputs(g, "cuda")
    keep = torch.from_numpy(g["train_keep"]).cuda()
    loc, w = m.sampling_and_weights(t["instance_feature"], t["anchor"], t["anchor_embed"],
                                    dict(projection_mat=t["projection_mat"], image_wh=t["image_wh"]), keep=keep)
    ref = torch.from_numpy(g["train_weights"]).permute(0, 1, 4, 2, 3, 5)      # [bs,A,P,K,L,G]
    assert torch.equal(w.detach().cpu() == 0, ref == 0)
    assert_close(w.detach(), ref, RTOL_F32, "training-mode weights")


@pytest.mark.gpu
def test_released_config_full_size_matches_the_reference_module():
    """Released SimPB+ R50 configuration at full size on the GPU (fused inference kernel and the
    three-kernel path): sampling locations and attention weights against the tensors of the unmodified
    reference module (tests/golden/module_released_r50.npz), the module output against the masked
    grid_sample restatement, which the CPU suite pins to the same fixture."""
    from simpb_b200 import blocks, feature_maps_format, synthetic
    from test_oracle_golden import RELEASED, released_ref
    g = load_golden("module_released_r50")
    n, seed = int(g["n_keep"]), int(g["seed"])
    m = blocks.DeformableFeatureAggregation(
        embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.15, use_deformable_func=True,
        use_camera_embed=True, residual_mode="cat",
        kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6, fix_scale=RELEASED["fix_scale"]))
    m.load_state_dict(seeded_state_dict(m, seed))
    m = m.cuda().eval()
    d = synthetic.module_inputs_rig(bs=1, seed=seed)
    t = {k: v.cuda() for k, v in d.items() if isinstance(v, torch.Tensor)}
    metas = dict(projection_mat=t["projection_mat"], image_wh=t["image_wh"])
    with torch.no_grad():
        loc, w = m.sampling_and_weights(t["instance_feature"], t["anchor"], t["anchor_embed"], metas)
        out = m(t["instance_feature"], t["anchor"], t["anchor_embed"],
                feature_maps_format([x.cuda() for x in d["feature_maps"]]), metas)
        want = released_ref(seed, "grid_sample_masked")(
            d["instance_feature"], d["anchor"], d["anchor_embed"], d["feature_maps"], d["projection_mat"],
            d["image_wh"])
    ref_w = torch.from_numpy(g["weights"]).permute(0, 1, 4, 2, 3, 5)          # [1,n,P,K,L,G]
    assert_close(w[:, :n], ref_w, RTOL_F32, "weights")
    ref_uv = torch.from_numpy(g["points_2d"]).permute(0, 2, 3, 1, 4)          # [1,n,P,K,2]
    sel = (ref_uv.abs() < 4).all(-1)
    assert (loc[:, :n].cpu() - ref_uv)[sel].abs().max() < 1e-4
    assert_close(out, want, RTOL_F32, "module output, released config")
    # all 900 anchors against the REFERENCE MODULE's own output (grid_sample path, module_released_r50_out
    # .npz): anchors whose samples stay clear of grid_sample's half-pixel border band — where the op's
    # exclusive (0,1) mask and grid_sample differ by design — must agree to 1e-5
    ref_out = torch.from_numpy(load_golden("module_released_r50_out")["out"])          # [1,900,256]
    bx, by = 0.5 / 22 + 1e-6, 0.5 / 8 + 1e-6      # half a pixel of the coarsest level (8 x 22)
    l = loc.cpu()
    x, y = l[..., 0], l[..., 1]
    inside = (x > 0) & (x < 1) & (y > 0) & (y < 1)
    in_band = (x > -bx) & (x < 1 + bx) & (y > -by) & (y < 1 + by)      # grid_sample still sees the map
    risky = (in_band & ~inside).flatten(2).any(-1)[0]                                   # [900] anchors
    clean = ~risky
    assert int(clean.sum()) > 600
    scale = ref_out.abs().max()
    err = (out[..., :256].cpu() - ref_out)[0, clean].abs().max() / scale
    assert float(err) <= RTOL_F32, "fused forward vs reference module, %d clean anchors: %.3e" % (int(clean.sum()), err)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_module_backward_matches_torch_autograd_of_the_reference_path(name):
    """Every gradient the module produces (inputs, feature maps, all parameters) against fp64
    autograd through the reference's grid_sample path with the op's mask, including a fixed
    attn-drop keep mask."""
    from simpb_b200 import feature_maps_format
    g = load_golden(name)
    cams, n_learn = int(g["cfg"][3]), int(g["cfg"][4])
    m, sd = build_from_golden(g)
    m.load_state_dict(sd)
    m = m.cuda().train()
    ref = ref_from_golden(g).double().train()
    t, maps = inputs(g, "cuda")
    tc, maps_c = inputs(g, dtype=torch.float64)
    bs, A = t["anchor"].shape[:2]
    gen = torch.Generator().manual_seed(3)
    keep = torch.rand(bs, A, cams, m.num_pts, generator=gen) > 0.15
    go = torch.randn(*g["out"].shape, generator=gen)
    names = ("instance_feature", "anchor", "anchor_embed")
    for d in (t, tc):
        for n in names:
            d[n].requires_grad_()
    for x in maps + maps_c:
        x.requires_grad_()
    out = m(t["instance_feature"], t["anchor"], t["anchor_embed"], feature_maps_format(maps),
            dict(projection_mat=t["projection_mat"], image_wh=t["image_wh"]), attn_keep_mask=keep.cuda())
    out.backward(go.cuda())
    want = ref(tc["instance_feature"], tc["anchor"], tc["anchor_embed"], maps_c, tc["projection_mat"],
               tc["image_wh"], drop_mask=keep[:, :, :, None, :, None])
    want.backward(go.double())
    assert_close(out, want, RTOL_F32, "training-mode output")
    for n in names:
        got, exp = t[n].grad, tc[n].grad
        if n == "anchor":      # velocity entries do not enter the module
            assert float(got[..., 8:].abs().max()) == 0.0
            got, exp = got[..., :8], exp[..., :8]
        assert_close(got, exp, 2e-5, "grad " + n)
    for l, (a, b) in enumerate(zip(maps, maps_c)):
        assert_close(a.grad, b.grad, RTOL_F32, "grad feature map %d" % l)
    ref_params = dict(ref.named_parameters())
    for n, p in m.named_parameters():
        if n == "kps_generator.fix_scale":
            continue
        assert p.grad is not None, n
        assert_close(p.grad, ref_params[n].grad, 2e-5, "grad " + n)
    assert n_learn == 0 or m.kps_generator.learnable_fc.weight.grad.abs().max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_inference_kernel_matches_the_three_kernel_path(dtype):
    """dfa_forward_fused (key points + projection + softmax + gather in one launch) against the
    separate kernels at the released config: identical sampling locations, outputs to rounding; and
    the module takes the fused path exactly when no gradient is needed."""
    from simpb_b200 import blocks, cabi, feature_maps_format, synthetic
    torch.manual_seed(0)
    m = blocks.DeformableFeatureAggregation(
        embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.15, use_camera_embed=True,
        residual_mode="cat", kps_generator=dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                                                fix_scale=synthetic.FIX_SCALE)).cuda().eval()
    d = synthetic.module_inputs_rig(bs=2, A=300, levels=((16, 44), (8, 22), (4, 11), (2, 6)), seed=11)
    g = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
    fm = feature_maps_format([x.cuda() for x in d["feature_maps"]], dtype=dtype)
    metas = dict(projection_mat=g["projection_mat"], image_wh=g["image_wh"])
    args = (g["instance_feature"], g["anchor"], g["anchor_embed"])
    with torch.no_grad():
        loc, w = m.sampling_and_weights(*args, metas)
        want = cabi.forward(fm[0], fm[1].int(), fm[2].int(), loc, w)
        got = m.fused_features(*args, fm, metas)
        assert got is not None
        assert_close(got, want, 2e-6, "fused features")
        gen = m.kps_generator
        cam = m.camera_encoder(metas["projection_mat"][:, :, :3].reshape(2, 6, -1))
        out2, loc2 = cabi.forward_fused(
            fm[0], fm[1].int(), fm[2].int(), g["anchor"], gen.fix_scale, gen.offset_logits(g["instance_feature"]),
            g["projection_mat"], g["image_wh"], m.weights_fc(g["instance_feature"] + g["anchor_embed"]),
            torch.nn.functional.linear(cam, m.weights_fc.weight).contiguous(), (2, 300, 6, 4, 13, 8),
            want_locations=True)
        assert torch.equal(loc2, loc)                       # same device code: bit-identical locations
        assert torch.equal(out2, got)
        y_fused = m(*args, fm, metas)                       # eval + no_grad: fused path
    y_sep = m(*args, fm, metas)                             # parameters require grad: separate kernels
    assert_close(y_fused, y_sep, 2e-6, "module output, fused vs separate")
    # full (unsplit) logits through the same entry point
    m2 = blocks.DeformableFeatureAggregation(embed_dims=256, num_groups=8, num_levels=4, num_cams=6,
                                             kps_generator=dict(type="SparseBox3DKeyPointsGenerator",
                                                                fix_scale=synthetic.FIX_SCALE)).cuda().eval()
    with torch.no_grad():
        loc, w = m2.sampling_and_weights(*args, metas)
        want = cabi.forward(fm[0], fm[1].int(), fm[2].int(), loc, w)
        assert_close(m2.fused_features(*args, fm, metas), want, 2e-6, "fused features, no camera embedding")


@pytest.mark.gpu
def test_softmax_weights_kernel_r50_shape_vs_torch():
    from simpb_b200 import cabi
    gen = torch.Generator().manual_seed(1)
    bs, A, K, L, P, G = 2, 900, 6, 4, 13, 8
    logits = (3 * torch.randn(bs, A, K, L * P * G, generator=gen)).cuda()
    keep = (torch.rand(bs, A, K, P, generator=gen) > 0.15)
    w = cabi.softmax_weights(logits, (bs, A, K, L, P, G))
    ref = logits.double().reshape(bs, A, -1, G).softmax(-2).reshape(bs, A, K, L, P, G)
    assert_close(w, ref.permute(0, 1, 4, 2, 3, 5), 1e-6, "softmax + permute")
    assert_close(w.sum(dim=(2, 3, 4)), torch.ones(bs, A, G), 1e-5, "rows sum to one")
    wk = cabi.softmax_weights(logits, (bs, A, K, L, P, G), keep.to(torch.uint8).cuda(), 1 / 0.85)
    refk = ref * keep[:, :, :, None, :, None].cuda() / 0.85
    assert_close(wk, refk.permute(0, 1, 4, 2, 3, 5), 1e-6, "softmax + keep mask + permute")
    # backward against autograd
    x = logits.double().requires_grad_()
    y = (x.reshape(bs, A, -1, G).softmax(-2).reshape(bs, A, K, L, P, G)
         * keep[:, :, :, None, :, None].cuda() / 0.85).permute(0, 1, 4, 2, 3, 5)
    gw = torch.randn(bs, A, P, K, L, G, generator=gen).cuda()
    y.backward(gw.double())
    gx = cabi.softmax_weights_backward(logits, (bs, A, K, L, P, G), keep.to(torch.uint8).cuda(), 1 / 0.85, gw)
    assert_close(gx, x.grad.reshape(gx.shape), 1e-5, "softmax backward")


@pytest.mark.gpu
def test_keypoints_project_backward_vs_torch_autograd():
    from simpb_b200 import cabi, synthetic
    d = synthetic.module_inputs_rig(bs=2, A=300, seed=8, feat=False)
    gen = torch.Generator().manual_seed(9)
    logits = torch.randn(2, 300, 18, generator=gen)
    fix = torch.tensor(synthetic.FIX_SCALE)
    a64 = d["anchor"].double().requires_grad_()
    l64 = logits.double().requires_grad_()
    kp = module_ref.key_points(a64, fix.double(), l64)
    uv = module_ref.project_points(kp, d["projection_mat"].double(), d["image_wh"].double())
    uv = uv.permute(0, 2, 3, 1, 4)                                             # [bs,A,P,K,2]
    # weight the gradient towards well-conditioned projections (in front of the camera, O(1))
    ok = ((uv.detach().abs() < 3).all(-1, keepdim=True)).double()
    go = torch.randn(uv.shape, generator=gen).double() * ok
    (uv * go).sum().backward()
    ga, gl = cabi.keypoints_project_backward(d["anchor"].cuda(), fix.cuda(), logits.cuda(),
                                             d["projection_mat"].cuda(), d["image_wh"].cuda(),
                                             go.float().cuda())
    assert float(ga[..., 8:].abs().max()) == 0.0
    assert_close(ga[..., :8], a64.grad[..., :8], 2e-5, "grad anchor")
    assert_close(gl, l64.grad, 2e-5, "grad learnable-offset logits")
