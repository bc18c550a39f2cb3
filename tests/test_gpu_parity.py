"""Parity of the CUDA path (through the C ABI) with the oracle, the golden fixtures and — when
the prebuilt object is present — the unmodified reference CUDA op."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import module_ref
from helpers import (RTOL_BF16, RTOL_F32, assert_close, golden_op_inputs, load_golden, rel_err)

pytestmark = pytest.mark.gpu


def dev(d, dtype=torch.float32):
    """Move an op-input dict to the GPU in the dtypes the C ABI takes."""
    from simpb_b200 import synthetic  # noqa: F401
    g = {}
    g["feat"] = d["mc_ms_feat"].cuda().to(dtype).contiguous()
    g["shape"] = d["spatial_shape"].int().cuda()
    g["start"] = d["scale_start_index"].int().cuda()
    g["loc"] = d["sampling_location"].cuda().contiguous()
    g["w"] = d["weights"].cuda().contiguous()
    g["go"] = d["grad_output"].cuda().contiguous()
    return g


def small_case(seed, bs, A, P, K, sizes, C, G, lo=-0.15, hi=1.15):
    from simpb_b200 import synthetic
    return synthetic.op_inputs_uniform(bs=bs, A=A, P=P, K=K, levels=sizes, C=C, G=G, seed=seed,
                                       lo=lo, hi=hi)


def check_case(d, dtype=torch.float32, rtol=RTOL_F32, backward=True):
    from simpb_b200 import cabi
    g = dev(d, dtype)
    feat_ref = g["feat"].float().cpu()      # the oracle sees exactly the values the kernel sees
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    ref = oracle.forward(feat_ref, d["spatial_shape"], d["scale_start_index"],
                         d["sampling_location"], d["weights"])
    assert_close(out, ref, rtol if dtype == torch.float32 else RTOL_F32, "forward")
    if not backward:
        return
    rgf, rgl, rgw = oracle.backward(feat_ref, d["spatial_shape"], d["scale_start_index"],
                                    d["sampling_location"], d["weights"], d["grad_output"])
    # (1) library-managed buffers: small gradients fully written, grad_feat zeroed by the library
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    assert_close(gf, rgf, rtol, "grad_feat")
    assert_close(gl, rgl, rtol, "grad_loc")
    assert_close(gw, rgw, rtol, "grad_weights")
    # (2) reference contract: accumulate into caller-provided buffers (here pre-filled with 1)
    gf2 = torch.ones_like(gf); gl2 = torch.ones_like(gl); gw2 = torch.ones_like(gw)
    cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], gf2, gl2, gw2)
    assert_close(gf2 - 1, rgf, 4 * rtol, "grad_feat (accumulate)")
    assert_close(gl2 - 1, rgl, 4 * rtol, "grad_loc (accumulate)")
    assert_close(gw2 - 1, rgw, 4 * rtol, "grad_weights (accumulate)")


# ------------------------------------------------------------------ indices / masks: bit-exact
def tie_locations(sizes, n=4096, seed=0):
    """Locations whose pixel coordinate sits within a few ulp of an integer or of the (0,1)
    borders — where the one-rounding FMA and a two-rounding evaluation can disagree."""
    rng = np.random.default_rng(seed)
    vals = []
    for h, w in sizes:
        for size in (h, w):
            k = rng.integers(0, size, n // 8)
            base = ((k + 0.5) / size).astype(np.float32)
            for ulps in (-2, -1, 0, 1, 2):
                v = base.copy()
                for _ in range(abs(ulps)):
                    v = np.nextafter(v, np.float32(2.0 if ulps > 0 else -2.0))
                vals.append(v)
    vals.append(np.array([0.0, 1.0, np.nextafter(np.float32(0), np.float32(1)),
                          np.nextafter(np.float32(1), np.float32(0)), -0.0, 1e-30, 0.5],
                         np.float32))
    v = np.concatenate(vals)
    rng.shuffle(v)
    return v


def test_indices_and_masks_bit_exact_r50():
    from simpb_b200 import cabi, synthetic
    for maker, seed in ((synthetic.op_inputs_uniform, 0), (synthetic.rig_op_inputs, 1)):
        d = maker(bs=2, seed=seed, feat=False)
        valid, rows = cabi.debug_indices(d["spatial_shape"].int().cuda(),
                                         d["scale_start_index"].int().cuda(),
                                         d["sampling_location"].cuda())
        dummy = np.zeros((2, d["num_feat"], 8), np.float32)
        _, rv, rr = oracle.forward(dummy, d["spatial_shape"], d["scale_start_index"],
                                   d["sampling_location"], np.zeros(d["weights"].shape[:5] + (8,), np.float32),
                                   side_channel=True)
        np.testing.assert_array_equal(valid.cpu().numpy(), rv)
        np.testing.assert_array_equal(rows.cpu().numpy(), rr)


def test_indices_bit_exact_on_rounding_ties():
    from simpb_b200 import cabi, synthetic
    shape, start, num_feat = synthetic.level_tables(synthetic.R50_LEVELS, 6)
    v = tie_locations(synthetic.R50_LEVELS)
    n = (len(v) // (13 * 6 * 2)) * (13 * 6 * 2)
    loc = torch.from_numpy(v[:n].copy()).reshape(1, -1, 13, 6, 2)
    valid, rows = cabi.debug_indices(shape.int().cuda(), start.int().cuda(), loc.cuda())
    A = loc.shape[1]
    _, rv, rr = oracle.forward(np.zeros((1, num_feat, 8), np.float32), shape, start, loc,
                               np.zeros((1, A, 13, 6, 4, 8), np.float32), side_channel=True)
    np.testing.assert_array_equal(valid.cpu().numpy(), rv)
    np.testing.assert_array_equal(rows.cpu().numpy(), rr)


# ------------------------------------------------------------------ floating parity vs the oracle
SIZES3 = ((8, 12), (4, 6), (2, 3))


@pytest.mark.parametrize("cfg", [
    dict(bs=2, A=7, P=5, K=3, sizes=SIZES3, C=32, G=4),      # 2 vector lanes per group row
    dict(bs=1, A=33, P=13, K=6, sizes=SIZES3, C=256, G=8),   # released group geometry (8 lanes)
    dict(bs=3, A=5, P=4, K=2, sizes=SIZES3, C=64, G=4),      # 4 lanes per group row
    dict(bs=1, A=9, P=3, K=2, sizes=SIZES3, C=20, G=5),      # 1 lane per group row
    dict(bs=2, A=4, P=3, K=1, sizes=((5, 7),), C=6, G=3),    # generic kernels (2 channels / group)
    dict(bs=1, A=6, P=3, K=3, sizes=SIZES3, C=32, G=1),      # weights block not 16-byte sized: no TMA
    dict(bs=1, A=3, P=32, K=6, sizes=SIZES3, C=256, G=8),    # many key points
    dict(bs=1, A=2, P=2, K=2, sizes=SIZES3, C=512, G=16),    # 16 warps per anchor
])
def test_small_shapes_vs_oracle(cfg):
    check_case(small_case(11, **cfg))


def test_r50_shape_uniform_vs_oracle():
    from simpb_b200 import synthetic
    check_case(synthetic.op_inputs_uniform(bs=1, seed=0))


def test_r50_full_size_vs_reference_fallback_fixture():
    """The CUDA op (default dispatch, fp32 and bf16 tables) against the reference's own grid_sample
    path at the benchmark shape: tests/golden/op_r50_rig_full.npz holds the reference output for the
    seeded rig inputs (generated by tests/golden/make_golden.py from the unmodified reference)."""
    from simpb_b200 import cabi, synthetic
    gold = load_golden("op_r50_rig_full")
    d = synthetic.rig_op_inputs(bs=1, seed=int(gold["seed"]))
    g = dev(d)
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    assert_close(out, gold["out"], RTOL_F32, "R50 full size vs reference fallback")
    outh = cabi.forward(g["feat"].bfloat16(), g["shape"], g["start"], g["loc"], g["w"])
    assert_close(outh, gold["out"], RTOL_BF16, "R50 full size, bf16 table, vs reference fallback")
    # the three gradients (the two large ones through the reductions the fixture stores)
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    assert_close(gl, gold["grad_loc"], 5 * RTOL_F32, "R50 full size grad_loc")
    assert_close(gw.sum(dim=(2, 3, 4)), gold["grad_weights_sum"], RTOL_F32, "R50 full size grad_weights (sums)")
    shape, start = d["spatial_shape"], d["scale_start_index"]
    sums = torch.stack([torch.stack([gf[0, int(start[k, l]):int(start[k, l]) + int(shape[k, l, 0] * shape[k, l, 1])]
                                     .double().sum(dim=0) for l in range(start.shape[1])])
                        for k in range(start.shape[0])])
    assert_close(sums, gold["grad_feat_level_sum"], RTOL_F32, "R50 full size grad_feat (sums)")


def test_r50_shape_rig_vs_oracle():
    from simpb_b200 import synthetic
    check_case(synthetic.rig_op_inputs(bs=2, seed=3))


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 30, 31, 32, 33])
def test_every_forward_kernel_variant_vs_oracle(variant, monkeypatch):
    """DFA_FWD_VARIANT selects the forward kernel family / tuning point; all of them must agree
    with the oracle (fp32 and bf16 feature tables, sparse rig and dense uniform locations)."""
    from simpb_b200 import cabi, synthetic
    monkeypatch.setenv("DFA_FWD_VARIANT", str(variant))
    cases = [synthetic.rig_op_inputs(bs=1, A=300, seed=21),
             synthetic.op_inputs_uniform(bs=1, A=100, seed=22),
             small_case(23, bs=2, A=17, P=13, K=6, sizes=SIZES3, C=256, G=8),
             small_case(24, bs=1, A=9, P=5, K=2, sizes=SIZES3, C=128, G=8),
             small_case(25, bs=1, A=5, P=3, K=3, sizes=SIZES3, C=256, G=8),     # odd P*K: no TMA
             small_case(26, bs=1, A=3, P=40, K=6, sizes=SIZES3, C=256, G=8)]    # > 32 valid samples
    for d in cases:
        for dtype in (torch.float32, torch.bfloat16):
            g = dev(d, dtype)
            out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
            ref = oracle.forward(g["feat"].float().cpu(), d["spatial_shape"], d["scale_start_index"],
                                 d["sampling_location"], d["weights"])
            assert_close(out, ref, RTOL_F32, "variant %d %s" % (variant, dtype))


@pytest.mark.parametrize("anchors", [900, 950, 889])
def test_forward_channel_split(anchors, monkeypatch):
    """Row-sliced forward with channel-split CTAs: by default the anchors beyond the first wave of
    resident CTAs (888 on a 148-SM part) are shared four ways by channel blocks; DFA_FWD_SPLIT=2/3
    split every anchor.  Every mode agrees with the oracle; modes differ from each other only in
    summation order."""
    from simpb_b200 import cabi, synthetic
    for d, dtype in ((synthetic.rig_op_inputs(bs=1, A=anchors, seed=51), torch.float32),
                     (synthetic.op_inputs_uniform(bs=1, A=anchors, seed=52), torch.bfloat16)):
        g = dev(d, dtype)
        ref = oracle.forward(g["feat"].float().cpu(), d["spatial_shape"], d["scale_start_index"],
                             d["sampling_location"], d["weights"])
        outs = {}
        for mode in ("0", "1", "2", "3"):
            monkeypatch.setenv("DFA_FWD_SPLIT", mode)
            junk = torch.full((1, anchors, g["feat"].shape[2]), float("nan"), device="cuda")
            outs[mode] = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=junk)
            assert_close(outs[mode], ref, RTOL_F32, "split mode %s %s" % (mode, dtype))
            assert torch.equal(outs[mode], cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"]))
        # anchors of the first wave are untouched by the default mode
        assert torch.equal(outs["1"][:, :888], outs["0"][:, :888])      # 148 SMs x 6 CTAs


@pytest.mark.parametrize("mode", ["2", "3"])
def test_forward_channel_split_small_shapes(mode, monkeypatch):
    """Channel-split CTAs on shapes other than SimPB's (forced for every anchor): channels per block
    down to one 16-byte vector, 1..16 groups, ragged levels, no-TMA operand sizes; shapes the split
    does not fit must fall back silently."""
    monkeypatch.setenv("DFA_FWD_SPLIT", mode)
    cfgs = [dict(bs=2, A=7, P=5, K=3, sizes=SIZES3, C=32, G=2), dict(bs=1, A=9, P=4, K=2, sizes=SIZES3, C=64, G=8),
            dict(bs=1, A=5, P=13, K=6, sizes=SIZES3, C=128, G=4), dict(bs=2, A=3, P=3, K=3, sizes=SIZES3, C=256, G=8),
            dict(bs=1, A=4, P=6, K=2, sizes=SIZES3, C=512, G=16), dict(bs=1, A=6, P=2, K=2, sizes=SIZES3, C=16, G=1),
            dict(bs=1, A=3, P=7, K=5, sizes=((5, 7),), C=48, G=3), dict(bs=1, A=2, P=40, K=6, sizes=SIZES3, C=256, G=8)]
    for i, cfg in enumerate(cfgs):
        d = small_case(700 + i, **cfg)
        check_case(d, backward=False)
        check_case(d, dtype=torch.bfloat16, backward=False)


@pytest.mark.parametrize("variant", [0, 10, 11, 12])
def test_every_backward_kernel_variant_vs_oracle(variant, monkeypatch):
    """DFA_BWD_VARIANT: row-merging backward (tuning points) and the one-warp-per-group kernel, both
    buffer contracts, fp32 and bf16 feature tables, sparse / dense / many-sample anchors."""
    monkeypatch.setenv("DFA_BWD_VARIANT", str(variant))
    from simpb_b200 import synthetic
    cases = [synthetic.rig_op_inputs(bs=1, A=200, seed=31),
             synthetic.op_inputs_uniform(bs=1, A=60, seed=32),
             small_case(33, bs=2, A=17, P=13, K=6, sizes=SIZES3, C=256, G=8),
             small_case(34, bs=1, A=9, P=5, K=2, sizes=SIZES3, C=128, G=8),
             small_case(35, bs=1, A=5, P=3, K=3, sizes=SIZES3, C=256, G=8),     # odd P*K: no TMA
             small_case(36, bs=1, A=3, P=40, K=6, sizes=SIZES3, C=256, G=8)]    # several rounds
    for d in cases:
        check_case(d)
    check_case(cases[2], dtype=torch.bfloat16)


def test_fuzz_random_shapes_vs_oracle():
    """60 seeded random configurations — group counts, channels per group, levels of odd sizes,
    1..7 cameras, 1..40 key points, fp32 / bf16 tables, sparse to dense masks — forward and all three
    gradients against the oracle.  Covers every dispatch path (row-sliced, merging, per-group, generic,
    TMA and plain staging)."""
    import random
    rng = random.Random(1234)
    seen = set()
    for case in range(60):
        G = rng.choice([1, 2, 3, 4, 8, 8, 8, 16])
        cpg = rng.choice([1, 2, 4, 8, 16, 32, 32, 5])
        C = G * cpg
        if C > 512:
            continue
        L = rng.randint(1, 4)
        sizes = tuple((rng.randint(1, 9), rng.randint(1, 13)) for _ in range(L))
        K, P = rng.randint(1, 7), rng.choice([1, 2, 3, 5, 13, 13, 20, 40])
        A, bs = rng.randint(1, 12), rng.randint(1, 3)
        lo, hi = rng.choice([(-0.15, 1.15), (0.05, 0.95), (-1.0, 2.0), (0.4, 0.6)])
        d = small_case(1000 + case, bs=bs, A=A, P=P, K=K, sizes=sizes, C=C, G=G, lo=lo, hi=hi)
        dtype = torch.bfloat16 if case % 4 == 3 else torch.float32
        seen.add((C * (2 if dtype == torch.bfloat16 else 4), G))
        try:
            check_case(d, dtype=dtype)
        except AssertionError as e:
            raise AssertionError("case %d: bs=%d A=%d P=%d K=%d sizes=%s C=%d G=%d %s: %s"
                                 % (case, bs, A, P, K, sizes, C, G, dtype, e))
    assert (1024, 8) in seen and (512, 8) in seen          # the merging / row-sliced fast paths were hit


def test_training_anchor_count_vs_oracle():
    from simpb_b200 import synthetic
    check_case(synthetic.rig_op_inputs(bs=1, A=1220, seed=4))


@pytest.mark.parametrize("cfg", [
    dict(bs=2, A=7, P=5, K=3, sizes=SIZES3, C=32, G=4),
    dict(bs=1, A=33, P=13, K=6, sizes=SIZES3, C=256, G=8),
    dict(bs=1, A=5, P=3, K=2, sizes=SIZES3, C=24, G=4),      # 6 bf16 per group: generic kernel
])
def test_bf16_features_vs_oracle(cfg):
    d = small_case(12, **cfg)
    # against the oracle on the bf16-rounded values: fp32 tolerance (the arithmetic is fp32) …
    check_case(d, dtype=torch.bfloat16, rtol=RTOL_F32)
    # … and against the fp32 features: the north-star 1e-2 bound
    from simpb_b200 import cabi
    g = dev(d, torch.bfloat16)
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    ref = oracle.forward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                         d["sampling_location"], d["weights"])
    assert_close(out, ref, RTOL_BF16, "bf16 features vs fp32 oracle")


@pytest.mark.parametrize("name", ["op_masked_f64", "op_masked_f32", "op_inner_f64"])
def test_golden_reference_fixtures(name):
    """The reference's own grid_sample path (fixtures made by tests/golden/make_golden.py)."""
    from simpb_b200 import cabi
    g = load_golden(name)
    col, shape, start, gcol = golden_op_inputs(g)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().cuda()  # noqa: E731
    out = cabi.forward(col.cuda(), shape.int().cuda(), start.int().cuda(), t(g["loc"]), t(g["weights"]))
    assert_close(out, g["out"], RTOL_F32, name)
    gf, gl, gw = cabi.backward(col.cuda(), shape.int().cuda(), start.int().cuda(), t(g["loc"]),
                               t(g["weights"]), t(g["grad_out"]))
    assert_close(gf, gcol.numpy(), RTOL_F32, name + " grad_feat")
    assert_close(gw, g["grad_weights"], RTOL_F32, name + " grad_weights")
    assert_close(gl, g["grad_loc"], 5 * RTOL_F32, name + " grad_loc")


# ------------------------------------------------------------------ edge cases
def test_all_samples_masked_gives_zeros():
    from simpb_b200 import cabi
    d = small_case(5, bs=1, A=4, P=3, K=2, sizes=SIZES3, C=32, G=4, lo=1.0, hi=1.5)
    g = dev(d)
    out = torch.full((1, 4, 32), 7.0, device="cuda")
    cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=out)
    assert (out == 0).all()
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    assert (gf == 0).all() and (gl == 0).all() and (gw == 0).all()


def test_empty_inputs():
    """No anchors / empty batch: the reference launches zero threads and returns empty tensors."""
    from simpb_b200 import cabi, deformable_aggregation_function
    d = small_case(8, bs=2, A=3, P=3, K=2, sizes=SIZES3, C=64, G=8)
    g = dev(d)
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"][:, :0].contiguous(),
                       g["w"][:, :0].contiguous())
    assert tuple(out.shape) == (2, 0, 64)
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"][:, :0].contiguous(),
                               g["w"][:, :0].contiguous(), g["go"][:, :0].contiguous())
    assert tuple(gl.shape) == (2, 0, 3, 2, 2) and float(gf.abs().max()) == 0.0
    out = deformable_aggregation_function(g["feat"][:0], g["shape"], g["start"], g["loc"][:0], g["w"][:0])
    assert tuple(out.shape) == (0, 3, 64)


def test_single_anchor_single_point_single_camera():
    check_case(small_case(9, bs=1, A=1, P=1, K=1, sizes=((3, 5),), C=8, G=8))   # 1 channel per group
    check_case(small_case(10, bs=1, A=1, P=2, K=1, sizes=((1, 1),), C=256, G=8))  # 1x1 feature map


def test_r101_shape_vs_oracle():
    """BASELINE.json config #4: 1408x512 maps (359,040 rows per sample), forward and the two small
    gradients against the oracle; the feature gradient through its adjoint identity."""
    from simpb_b200 import cabi, synthetic
    d = synthetic.rig_op_inputs(bs=1, A=900, levels=synthetic.R101_LEVELS, seed=17)
    g = dev(d)
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    ref = oracle.forward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                         d["sampling_location"], d["weights"])
    assert_close(out, ref, RTOL_F32, "R101 forward")
    _, rgl, rgw = oracle.backward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                                  d["sampling_location"], d["weights"], d["grad_output"], need_feat=False)
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    assert_close(gl, rgl, RTOL_F32, "R101 grad_loc")
    assert_close(gw, rgw, RTOL_F32, "R101 grad_weights")
    # out is linear in the features: <out, go> == <feat, grad_feat>
    lhs = (torch.from_numpy(ref) * d["grad_output"].double()).sum()
    rhs = (g["feat"].double() * gf.double()).sum().cpu()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def test_feature_maps_format_camera_groups_with_different_resolutions():
    """ops/__init__.py:56-61: a list of per-group lists is formatted group by group and concatenated."""
    from simpb_b200 import cabi, feature_maps_format
    gen = torch.Generator().manual_seed(4)
    grp_a = [torch.randn(2, 2, 16, h, w, generator=gen) for h, w in ((8, 12), (4, 6))]
    grp_b = [torch.randn(2, 1, 16, h, w, generator=gen) for h, w in ((6, 10), (3, 5))]
    col, shape, start = feature_maps_format([[m.cuda() for m in grp_a], [m.cuda() for m in grp_b]])
    ca, sa, _ = module_ref.flatten_feature_maps(grp_a)
    cb, sb, _ = module_ref.flatten_feature_maps(grp_b)
    assert torch.equal(col.cpu(), torch.cat([ca, cb], dim=1))
    assert torch.equal(shape.cpu(), torch.cat([sa, sb], dim=0))
    assert start.cpu().flatten().tolist() == [0, 96, 120, 216, 240, 300]
    # the op consumes the ragged table: sample the third camera (group b) at its first level's centre
    loc = torch.full((2, 1, 1, 3, 2), -1.0)
    loc[:, :, :, 2] = 0.5
    w = torch.zeros(2, 1, 1, 3, 2, 1)
    w[:, :, :, 2, 0] = 1.0
    out = cabi.forward(col, shape.int(), start.int(), loc.cuda(), w.cuda())
    ref = oracle.forward(col.cpu(), shape.cpu(), start.cpu(), loc, w)
    assert_close(out, ref, RTOL_F32, "ragged camera groups")


def test_nested_camera_groups_vs_reference_fixture(golden_dir):
    """tests/golden/flatten_nested.npz holds what the REFERENCE's feature_maps_format returns for a list
    of three camera groups (ops/__init__.py:56-61): col_feats and spatial_shape agree bit for bit; its
    scale_start_index restarts at 0 in every group, which `reference_start_index=True` reproduces and the
    default replaces by offsets from the beginning of col_feats (differences)."""
    from simpb_b200 import feature_maps_format
    g = np.load(os.path.join(golden_dir, "flatten_nested.npz"))
    nested = [[torch.from_numpy(g["g%d_map%d" % (gi, l)]).cuda() for l in range(len(g["g%d_sizes" % gi]))]
              for gi in range(int(g["n_groups"]))]
    col, shape, start = feature_maps_format(nested, reference_start_index=True)
    assert torch.equal(col.cpu(), torch.from_numpy(g["col"]))
    assert torch.equal(shape.cpu(), torch.from_numpy(g["shape"]))
    assert torch.equal(start.cpu(), torch.from_numpy(g["start"]))
    assert start.cpu()[2].tolist() == [0, 24]                    # the reference's restart
    col2, shape2, start2 = feature_maps_format(nested)
    assert torch.equal(col2, col) and torch.equal(shape2, shape)
    counts = (shape2[..., 0] * shape2[..., 1]).flatten().cpu()
    assert start2.cpu().flatten().tolist() == (counts.cumsum(0) - counts).tolist()
    for tables in ([col, shape, start], [col2, shape2, start2]):   # the inverse reads sizes only
        back = feature_maps_format(tables, inverse=True)
        assert len(back) == len(nested)
        for a, b in zip(back, nested):
            assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_border_values_are_excluded():
    """loc exactly 0 or 1 is masked (exclusive test, …_cuda.cu:168-171)."""
    from simpb_b200 import cabi
    d = small_case(6, bs=1, A=2, P=2, K=1, sizes=((4, 4),), C=32, G=4, lo=0.2, hi=0.8)
    loc = d["sampling_location"]
    loc[0, 0, 0, 0, 0] = 0.0
    loc[0, 0, 1, 0, 1] = 1.0
    loc[0, 1, 0, 0, 0] = float(np.nextafter(np.float32(1), np.float32(0)))   # still valid
    check_case(d)
    valid, _ = cabi.debug_indices(d["spatial_shape"].int().cuda(), d["scale_start_index"].int().cuda(),
                                  loc.cuda())
    assert valid.flatten().tolist() == [0, 0, 1, 1]


def test_output_buffer_needs_no_zero_fill():
    from simpb_b200 import cabi
    d = small_case(7, bs=2, A=5, P=3, K=2, sizes=SIZES3, C=64, G=8)
    g = dev(d)
    a = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    junk = torch.full_like(a, float("nan"))
    b = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], out=junk)
    assert rel_err(b, a) < 1e-6 and torch.isfinite(b).all()


def test_run_to_run_reproducibility():
    """The forward merges duplicate rows and sums in a fixed order, and the two small gradients are
    produced without atomics: all three are bitwise reproducible (the reference's float atomics
    are not)."""
    from simpb_b200 import cabi, synthetic
    g = dev(synthetic.op_inputs_uniform(bs=1, A=300, seed=2))
    a = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    for _ in range(3):
        assert torch.equal(cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"]), a)
    _, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    _, gl2, gw2 = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    assert torch.equal(gl, gl2) and torch.equal(gw, gw2)


# ------------------------------------------------------------------ size-independent properties
def test_full_size_properties_bs8():
    """Training shape (bs=8, A=1220): linearity in features and weights, batch independence."""
    from simpb_b200 import cabi, synthetic
    d = synthetic.rig_op_inputs(bs=8, A=1220, seed=9)
    g = dev(d)
    f = lambda feat, w: cabi.forward(feat, g["shape"], g["start"], g["loc"], w)  # noqa: E731
    base = f(g["feat"], g["w"])
    assert_close(f(g["feat"] * 2, g["w"]), (2 * base), 1e-6, "linear in features")
    assert_close(f(g["feat"], g["w"] * 0.5), (0.5 * base), 1e-6, "linear in weights")
    other = torch.randn_like(g["feat"])
    assert_close(f(g["feat"] + other, g["w"]), base + f(other, g["w"]), 2e-6, "additive in features")
    # batch items are independent: item 3 alone reproduces row 3 bit for bit
    one = cabi.forward(g["feat"][3:4].contiguous(), g["shape"], g["start"],
                       g["loc"][3:4].contiguous(), g["w"][3:4].contiguous())
    assert rel_err(one[0], base[3]) < 1e-6
    # anchors are independent: reversing anchor order reverses the output
    rev = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"].flip(1).contiguous(),
                       g["w"].flip(1).contiguous())
    assert rel_err(rev.flip(1), base) < 1e-6
    # adjoint identity <out, go> == <w, grad_w> (out is linear in w)
    _, _, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    lhs = (base.double() * g["go"].double()).sum()
    rhs = (g["w"].double() * gw.double()).sum()
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


# ------------------------------------------------------------------ python surface
def test_autograd_function_and_torch_extension():
    from simpb_b200 import deformable_aggregation_function
    from simpb_b200.ops import deformable_aggregation_ext as ext
    d = small_case(13, bs=2, A=6, P=4, K=3, sizes=SIZES3, C=64, G=8)
    g = dev(d)
    feat = g["feat"].clone().requires_grad_()
    loc = g["loc"].clone().requires_grad_()
    w = g["w"].clone().requires_grad_()
    # int64 tables straight from feature_maps_format are accepted, as in the reference
    out = deformable_aggregation_function(feat, d["spatial_shape"].cuda(), d["scale_start_index"].cuda(),
                                          loc, w)
    out.backward(g["go"])
    rgf, rgl, rgw = oracle.backward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                                    d["sampling_location"], d["weights"], d["grad_output"])
    assert_close(feat.grad, rgf, RTOL_F32, "autograd grad_feat")
    assert_close(loc.grad, rgl, RTOL_F32, "autograd grad_loc")
    assert_close(w.grad, rgw, RTOL_F32, "autograd grad_weights")
    # the pybind module with the reference's two entry points
    o2 = ext.deformable_aggregation_forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    assert rel_err(o2, out.detach()) < 1e-6
    gf = torch.zeros_like(g["feat"]); gl = torch.zeros_like(g["loc"]); gw = torch.zeros_like(g["w"])
    ext.deformable_aggregation_backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"],
                                        gf, gl, gw)
    assert_close(gf, rgf, RTOL_F32, "ext grad_feat")
    assert torch.equal(gl, loc.grad) and torch.equal(gw, w.grad)
    with pytest.raises(RuntimeError):
        ext.deformable_aggregation_forward(g["feat"], g["shape"].long(), g["start"], g["loc"], g["w"])


def test_frozen_features_skip_the_scatter_and_retained_graphs_work():
    from simpb_b200 import cabi, deformable_aggregation_function
    d = small_case(16, bs=2, A=12, P=13, K=6, sizes=SIZES3, C=256, G=8)
    g = dev(d)
    rgf, rgl, rgw = oracle.backward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                                    d["sampling_location"], d["weights"], d["grad_output"])
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"], need_feat=False)
    assert gf is None
    assert_close(gl, rgl, RTOL_F32, "grad_loc without feature gradient")
    assert_close(gw, rgw, RTOL_F32, "grad_weights without feature gradient")
    # autograd: features frozen
    loc = g["loc"].clone().requires_grad_(); w = g["w"].clone().requires_grad_()
    deformable_aggregation_function(g["feat"], g["shape"], g["start"], loc, w).backward(g["go"])
    assert_close(loc.grad, rgl, RTOL_F32, "autograd grad_loc (frozen features)")
    # autograd: two backward passes through one retained graph (the pre-zeroed buffer is used once)
    feat = g["feat"].clone().requires_grad_()
    out = deformable_aggregation_function(feat, g["shape"], g["start"], g["loc"], g["w"])
    out.backward(g["go"], retain_graph=True)
    first = feat.grad.clone()
    feat.grad = None
    out.backward(g["go"])
    assert_close(first, rgf, RTOL_F32, "grad_feat, first pass")
    assert_close(feat.grad, rgf, RTOL_F32, "grad_feat, second pass")


def test_runs_on_a_side_stream():
    from simpb_b200 import cabi
    d = small_case(14, bs=1, A=50, P=13, K=6, sizes=SIZES3, C=256, G=8)
    g = dev(d)
    ref = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    s.synchronize()
    assert rel_err(out, ref) < 1e-6


def test_host_buffer_entry_point():
    from simpb_b200 import cabi
    d = small_case(15, bs=2, A=9, P=13, K=6, sizes=SIZES3, C=256, G=8)
    dims = cabi.Dims(2, 6, d["num_feat"], 256, 3, 9, 13, 8)
    hf = cabi.HostForward(dims)
    pin = lambda t: t.contiguous().pin_memory()  # noqa: E731
    h_out = torch.empty(2, 9, 256).pin_memory()
    hf(pin(d["mc_ms_feat"]), pin(d["spatial_shape"].int()), pin(d["scale_start_index"].int()),
       pin(d["sampling_location"]), pin(d["weights"]), h_out)
    ref = oracle.forward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                         d["sampling_location"], d["weights"])
    assert_close(h_out, ref, RTOL_F32, "dfa_forward_host")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_host_buffer_pull_mode_moves_exactly_what_the_forward_reads(dtype, monkeypatch):
    """dfa_forward_host with pinned (mapped) host buffers pulls only the feature rows and weight lines the
    forward reads.  The result must be bit-identical to the whole-copy path (pageable buffers, and pinned
    ones with DFA_HOST_PULL=0) — also when the workspace still holds OTHER data from a previous call — and
    the number of rows pulled must equal the oracle's count of distinct referenced rows."""
    from simpb_b200 import cabi, synthetic
    cases = [synthetic.rig_op_inputs(bs=2, A=300, seed=71),
             small_case(72, bs=2, A=9, P=13, K=6, sizes=SIZES3, C=256, G=8),
             small_case(73, bs=1, A=5, P=3, K=3, sizes=SIZES3, C=64, G=8)]
    for ci, d in enumerate(cases):
        bs, A, P, K = d["sampling_location"].shape[:4]
        L, G = d["weights"].shape[4:6]
        C = d["mc_ms_feat"].shape[2]
        dims = cabi.Dims(bs, K, d["num_feat"], C, L, A, P, G)
        hf = cabi.HostForward(dims, dtype)
        hf.workspace.fill_(0xFF)                                  # stale garbage (NaN patterns) everywhere
        host = [d["mc_ms_feat"].to(dtype).contiguous(), d["spatial_shape"].int().contiguous(),
                d["scale_start_index"].int().contiguous(), d["sampling_location"].contiguous(),
                d["weights"].contiguous()]
        pinned = [t.pin_memory() for t in host]
        out_pull = hf(*pinned, torch.empty(bs, A, C).pin_memory()).clone()
        h2d, rows, wbytes = hf.stats()
        U = oracle.distinct_rows(d["spatial_shape"], d["scale_start_index"], d["sampling_location"], d["num_feat"])
        valid = ((d["sampling_location"] > 0) & (d["sampling_location"] < 1)).all(-1).sum().item()
        assert rows == U and wbytes == valid * L * G * 4
        small = sum(t.numel() * t.element_size() for t in host[1:4])
        assert h2d == rows * C * host[0].element_size() + wbytes + small
        if ci == 0:      # camera-rig inputs: most of the table is never referenced
            assert h2d < 0.5 * sum(t.numel() * t.element_size() for t in host)
        hf.workspace.fill_(0xFF)
        out_copy = hf(*host, torch.empty(bs, A, C)).clone()       # pageable: whole copies
        assert hf.stats()[1] == bs * d["num_feat"]
        monkeypatch.setenv("DFA_HOST_PULL", "0")
        out_copy2 = hf(*pinned, torch.empty(bs, A, C).pin_memory()).clone()
        monkeypatch.delenv("DFA_HOST_PULL")
        assert torch.equal(out_pull, out_copy) and torch.equal(out_pull, out_copy2)
        ref = oracle.forward(host[0].float(), d["spatial_shape"], d["scale_start_index"],
                             d["sampling_location"], d["weights"])
        assert_close(out_pull, ref, RTOL_F32, "dfa_forward_host pull mode")


# ------------------------------------------------------------------ flatten / key points
def test_flatten_maps_matches_reference_layout():
    from simpb_b200 import feature_maps_format
    g = load_golden("flatten_small")
    maps = [torch.from_numpy(g["map%d" % l]).cuda() for l in range(g["sizes"].shape[0])]
    col, shape, start = feature_maps_format(maps)
    assert shape.dtype == torch.int64 and start.dtype == torch.int64
    np.testing.assert_array_equal(col.cpu().numpy(), g["col"])        # data movement: bit-exact
    np.testing.assert_array_equal(shape.cpu().numpy(), g["shape"])
    np.testing.assert_array_equal(start.cpu().numpy(), g["start"])
    back = feature_maps_format([col, shape, start], inverse=True)
    assert len(back) == int(g["inverse_n_groups"]) and len(back[0]) == int(g["inverse_n_levels"])
    for l, m in enumerate(back[0]):
        np.testing.assert_array_equal(m.cpu().numpy(), g["inv%d" % l])
    colh, _, _ = feature_maps_format(maps, dtype=torch.bfloat16)
    assert torch.equal(colh, col.bfloat16())


def test_flatten_maps_r50_shape():
    from simpb_b200 import feature_maps_format, synthetic
    gen = torch.Generator().manual_seed(0)
    maps = [torch.randn(1, 6, 256, h, w, generator=gen) for h, w in synthetic.R50_LEVELS]
    col, shape, start = feature_maps_format([m.cuda() for m in maps])
    rcol, rshape, rstart = module_ref.flatten_feature_maps(maps)
    assert torch.equal(col.cpu(), rcol) and torch.equal(shape.cpu(), rshape)
    assert torch.equal(start.cpu(), rstart)


def test_keypoints_project_vs_oracle():
    from simpb_b200 import cabi, synthetic
    d = synthetic.module_inputs_rig(bs=2, A=900, seed=5, feat=False)
    gen = torch.Generator().manual_seed(6)
    logits = torch.randn(2, 900, 18, generator=gen)
    fix = torch.tensor(synthetic.FIX_SCALE)
    loc, kp = cabi.keypoints_project(d["anchor"].cuda(), fix.cuda(), logits.cuda(),
                                     d["projection_mat"].cuda(), d["image_wh"].cuda(),
                                     want_key_points=True)
    rkp = module_ref.key_points(d["anchor"].double(), fix.double(), logits.double())
    ruv = module_ref.project_points(rkp, d["projection_mat"].double(), d["image_wh"].double())
    ruv = ruv.permute(0, 2, 3, 1, 4)
    assert_close(kp, rkp, 1e-6, "key points")
    # compare where the projection is well conditioned (in front of the camera, O(1) coords)
    sel = (ruv.abs() < 3).all(-1)
    err = (loc.cpu().double() - ruv)[sel].abs().max()
    assert err < 2e-5, err
    # validity masks: identical except for points within rounding distance of a border
    mine = module_ref.op_valid_mask(loc.cpu())
    ref = module_ref.op_valid_mask(ruv.float())
    flips = (mine != ref)
    if flips.any():
        dist = torch.minimum(ruv.abs(), (ruv - 1).abs()).min(-1).values
        assert (dist[flips] < 1e-5).all()
    assert flips.float().mean() < 1e-4


# ------------------------------------------------------------------ the unmodified reference op
def _ref_ext():
    from oracle import build_ref
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref/ not built (needs /root/reference at build time)")
    return build_ref.load()


def test_against_reference_cuda_op_r50():
    from simpb_b200 import cabi, synthetic
    ref_ext = _ref_ext()
    for d in (synthetic.op_inputs_uniform(bs=1, seed=0), synthetic.rig_op_inputs(bs=2, seed=1)):
        g = dev(d)
        out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
        rout = ref_ext.deformable_aggregation_forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
        o64 = oracle.forward(d["mc_ms_feat"], d["spatial_shape"], d["scale_start_index"],
                             d["sampling_location"], d["weights"])
        assert_close(out, rout, RTOL_F32, "forward vs reference op")
        assert_close(rout, o64, RTOL_F32, "reference op vs oracle")
        gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
        rgf = torch.zeros_like(g["feat"]); rgl = torch.zeros_like(g["loc"]); rgw = torch.zeros_like(g["w"])
        ref_ext.deformable_aggregation_backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"],
                                                g["go"], rgf, rgl, rgw)
        assert_close(gf, rgf, RTOL_F32, "grad_feat vs reference op")
        assert_close(gw, rgw, RTOL_F32, "grad_weights vs reference op")
        assert_close(gl, rgl, 2 * RTOL_F32, "grad_loc vs reference op")


def _vs_reference_op(g, what, loc_tol=2 * RTOL_F32):
    from simpb_b200 import cabi
    ref_ext = _ref_ext()
    out = cabi.forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    rout = ref_ext.deformable_aggregation_forward(g["feat"], g["shape"], g["start"], g["loc"], g["w"])
    assert_close(out, rout, RTOL_F32, what + ": forward vs reference op")
    del out, rout
    gf, gl, gw = cabi.backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"])
    rgf = torch.zeros_like(g["feat"]); rgl = torch.zeros_like(g["loc"]); rgw = torch.zeros_like(g["w"])
    ref_ext.deformable_aggregation_backward(g["feat"], g["shape"], g["start"], g["loc"], g["w"], g["go"],
                                            rgf, rgl, rgw)
    assert_close(gf, rgf, RTOL_F32, what + ": grad_feat vs reference op")
    assert_close(gw, rgw, RTOL_F32, what + ": grad_weights vs reference op")
    assert_close(gl, rgl, loc_tol, what + ": grad_loc vs reference op")


def test_training_shape_bs8_all_gradients_vs_reference_cuda_op():
    """bs=8 x 900 anchors (the window-merging forward is the default there) and the backward's three
    gradients, every element, against the unmodified reference binary on the same inputs."""
    from simpb_b200 import synthetic
    _vs_reference_op(dev(synthetic.rig_op_inputs(bs=8, seed=61)), "R50 bs=8")


@pytest.mark.parametrize("corner", ["A3600_P32", "R101_bs8"])
def test_sweep_corners_vs_reference_cuda_op(corner):
    """The extreme points of BASELINE.json configs #4 / #5 (tools/op_sweep.py times them): 3,600
    anchors x 32 key points at bs=1, and R101 1408x512 maps at bs=8."""
    from simpb_b200 import synthetic
    if corner == "A3600_P32":
        d = synthetic.rig_op_inputs(bs=1, A=3600, P=32, seed=62)
    else:
        d = synthetic.rig_op_inputs(bs=8, levels=synthetic.R101_LEVELS, seed=63, feat=False)
        gen = torch.Generator().manual_seed(63)
        d["mc_ms_feat"] = torch.randn(8, d["num_feat"], 256, generator=gen)
    _vs_reference_op(dev(d), corner)
    torch.cuda.empty_cache()


def test_reference_binary_pins_indices_and_masks():
    """Bit-exact index/mask pin against the reference BINARY.  Every (a,p,k) sample gets a private
    channel (G == C, one-hot weights and grad_output), so the non-zero pattern of the reference
    op's grad_mc_ms_feat[b, :, channel] is exactly the set of rows that sample touched.  Uses
    rounding-tie locations, where a two-rounding evaluation of loc*size-0.5 would differ."""
    from simpb_b200 import cabi, synthetic
    ref_ext = _ref_ext()
    sizes = synthetic.R50_LEVELS
    K, L, P, A = 6, 4, 4, 10
    C = A * P * K                                   # 240 channels, one per sample
    shape, start, num_feat = synthetic.level_tables(sizes, K)
    v = tie_locations(sizes, seed=3)
    per = A * P * K * 2
    nb = min(len(v) // per, 24)
    loc = torch.from_numpy(v[:nb * per].copy()).reshape(nb, A, P, K, 2)
    w = torch.zeros(nb, A, P, K, L, C)
    go = torch.zeros(nb, A, C)
    for a in range(A):
        for p in range(P):
            for k in range(K):
                ch = (a * P + p) * K + k
                w[:, a, p, k, :, ch] = 1.0
                go[:, a, ch] = 1.0
    feat = torch.ones(nb, num_feat, C)
    sh, st = shape.int().cuda(), start.int().cuda()
    rgf = torch.zeros(nb, num_feat, C, device="cuda")
    rgl = torch.zeros_like(loc).cuda(); rgw = torch.zeros_like(w).cuda()
    ref_ext.deformable_aggregation_backward(feat.cuda(), sh, st, loc.cuda(), w.cuda(), go.cuda(),
                                            rgf, rgl, rgw)
    touched_ref = (rgf != 0).cpu().numpy()          # [nb, num_feat, C]
    valid, rows = cabi.debug_indices(sh, st, loc.cuda())
    valid, rows = valid.cpu().numpy(), rows.cpu().numpy()
    _, ov, orows = oracle.forward(np.zeros((nb, num_feat, 1), np.float32), shape, start, loc,
                                  np.zeros((nb, A, P, K, L, 1), np.float32), side_channel=True)
    np.testing.assert_array_equal(valid, ov)
    np.testing.assert_array_equal(rows, orows)
    mine = np.zeros_like(touched_ref)
    b_idx = np.arange(nb)[:, None, None, None, None, None]
    ch = ((np.arange(A)[:, None, None] * P + np.arange(P)[None, :, None]) * K
          + np.arange(K)[None, None, :])[None, :, :, :, None, None]
    sel = rows >= 0
    mine[np.broadcast_to(b_idx, rows.shape)[sel], rows[sel], np.broadcast_to(ch, rows.shape)[sel]] = True
    # a corner with an exactly-zero bilinear weight leaves no trace in the reference gradient:
    # the reference pattern must be a subset of ours, and equal wherever our weight is non-zero
    assert not (touched_ref & ~mine).any(), "reference touched a row our geometry does not"
    gf, _, _ = cabi.backward(feat.cuda(), sh, st, loc.cuda(), w.cuda(), go.cuda())
    np.testing.assert_array_equal((gf != 0).cpu().numpy(), touched_ref)
    # masks: a sample is valid iff the reference produced any gradient for its channel … unless
    # every corner weight is zero, which cannot happen for a valid sample (weights sum to 1 over
    # in-bounds + out-of-bounds corners; at least one in-bounds corner has weight > 0)
    ref_valid = touched_ref.any(axis=1).reshape(nb, A, P, K)
    np.testing.assert_array_equal(ref_valid, valid.astype(bool))
