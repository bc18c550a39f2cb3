"""Multi-scale deformable attention (the 2-D query branch's gather).  mmcv-full 1.7.1 is not
vendored, so there are no vectors from mmcv itself: the C oracle (mmcv's CUDA kernel restated) and
the grid_sample formulation (mmcv's CPU fallback restated) are checked against each other and against
a fixture produced by an independent third-party implementation of the same published function
(HuggingFace transformers' MultiScaleDeformableAttention, tests/golden/msda_hf.npz), then the CUDA
kernels against the oracle and the fixture."""
import numpy as np
import pytest
import torch

from oracle import msda_ref
from helpers import RTOL_BF16, RTOL_F32, assert_close, load_golden

SIZES3 = ((8, 12), (4, 6), (2, 3))


def make_case(seed, bs, Q, M, D, sizes, P, lo=-0.15, hi=1.15):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.tensor(sizes, dtype=torch.int64)
    counts = shapes[:, 0] * shapes[:, 1]
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])
    S, L = int(counts.sum()), len(sizes)
    return dict(value=torch.randn(bs, S, M, D, generator=g), shapes=shapes, start=start,
                loc=torch.rand(bs, Q, M, L, P, 2, generator=g) * (hi - lo) + lo,
                w=torch.rand(bs, Q, M, L, P, generator=g).reshape(bs, Q, M, -1).softmax(-1)
                .reshape(bs, Q, M, L, P),
                go=torch.randn(bs, Q, M * D, generator=g))


# ------------------------------------------------------------------ CPU: the two oracles agree
@pytest.mark.parametrize("cfg", [dict(bs=2, Q=7, M=4, D=8, sizes=SIZES3, P=4),
                                 dict(bs=1, Q=5, M=8, D=32, sizes=SIZES3, P=4),
                                 dict(bs=1, Q=3, M=2, D=5, sizes=((5, 7),), P=3)])
def test_c_oracle_matches_grid_sample_formulation(cfg):
    d = make_case(1, **cfg)
    out = msda_ref.forward(d["value"], d["shapes"], d["start"], d["loc"], d["w"])
    v, l, w = (d[k].double().requires_grad_() for k in ("value", "loc", "w"))
    ref = msda_ref.msda_grid_sample(v, d["shapes"], l, w)
    assert_close(out, ref, 2e-6, "forward")
    ref.backward(d["go"].double())
    gv, gl, gw = msda_ref.backward(d["value"], d["shapes"], d["start"], d["loc"], d["w"], d["go"])
    assert_close(gv, v.grad, 2e-6, "grad_value")
    assert_close(gw, w.grad, 2e-6, "grad_attn_weight")
    assert_close(gl, l.grad, 1e-5, "grad_sampling_loc")


def hf_case():
    g = load_golden("msda_hf")
    shapes = torch.from_numpy(g["sizes"])
    counts = shapes[:, 0] * shapes[:, 1]
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])
    d = dict(value=torch.from_numpy(g["value"]), shapes=shapes, start=start, loc=torch.from_numpy(g["loc"]),
             w=torch.from_numpy(g["w"]), go=torch.from_numpy(g["go"]))
    return g, d


def test_oracles_match_the_third_party_fixture():
    """Forward and the three gradients of both restatements against HuggingFace transformers'
    implementation (fp64 autograd), stored in tests/golden/msda_hf.npz; and, when transformers is
    importable, against a live call on the same inputs."""
    g, d = hf_case()
    out = msda_ref.forward(d["value"], d["shapes"], d["start"], d["loc"], d["w"])
    assert_close(out, g["out"], 2e-6, "C oracle forward")
    gv, gl, gw = msda_ref.backward(d["value"], d["shapes"], d["start"], d["loc"], d["w"], d["go"])
    assert_close(gv, g["grad_value"], 2e-6, "C oracle grad_value")
    assert_close(gw, g["grad_w"], 2e-6, "C oracle grad_attn_weight")
    assert_close(gl, g["grad_loc"], 1e-5, "C oracle grad_sampling_loc")
    v, l, w = (d[k].double().requires_grad_() for k in ("value", "loc", "w"))
    ref = msda_ref.msda_grid_sample(v, d["shapes"], l, w)
    assert_close(ref, g["out"], 1e-12, "grid_sample restatement forward")
    ref.backward(d["go"].double())
    assert_close(v.grad, g["grad_value"], 1e-12, "grid_sample restatement grad_value")
    assert_close(l.grad, g["grad_loc"], 1e-10, "grid_sample restatement grad_sampling_loc")
    try:
        from transformers.models.deformable_detr.modeling_deformable_detr import MultiScaleDeformableAttention
    except Exception:
        return
    live = MultiScaleDeformableAttention()(d["value"].double(), d["shapes"], [tuple(x) for x in g["sizes"].tolist()],
                                           d["start"], d["loc"].double(), d["w"].double(), 64)
    assert_close(live, g["out"], 1e-12, "live transformers call")


def test_taps_outside_the_border_band_contribute_nothing():
    """h_im > -1 && w_im > -1 && h_im < H && w_im < W: one pixel beyond the map is the limit."""
    d = make_case(2, bs=1, Q=2, M=1, D=4, sizes=((4, 4),), P=2)
    d["loc"][0, 0, 0, 0, 0] = torch.tensor([-0.13, 0.5])       # w_im = -1.02: not taken
    d["loc"][0, 0, 0, 0, 1] = torch.tensor([-0.12, 0.5])       # w_im = -0.98: taken, weight on column 0 only
    d["loc"][0, 1, 0, 0, 0] = torch.tensor([0.5, 1.13])        # h_im = 4.02: not taken
    d["loc"][0, 1, 0, 0, 1] = torch.tensor([0.5, 1.12])        # h_im = 3.98: taken
    out = msda_ref.forward(d["value"], d["shapes"], d["start"], d["loc"], d["w"])
    w0 = d["w"].clone(); w0[0, :, 0, 0, 0] = 0                  # dropping the "not taken" taps changes nothing
    assert_close(msda_ref.forward(d["value"], d["shapes"], d["start"], d["loc"], w0), out, 0, "untaken taps")
    w1 = d["w"].clone(); w1[0, :, 0, 0, 1] = 0
    assert np.abs(msda_ref.forward(d["value"], d["shapes"], d["start"], d["loc"], w1) - out).max() > 1e-4


def test_module_interface_matches_mmcv_msda_parameters():
    from simpb_b200 import msda
    m = msda.QueryGroupMultiScaleDeformableAttention(embed_dims=256, num_heads=8, num_levels=4,
                                                     num_points=4, num_cams=6)
    sd = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert sd == {"sampling_offsets.weight": (256, 256), "sampling_offsets.bias": (256,),
                  "attention_weights.weight": (128, 256), "attention_weights.bias": (128,),
                  "value_proj.weight": (256, 256), "value_proj.bias": (256,),
                  "output_proj.weight": (256, 256), "output_proj.bias": (256,)}
    assert float(m.sampling_offsets.weight.abs().max()) == 0.0
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1.0, 2.0, 3.0, 4.0]))    # head 0 looks along +x


# ------------------------------------------------------------------ GPU: kernels vs the oracle
def gpu(d, dtype=torch.float32):
    return dict(value=d["value"].cuda().to(dtype), shapes=d["shapes"].int().cuda(),
                start=d["start"].int().cuda(), loc=d["loc"].cuda(), w=d["w"].cuda(), go=d["go"].cuda())


def check(d, dtype=torch.float32):
    from simpb_b200 import cabi
    g = gpu(d, dtype)
    vref = g["value"].float().cpu()
    out = cabi.msda_forward(g["value"], g["shapes"], g["start"], g["loc"], g["w"])
    assert_close(out, msda_ref.forward(vref, d["shapes"], d["start"], d["loc"], d["w"]), RTOL_F32, "forward")
    gv, gl, gw = cabi.msda_backward(g["value"], g["shapes"], g["start"], g["loc"], g["w"], g["go"])
    rgv, rgl, rgw = msda_ref.backward(vref, d["shapes"], d["start"], d["loc"], d["w"], d["go"])
    assert_close(gv, rgv, RTOL_F32, "grad_value")
    assert_close(gl, rgl, RTOL_F32, "grad_sampling_loc")
    assert_close(gw, rgw, RTOL_F32, "grad_attn_weight")
    _, gl2, gw2 = cabi.msda_backward(g["value"], g["shapes"], g["start"], g["loc"], g["w"], g["go"],
                                     need_value=False)
    assert torch.equal(gl, gl2) and torch.equal(gw, gw2)          # no atomics: bitwise reproducible


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [
    dict(bs=2, Q=7, M=4, D=8, sizes=SIZES3, P=4),        # 2 vector lanes per head
    dict(bs=1, Q=33, M=8, D=32, sizes=SIZES3, P=4),      # SimPB head geometry (8 lanes)
    dict(bs=1, Q=9, M=8, D=16, sizes=SIZES3, P=3),       # 4 lanes, L*P not a multiple of the batch
    dict(bs=2, Q=5, M=3, D=4, sizes=SIZES3, P=2),        # 1 lane per head, odd head count: no TMA
    dict(bs=1, Q=4, M=2, D=5, sizes=((5, 7),), P=3),     # generic kernels
    dict(bs=1, Q=3, M=16, D=32, sizes=SIZES3, P=8),      # 16 heads, 24 taps per head
])
def test_kernels_vs_oracle_small_shapes(cfg):
    check(make_case(3, **cfg))


@pytest.mark.gpu
def test_kernels_vs_third_party_fixture():
    """CUDA forward / backward against the HuggingFace-transformers fixture."""
    from simpb_b200 import cabi
    g, d = hf_case()
    c = {k: v.cuda() for k, v in d.items()}
    shapes, start = c["shapes"].int(), c["start"].int()
    out = cabi.msda_forward(c["value"], shapes, start, c["loc"], c["w"])
    assert_close(out, g["out"], RTOL_F32, "msda forward")
    gv, gl, gw = cabi.msda_backward(c["value"], shapes, start, c["loc"], c["w"], c["go"])
    assert_close(gv, g["grad_value"], RTOL_F32, "grad_value")
    assert_close(gw, g["grad_w"], RTOL_F32, "grad_attn_weight")
    assert_close(gl, g["grad_loc"], 5 * RTOL_F32, "grad_sampling_loc")


@pytest.mark.gpu
def test_kernels_vs_oracle_simpb_shape():
    """One camera of the R50 704x256 pyramid (14,960 rows x 256), 300 queries, 8 heads x 4 x 4."""
    from simpb_b200 import synthetic
    check(make_case(4, bs=2, Q=300, M=8, D=32, sizes=synthetic.R50_LEVELS, P=4, lo=-0.05, hi=1.05))


@pytest.mark.gpu
def test_grouped_launch_equals_per_group_calls():
    """value [bs,K,S,M,D] + query_table in one launch == the reference's loop over camera groups."""
    from simpb_b200 import cabi
    d = make_case(6, bs=2, Q=30, M=8, D=32, sizes=SIZES3, P=4)
    g = gpu(d)
    K = 3
    gen = torch.Generator().manual_seed(7)
    value = torch.randn(2, K, d["value"].shape[1], 8, 32, generator=gen).cuda()
    groups = [(0, 11), (11, 11), (11, 30)]
    table = torch.tensor([0] * 11 + [2] * 19, dtype=torch.int32).cuda()
    out = cabi.msda_forward(value, g["shapes"], g["start"], g["loc"], g["w"], table)
    gv, gl, gw = cabi.msda_backward(value, g["shapes"], g["start"], g["loc"], g["w"], g["go"], query_table=table)
    for i, (a, b) in enumerate(groups):
        if b == a:
            assert float(gv[:, i].abs().max()) == 0.0
            continue
        sl = lambda t: t[:, a:b].contiguous()  # noqa: E731
        o = cabi.msda_forward(value[:, i].contiguous(), g["shapes"], g["start"], sl(g["loc"]), sl(g["w"]))
        assert torch.equal(out[:, a:b], o)
        v2, l2, w2 = cabi.msda_backward(value[:, i].contiguous(), g["shapes"], g["start"], sl(g["loc"]),
                                        sl(g["w"]), sl(g["go"]))
        assert torch.equal(gl[:, a:b], l2) and torch.equal(gw[:, a:b], w2)
        assert_close(gv[:, i], v2, 1e-6, "grad_value of group %d" % i)


@pytest.mark.gpu
def test_bf16_value_table():
    from simpb_b200 import cabi
    d = make_case(5, bs=1, Q=40, M=8, D=32, sizes=SIZES3, P=4)
    check(d, torch.bfloat16)
    g = gpu(d, torch.bfloat16)
    out = cabi.msda_forward(g["value"], g["shapes"], g["start"], g["loc"], g["w"])
    assert_close(out, msda_ref.forward(d["value"], d["shapes"], d["start"], d["loc"], d["w"]), RTOL_BF16,
                 "bf16 value vs fp32 oracle")


@pytest.mark.gpu
def test_autograd_function_and_query_group_module():
    """The module on camera groups against the same module evaluated with the grid_sample
    formulation (mmcv's CPU fallback) in fp64 — outputs and every gradient."""
    from simpb_b200 import msda, synthetic
    torch.manual_seed(0)
    bs, cams, C = 2, 3, 64
    sizes = SIZES3
    shapes = torch.tensor(sizes)
    counts = shapes[:, 0] * shapes[:, 1]
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])
    S = int(counts.sum())
    groups = [(0, 5), (5, 5), (5, 12)]                    # camera 1 has no query
    m = msda.QueryGroupMultiScaleDeformableAttention(embed_dims=C, num_heads=4, num_levels=3, num_points=4,
                                                     num_cams=cams, query_groups=groups, dropout=0.0,
                                                     batch_first=True, residual_mode="cat")
    with torch.no_grad():                                  # the default init makes every query look alike
        m.sampling_offsets.weight.normal_(0, 0.05)
        m.attention_weights.weight.normal_(0, 0.5)
    g = torch.Generator().manual_seed(1)
    query = torch.randn(bs, 12, C, generator=g)
    value = torch.randn(bs * cams, S, C, generator=g)
    ref_pts = torch.rand(bs, 12, 3, 2, generator=g)
    go = torch.randn(bs, 12, 2 * C, generator=g)

    def run(mod, dev, dt, fn):
        q = query.to(dev, dt).requires_grad_()
        v = value.to(dev, dt).requires_grad_()
        mod = mod.to(dev, dt)
        mod.zero_grad()
        val = mod.value_proj(v).view(bs, cams, S, mod.num_heads, -1)
        off = mod.sampling_offsets(q).view(bs, 12, mod.num_heads, 3, 4, 2)
        w = mod.attention_weights(q).view(bs, 12, mod.num_heads, 12).softmax(-1).view(bs, 12, mod.num_heads, 3, 4)
        loc = mod.sampling_locations(ref_pts.to(dev, dt), off, shapes.to(dev))
        outs = [fn(val[:, i], loc[:, a:b], w[:, a:b]) for i, (a, b) in enumerate(groups) if b > a]
        out = torch.cat([mod.output_proj(torch.cat(outs, 1)), q], -1)
        out.backward(go.to(dev, dt))
        return out, q.grad, v.grad, {n: p.grad.clone() for n, p in mod.named_parameters()}

    ref = run(m, "cpu", torch.float64,
              lambda val, loc, w: msda_ref.msda_grid_sample(val, shapes, loc, w))
    m = m.float().cuda()
    m.zero_grad()
    q = query.cuda().requires_grad_()
    v = value.cuda().requires_grad_()
    out = m(q, value=v, reference_points=ref_pts.cuda(), spatial_shapes=shapes.cuda(),
            level_start_index=start.cuda())
    out.backward(go.cuda())
    assert_close(out, ref[0], RTOL_F32, "module output")
    assert_close(q.grad, ref[1], 2e-5, "grad query")
    assert_close(v.grad, ref[2], 2e-5, "grad value")
    for n, p in m.named_parameters():
        assert_close(p.grad, ref[3][n], 2e-5, "grad " + n)


@pytest.mark.gpu
@pytest.mark.parametrize("C,heads,dtype", [(256, 8, torch.float32), (128, 4, torch.float32),
                                           (256, 8, torch.bfloat16)])
def test_gather_then_project_equals_project_then_gather(C, heads, dtype, monkeypatch):
    """Inference path of the module: whole rows of the UNPROJECTED table gathered per (query, head)
    (dfa_msda_forward_raw), value_proj applied afterwards.  Sampling is linear, so the result equals the
    reference's order (value_proj over the whole table, then sampling) up to fp32 rounding — also for
    taps on and beyond the border (the bias is weighted by the in-map part only), with a bias that is not
    zero, for empty camera groups, and against the kernel-level definition of the two outputs."""
    from simpb_b200 import cabi, msda
    torch.manual_seed(3)
    bs, cams = 2, 3
    sizes = ((8, 12), (4, 6), (2, 3), (1, 2))
    shapes = torch.tensor(sizes)
    counts = shapes[:, 0] * shapes[:, 1]
    start = torch.cat([counts.new_zeros(1), counts.cumsum(0)[:-1]])
    S = int(counts.sum())
    groups = [(0, 9), (9, 9), (9, 21)]                    # camera 1 has no query
    m = msda.QueryGroupMultiScaleDeformableAttention(embed_dims=C, num_heads=heads, num_levels=4, num_points=4,
                                                     num_cams=cams, query_groups=groups, dropout=0.0,
                                                     batch_first=True, residual_mode="cat").cuda().eval()
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.3)
        m.sampling_offsets.bias.mul_(3.0)                  # pushes many taps over the border
        m.attention_weights.weight.normal_(0, 0.5)
        m.value_proj.bias.normal_(0, 0.5)
    g = torch.Generator().manual_seed(4)
    query = torch.randn(bs, 21, C, generator=g).cuda()
    table = torch.randn(bs * cams, S, C, generator=g).cuda().to(dtype)
    ref_pts = torch.rand(bs, 21, 4, 2, generator=g).cuda()
    kw = dict(reference_points=ref_pts, spatial_shapes=shapes.cuda(), level_start_index=start.cuda())
    with torch.no_grad():
        fast = m(query, value=table.float() if dtype == torch.float32 else table, **kw)
        monkeypatch.setenv("SIMPB_B200_MSDA_PROJECT_FIRST", "1")
        slow = m(query, value=table.float(), **kw)
    tol = RTOL_F32 if dtype == torch.float32 else RTOL_BF16
    assert_close(fast, slow, tol, "gather-then-project vs project-then-gather")
    if dtype != torch.float32:
        return
    # the kernel's two outputs against their definition, head by head, through the projected kernel:
    # with W = identity blocks the projected kernel returns the head's slice of the gathered row
    M, D = heads, C // heads
    off = m.sampling_offsets(query).view(bs, 21, M, 4, 4, 2)
    w = m.attention_weights(query).view(bs, 21, M, 16).softmax(-1).view(bs, 21, M, 4, 4)
    loc = m.sampling_locations(ref_pts, off, shapes.cuda()).contiguous()
    qt = m._query_table(21, query.device)
    gth, ssum = cabi.msda_forward_raw(table.view(bs, cams, S, C), shapes.int().cuda(), start.int().cuda(), loc, w, qt)
    for h in range(M):
        lh = loc[:, :, h:h + 1].expand(-1, -1, M, -1, -1, -1).contiguous()     # every head samples like head h
        wh = w[:, :, h:h + 1].expand(-1, -1, M, -1, -1).contiguous()
        ref = cabi.msda_forward(table.view(bs, cams, S, M, D), shapes.int().cuda(), start.int().cuda(), lh, wh, qt)
        assert_close(gth[:, :, h], ref, RTOL_F32, "gathered rows, head %d" % h)
        ones = cabi.msda_forward(torch.ones_like(table).view(bs, cams, S, M, D), shapes.int().cuda(),
                                 start.int().cuda(), lh, wh, qt)
        assert_close(ssum[:, :, h], ones[..., 0], RTOL_F32, "weight sums, head %d" % h)
