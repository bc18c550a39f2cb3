"""The SimPB+ R50 decoder frame (simpb_b200/decoder.py): a structural restatement for the frames/sec
figure.  Checked here: the released operation order, the 2-D query allocation (static slots == the
data-dependent layout on every valid slot), and — on the GPU — that frames run, the temporal branch becomes
live on the second frame, and the static-shape frame agrees with the eager one."""
import collections

import pytest
import torch


def test_operation_order_is_the_released_one():
    from simpb_b200 import decoder
    c = collections.Counter(decoder.OPERATION_ORDER)
    # SURVEY.md Appendix B / config :58-72: 50 ops
    assert len(decoder.OPERATION_ORDER) == 50
    assert c == dict(deformable=3, qg_cross_attn=3, gnn=3, temp_gnn=5, aggregation=3, qg_self_attn=3, ffn=6,
                     norm=12, refine3d=6, refine2d=3, allocation=3)
    assert decoder.OPERATION_ORDER[:9] == decoder.SINGLE_2D and decoder.OPERATION_ORDER[-1] == "refine3d"


def test_static_allocation_matches_the_data_dependent_layout():
    from simpb_b200 import decoder, synthetic
    alloc = decoder.DynamicQueryAllocation()
    proj, wh = synthetic.camera_rig(1)
    a = synthetic.rig_anchors(torch.Generator().manual_seed(3), 1, 900)
    ref, dep, trans, center, groups, none = alloc(a, proj, wh)
    assert none is None and trans.shape == (1, groups[-1][1], 900)
    assert (trans.sum(1) >= 0).all() and trans.sum(-1).eq(1).all()       # one anchor per 2-D query
    ref2, dep2, trans2, center2, groups2, valid = alloc(a, proj, (704.0, 256.0), cap=320)
    for (s, e), (s2, e2) in zip(groups, groups2):
        n = e - s
        assert valid[0, s2:s2 + n].all() and not valid[0, s2 + n:e2].any()
        for x, y in ((ref, ref2), (dep, dep2), (trans, trans2), (center, center2)):
            assert torch.equal(x[0, s:e], y[0, s2:s2 + n])
        assert trans2[0, s2 + n:e2].abs().sum() == 0                      # unused slots feed nothing


@pytest.mark.gpu
def test_frames_run_and_static_frame_matches_eager():
    from simpb_b200 import decoder, synthetic
    dev = "cuda"
    proj, wh = synthetic.camera_rig(1)
    outs = {}
    for cap in (None, 320):
        m = decoder.SimPBFrame(seed=1, static_queries=cap).to(dev).eval()
        gen = torch.Generator().manual_seed(5)
        T = torch.eye(4)[None].clone()
        T[0, 1, 3] = -2.5
        metas = dict(projection_mat=proj.to(dev), image_wh=wh.to(dev), img_wh=(704.0, 256.0),
                     T_temp2cur=T.to(dev), dt=torch.full((1,), 0.5, device=dev))
        res = []
        with torch.no_grad():
            for i in range(3):
                img = torch.randn(1, 6, 3, 256, 704, generator=gen).to(dev)
                anchor, cls, qt = m(img, metas)
                assert anchor.shape == (1, 900, 11) and cls.shape == (1, 900, 10) and qt.shape == (1, 900, 2)
                assert torch.isfinite(anchor).all() and torch.isfinite(cls).all()
                assert m.head.cached_anchor.shape == (1, 600, 11)        # temporal instances for the next frame
                res.append((anchor, cls))
        outs[cap] = res
    # same weights (seed), same inputs: padding the 2-D queries to static slots must not change the frame.
    # Frame 0 has no temporal instances and no confidence-ranked selection, so it is compared tightly; from
    # frame 1 on the bank keeps the top-k by confidence, and with random weights the scores are near-ties
    # (any rounding difference reorders them), so later frames are compared on order-free statistics.
    (a0, c0), (a1, c1) = outs[None][0], outs[320][0]
    da = float((a0 - a1).abs().max() / a0.abs().max())
    dc = float((c0 - c1).abs().max() / c0.abs().max())
    print("frame 0 static-vs-eager rel diff: anchor %.2e cls %.2e" % (da, dc))
    assert da < 2e-3 and dc < 2e-3
    for (a0, c0), (a1, c1) in zip(outs[None][1:], outs[320][1:]):
        assert abs(float(c0.mean() - c1.mean())) < 2e-2 * float(c0.abs().mean())
        assert abs(float(a0[..., :3].abs().mean() - a1[..., :3].abs().mean())) < 5e-2 * float(a0[..., :3].abs().mean())


@pytest.mark.gpu
def test_folded_batchnorm_is_the_same_backbone():
    from simpb_b200 import decoder
    torch.manual_seed(0)
    m = decoder.SimPBFrame(seed=2).cuda().eval()
    for mod in m.modules():          # non-trivial running statistics
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    img = torch.randn(1, 6, 3, 256, 704, device="cuda")
    with torch.no_grad():
        ref = m.extract_feat(img)[0]
        out = m.fold_batchnorm().extract_feat(img)[0]
    assert not any(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-2      # fp16 autocast on both sides


@pytest.mark.gpu
def test_bfloat16_table_from_the_neck():
    """the frame with the pyramid flattened into a bfloat16 table — both gathers (fused DFA
    forward, MSDA on the unprojected table) read it natively.  First frame against the fp32-table frame."""
    from simpb_b200 import decoder, synthetic
    proj, wh = synthetic.camera_rig(1)
    metas = dict(projection_mat=proj.cuda(), image_wh=wh.cuda(), img_wh=(704.0, 256.0))
    img = torch.randn(1, 6, 3, 256, 704, generator=torch.Generator().manual_seed(7)).cuda()
    outs = []
    for dt in (None, torch.bfloat16):
        m = decoder.SimPBFrame(seed=3, static_queries=320, table_dtype=dt).cuda().eval()
        with torch.no_grad():
            fm = m.extract_feat(img)
            assert fm[0].dtype == (torch.float32 if dt is None else torch.bfloat16)
            outs.append(m.head(fm, metas))
    (a0, c0, _), (a1, c1, _) = outs
    assert float((a0 - a1).abs().max() / a0.abs().max()) < 2e-2
    assert float((c0 - c1).abs().max() / c0.abs().max()) < 2e-2
