"""Pins the oracle (oracle/resnet_fpn_oracle.py) and the product's host-side mirror against the
LIVE, unmodified reference.  Runs only where /root/reference exists (the dev container); the same
facts are re-checked everywhere through the committed golden fixtures (test_oracle_golden.py)."""
import pytest
import torch

from oracle import reference_shim, resnet_fpn_oracle as orc
from tests import helpers

pytestmark = pytest.mark.skipif(not reference_shim.available(), reason="reference tree not present")


@pytest.mark.parametrize("depth", [18, 34, 50, 101])
def test_oracle_bit_identical_to_reference(depth):
    torch.set_num_threads(4)
    bb, neck = reference_shim.build_pair(depth, seed=3)
    x = torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        feats = bb(x)
        outs = neck(feats)
        f, p = orc.resnet_fpn_forward(bb.state_dict(), neck.state_dict(), x, depth)
    for a, b in zip(tuple(feats) + tuple(outs), tuple(f) + tuple(p)):
        assert torch.equal(a, b)


def test_oracle_odd_sizes_and_fpn_mismatch_error():
    """F2: the backbone runs at 1333-derived odd sizes; the FPN raises RuntimeError on them."""
    bb, neck = reference_shim.build_pair(18, seed=0)
    x = torch.randn(1, 3, 100, 167)
    with torch.no_grad():
        feats = bb(x)
        f = orc.resnet_forward(bb.state_dict(), x, 18)
    assert all(torch.equal(a, b) for a, b in zip(feats, f))
    with pytest.raises(RuntimeError):
        neck(feats)
    with pytest.raises(RuntimeError):
        orc.fpn_forward(neck.state_dict(), f, [64, 128, 256, 512], 256, 5)


def test_oracle_variants_match_reference():
    ref_backbone, ref_necks, obj_from_dict = reference_shim.load()
    torch.manual_seed(0)
    bb = obj_from_dict(dict(type="ResNet", depth=50, out_indices=(3,)), parent=ref_backbone)
    bb.init_weights()
    bb.eval()
    x = torch.randn(1, 3, 64, 64)
    with torch.no_grad():
        single = bb(x)
        mine = orc.resnet_forward(bb.state_dict(), x, 50, out_indices=(3,))
    assert isinstance(single, torch.Tensor) and torch.equal(single, mine)
    # RetinaNet-style extra convs, start_level=1
    neck = obj_from_dict(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256,
                              num_outs=5, start_level=1, add_extra_convs=True), parent=ref_necks)
    neck.init_weights()
    neck.eval()
    bb4 = obj_from_dict(dict(type="ResNet", depth=50), parent=ref_backbone)
    bb4.init_weights()
    bb4.eval()
    with torch.no_grad():
        feats = bb4(x)
        ref = neck(feats)
        mine = orc.fpn_forward(neck.state_dict(), feats, [256, 512, 1024, 2048], 256, 5,
                               start_level=1, add_extra_convs=True)
    assert len(ref) == len(mine) == 5
    assert all(torch.equal(a, b) for a, b in zip(ref, mine))


@pytest.mark.parametrize("depth", [18, 50, 101])
def test_product_modules_mirror_reference_state(depth):
    """Same seed -> same state_dict keys, shapes AND values as the reference (construction and
    init_weights consume the RNG identically), for backbone and neck."""
    bb_ref, neck_ref = reference_shim.build_pair(depth, seed=5)
    bb, neck = helpers.build_product_pair(depth, seed=5)
    for ref, mine in ((bb_ref, bb), (neck_ref, neck)):
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a.keys()) == list(b.keys())
        for k in a:
            assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


def test_golden_fixtures_are_reproducible_from_reference():
    for name in helpers.GOLDEN_CASES:
        meta, arrays = helpers.load_golden(name)
        bb, neck = reference_shim.build_pair(meta["depth"], seed=meta["seed"])
        if meta["bnstats"]:
            sd = bb.state_dict()
            orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(1000 + meta["seed"]))
            bb.load_state_dict(sd)
        assert helpers.state_hash(bb.state_dict()) == meta["bb_hash"]
        assert helpers.state_hash(neck.state_dict()) == meta["neck_hash"]


@pytest.mark.parametrize("depth", [18, 50])
def test_gradient_oracle_matches_reference_autograd(depth):
    """The gradient oracle (oracle/grad_oracle.plain_grads) against autograd through the live reference
    modules: same ATen forward and backward ops in the same order -> identical parameter gradients."""
    from oracle import grad_oracle
    torch.set_num_threads(4)
    bb, neck = reference_shim.build_pair(depth, seed=7)
    sd = bb.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(17))
    bb.load_state_dict(sd)
    x = torch.randn(2, 3, 64, 96)
    outs = neck(bb(x))
    g = torch.Generator().manual_seed(1)
    grads = [torch.randn(o.shape, generator=g) for o in outs]
    torch.autograd.backward(list(outs), grads)
    gb, gn, _, _ = grad_oracle.plain_grads(bb.state_dict(), neck.state_dict(), x, depth, grads,
                                           train_from_stage=1)
    ref_b = dict(bb.named_parameters())
    ref_n = dict(neck.named_parameters())
    assert len(gn) == 16
    for k, v in gn.items():
        assert torch.allclose(v, ref_n[k].grad, rtol=1e-5, atol=1e-7), k
    assert gb and all(not k.startswith("layer1") for k in gb)
    for k, v in gb.items():
        assert torch.allclose(v, ref_b[k].grad, rtol=1e-5, atol=1e-7), k


@pytest.mark.parametrize("activation", [None, "relu", "relu6"])
def test_pafpn_oracle_and_mirror_match_reference(activation):
    """Row f3: the PAFPN restatement is bit-identical to the live reference neck, and the product module
    mirrors its state_dict (keys, shapes, values for the same seed)."""
    from torch_detection_b200 import models as b200
    from torch_detection_b200.utils import obj_from_dict as b200_build
    ref_backbone, ref_necks, obj_from_dict = reference_shim.load()
    cfg = dict(type="PAFPN", in_channels=[64, 128, 256, 512], out_channels=256, num_outs=5, activation=activation)
    torch.manual_seed(2)
    neck = obj_from_dict(dict(cfg), parent=ref_necks)
    neck.init_weights()
    neck.eval()
    torch.manual_seed(2)
    mine = b200_build(dict(cfg), parent=b200.necks)
    mine.init_weights()
    a, b = neck.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(2, c, 64 // 2 ** i, 96 // 2 ** i, generator=g) for i, c in enumerate([64, 128, 256, 512])]
    with torch.no_grad():
        ref = neck([f.clone() for f in feats])
        got = orc.pafpn_forward(neck.state_dict(), [f.clone() for f in feats], [64, 128, 256, 512], 256, 5,
                                activation=activation)
    assert len(ref) == len(got) == 5
    assert all(torch.equal(x, y) for x, y in zip(ref, got))


@pytest.mark.parametrize("neck_type", ["FPN", "PAFPN"])
def test_neck_with_batchnorm_matches_reference(neck_type):
    """Row f4: normalize=... puts an (eval-mode) BatchNorm after every neck conv (layers.py:57-135); the oracle
    and the product's parameter mirror against the live reference with randomised running statistics."""
    from torch_detection_b200 import models as b200
    from torch_detection_b200.utils import obj_from_dict as b200_build
    ref_backbone, ref_necks, obj_from_dict = reference_shim.load()
    cfg = dict(type=neck_type, in_channels=[64, 128, 256, 512], out_channels=256, num_outs=5,
               normalize=dict(type="BN"))
    torch.manual_seed(4)
    neck = obj_from_dict(dict(cfg), parent=ref_necks)
    neck.init_weights()
    neck.eval()
    torch.manual_seed(4)
    mine = b200_build(dict(cfg), parent=b200.necks)
    mine.init_weights()
    sd = neck.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(6))
    neck.load_state_dict(sd)
    mine.load_state_dict(sd)
    assert list(sd.keys()) == list(mine.state_dict().keys())
    g = torch.Generator().manual_seed(3)
    feats = [torch.randn(2, c, 64 // 2 ** i, 96 // 2 ** i, generator=g) for i, c in enumerate([64, 128, 256, 512])]
    fwd = orc.fpn_forward if neck_type == "FPN" else orc.pafpn_forward
    with torch.no_grad():
        ref = neck([f.clone() for f in feats])
        got = fwd(sd, [f.clone() for f in feats], [64, 128, 256, 512], 256, 5)
    assert all(torch.equal(x, y) for x, y in zip(ref, got))


def test_resnext_oracle_and_mirror_match_reference():
    """Row f4: ResNeXt-50 32x4d -- the product module mirrors the reference's parameters for the same seed and
    the oracle (grouped conv2 inferred from the parameter shape) is bit-identical to the live reference."""
    from torch_detection_b200 import models as b200
    from torch_detection_b200.utils import obj_from_dict as b200_build
    ref_backbone, ref_necks, obj_from_dict = reference_shim.load()
    cfg = dict(type="ResNeXt", depth=50, base_width=4, cardinality=32)
    torch.manual_seed(1)
    ref = obj_from_dict(dict(cfg), parent=ref_backbone)
    ref.init_weights()
    ref.eval()
    torch.manual_seed(1)
    mine = b200_build(dict(cfg), parent=b200.backbone)
    mine.init_weights()
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
    assert mine.resX_layers == ["layer1", "layer2", "layer3", "layer4"] and mine.feat_dim == ref.feat_dim
    x = torch.randn(1, 3, 64, 96)
    with torch.no_grad():
        want = ref(x)
        got = orc.resnet_forward(a, x, 50)
    assert all(torch.equal(u, v) for u, v in zip(want, got))


@pytest.mark.parametrize("depth", [18, 50])
def test_oracle_groupnorm_matches_reference(depth):
    """use_gn=True (models/utils/layers.py:50-54): GroupNorm(32, C) backbone and neck, eval mode, against the oracle's
    restatement; the affine parameters are randomised (init makes them identity)."""
    ref_backbone, ref_necks, obj_from_dict = reference_shim.load()
    torch.manual_seed(1)
    bb = obj_from_dict(dict(type="ResNet", depth=depth, use_gn=True), parent=ref_backbone)
    bb.init_weights()
    bb.eval()
    chans = [64, 128, 256, 512] if depth < 50 else [256, 512, 1024, 2048]
    neck = obj_from_dict(dict(type="FPN", in_channels=chans, out_channels=256, num_outs=5, normalize=dict(type="GN"),
                              use_gn=True), parent=ref_necks)
    neck.init_weights()
    neck.eval()
    g = torch.Generator().manual_seed(2)
    for m in list(bb.modules()) + list(neck.modules()):
        if isinstance(m, torch.nn.GroupNorm):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
    assert "gn1.weight" in bb.state_dict() and "layer1.0.gn2.bias" in bb.state_dict()
    x = torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        feats = bb(x)
        outs = neck(feats)
        f = orc.resnet_forward(bb.state_dict(), x, depth)
        p = orc.fpn_forward(neck.state_dict(), f, chans, 256, 5)
    for a, b in zip(tuple(feats) + tuple(outs), tuple(f) + tuple(p)):
        assert torch.equal(a, b)
