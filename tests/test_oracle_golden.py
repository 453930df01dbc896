"""Oracle vs the committed golden vectors (produced by the reference itself, oracle/make_golden.py).
CPU-only; runs on the GPU box too, where the reference tree is absent."""
import os

import pytest
import torch

from oracle import resnet_fpn_oracle as orc
from tests import helpers


@pytest.mark.parametrize("name", helpers.GOLDEN_CASES)
def test_oracle_matches_golden(name):
    torch.set_num_threads(1)
    meta, arrays = helpers.load_golden(name)
    bb, neck = helpers.build_product_pair(meta["depth"], seed=meta["seed"], bnstats=bool(meta["bnstats"]))
    # the product's host-side mirror rebuilds the reference's weights from the seed
    assert helpers.state_hash(bb.state_dict()) == meta["bb_hash"]
    assert helpers.state_hash(neck.state_dict()) == meta["neck_hash"]
    feats, outs = orc.resnet_fpn_forward(helpers.cpu_state(bb), helpers.cpu_state(neck), arrays["x"],
                                         meta["depth"])
    for i, t in enumerate(feats):
        assert torch.equal(t, arrays["C%d" % (i + 2)]), "C%d" % (i + 2)
    for i, t in enumerate(outs):
        assert torch.equal(t, arrays["P%d" % (i + 2)]), "P%d" % (i + 2)


def test_flop_model_matches_survey():
    assert orc.conv_flops(50, 800, 1344)[0] == 296961638400
    assert orc.conv_flops(101, 800, 1344)[0] == 456056832000
    assert orc.conv_flops(18, 800, 1344)[0] == 187136409600
    assert orc.conv_flops(50, 800, 1333, with_fpn=False)[0] == 174666598400


def test_bf16_emulation_is_within_gate():
    """The emulated fused-bf16 pipeline (what the kernels implement) vs the fp32 oracle: documents the
    margin the <=1e-2 gate leaves."""
    meta, arrays = helpers.load_golden("r50_fpn_64x96")
    bb, neck = helpers.build_product_pair(50, seed=0)
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    _, outs = orc.resnet_fpn_forward(bsd, nsd, arrays["x"], 50)
    _, emu = orc.resnet_fpn_forward_bf16_emulated(bsd, nsd, arrays["x"], 50)
    errs = [orc.rel_l2(a, b) for a, b in zip(emu, outs)]
    assert max(errs) < 1.5e-2, errs


def test_gradient_oracle_reproduces_reference_gradient_fixture():
    """tests/golden/r18_fpn_64x64_grads.npz holds fingerprints of the REFERENCE's own autograd gradients
    (oracle/make_golden.py, dev container).  The gradient oracle must reproduce them anywhere -- this is what
    pins oracle/grad_oracle.py on the GPU box, where the reference tree does not exist."""
    import numpy as np
    from oracle import grad_oracle, make_golden
    z = np.load(os.path.join(helpers.GOLDEN_DIR, "r18_fpn_64x64_grads.npz"))
    seed = int(z["meta_seed"])
    bb, neck = helpers.build_product_pair(18, seed=seed, bnstats=True)
    assert helpers.state_hash(bb.state_dict()) == str(z["meta_bb_hash"])
    assert helpers.state_hash(neck.state_dict()) == str(z["meta_neck_hash"])
    bsd, nsd = helpers.cpu_state(bb), helpers.cpu_state(neck)
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(int(z["meta_input_seed"])))
    with torch.no_grad():
        _, outs = orc.resnet_fpn_forward(bsd, nsd, x, 18)
    gg = torch.Generator().manual_seed(int(z["meta_grad_seed"]))
    grads = [torch.randn(o.shape, generator=gg) for o in outs]
    gb, gn, _, _ = grad_oracle.plain_grads(bsd, nsd, x, 18, grads, train_from_stage=1, bn_affine=True)
    keys = [k for k in z.files if not k.startswith("meta_")]
    assert len(keys) >= 40
    for k in keys:
        got = (gb if k.startswith("bb.") else gn)[k.split(".", 1)[1]]
        sig = make_golden.grad_signature(got)
        assert np.allclose(sig, z[k], rtol=2e-4, atol=1e-6 * abs(z[k][0])), (k, sig[:3], z[k][:3])


def test_oracle_groupnorm_matches_golden():
    """use_gn=True fixture produced by the unmodified reference (oracle/make_golden.py:make_groupnorm_golden): the
    product's host-side mirror rebuilds the same state (gn* keys) and the oracle's GroupNorm restatement reproduces
    every level bit for bit."""
    torch.set_num_threads(1)
    meta, arrays = helpers.load_golden("r50_gn_fpn_64x96")
    bb, neck = helpers.build_product_gn_pair(meta["depth"], seed=meta["seed"])
    assert "gn1.weight" in bb.state_dict() and "layer1.0.gn3.bias" in bb.state_dict()
    assert "lateral_convs.0.norm.weight" in neck.state_dict() and "lateral_convs.0.conv.bias" not in neck.state_dict()
    assert helpers.state_hash(bb.state_dict()) == meta["bb_hash"]
    assert helpers.state_hash(neck.state_dict()) == meta["neck_hash"]
    feats = orc.resnet_forward(helpers.cpu_state(bb), arrays["x"], 50)
    outs = orc.fpn_forward(helpers.cpu_state(neck), feats, [256, 512, 1024, 2048], 256, 5)
    for i, t in enumerate(feats):
        assert torch.equal(t, arrays["C%d" % (i + 2)]), "C%d" % (i + 2)
    for i, t in enumerate(outs):
        assert torch.equal(t, arrays["P%d" % (i + 2)]), "P%d" % (i + 2)
