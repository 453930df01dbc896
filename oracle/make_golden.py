"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference via the import
shim).  Dev-container only; the fixtures it writes are committed and travel to the GPU box.

    python oracle/make_golden.py

Each fixture holds: the build config, the seeds, a SHA-256 of the reference's state_dict bytes (so a
consumer can prove it rebuilt the same weights from the seed), the input batch and every output
level (C2..C5, P2..P6) in fp32, exactly as the reference produced them on CPU.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_shim, resnet_fpn_oracle as orc  # noqa: E402

CASES = [
    # name, depth, seed, input shape, randomise BN stats
    ("r18_fpn_64x64", 18, 0, (1, 3, 64, 64), False),
    ("r50_fpn_64x96", 50, 0, (1, 3, 64, 96), False),
    ("r50_fpn_64x64_bnstats", 50, 1, (1, 3, 64, 64), True),
]


def state_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def grad_signature(g):
    """Compact fingerprint of a gradient tensor: L2 norm, sum, a fixed pseudo-random projection, first 8 values."""
    g = g.detach().double().flatten()
    idx = torch.arange(g.numel(), dtype=torch.float64)
    proj = torch.cos(0.37 * idx + 0.11)
    return np.array([float(g.norm()), float(g.sum()), float((g * proj).sum())] + [float(v) for v in g[:8]])


def make_gradient_golden(out_dir):
    """tests/golden/r18_fpn_64x64_grads.npz: the REFERENCE's own autograd (training config: frozen BN statistics,
    stage 1 frozen) on the r18 fixture's input with seeded upstream gradients; one fingerprint per parameter."""
    depth, seed = 18, 0
    bb, neck = reference_shim.build_pair(depth, seed=seed)
    sd = bb.state_dict()
    orc.randomize_bn_stats(sd, generator=torch.Generator().manual_seed(1000 + seed))
    bb.load_state_dict(sd)
    g = torch.Generator().manual_seed(77 + seed)
    x = torch.randn(1, 3, 64, 64, generator=g)
    outs = neck(bb(x))
    gg = torch.Generator().manual_seed(99)
    grads = [torch.randn(o.shape, generator=gg) for o in outs]
    torch.autograd.backward(list(outs), grads)
    arrays = {}
    for prefix, mod in (("bb.", bb), ("neck.", neck)):
        for k, p in mod.named_parameters():
            if p.grad is not None and not (prefix == "bb." and (k.startswith("layer1") or not k.startswith("layer"))):
                arrays[prefix + k] = grad_signature(p.grad)
    meta = dict(depth=depth, seed=seed, input_seed=77 + seed, grad_seed=99,
                bb_hash=state_hash(bb.state_dict()), neck_hash=state_hash(neck.state_dict()))
    np.savez(os.path.join(out_dir, "r18_fpn_64x64_grads.npz"), **arrays,
             **{"meta_" + k: np.array(v) for k, v in meta.items()})
    print("r18_fpn_64x64_grads", len(arrays), "parameter gradients")


def make_groupnorm_golden(out_dir):
    """tests/golden/r50_gn_fpn_64x96.npz: use_gn=True backbone and neck (nn.GroupNorm(32, C) after every conv,
    models/utils/layers.py:50-54), built through the reference's registry, GroupNorm affines randomised."""
    depth, seed = 50, 2
    ref_backbone, ref_necks, obj_from_dict = reference_shim.load()
    torch.manual_seed(seed)
    bb = obj_from_dict(dict(type="ResNet", depth=depth, use_gn=True), parent=ref_backbone)
    bb.init_weights()
    bb.eval()
    neck = obj_from_dict(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5,
                              normalize=dict(type="GN"), use_gn=True), parent=ref_necks)
    neck.init_weights()
    neck.eval()
    g = torch.Generator().manual_seed(1000 + seed)
    for mod in (bb, neck):
        sd = mod.state_dict()
        orc.randomize_gn_affine(sd, generator=g)
        mod.load_state_dict(sd)
    # The input is bf16-representable: a chain of GroupNorms over 2x3 .. 16x24 maps amplifies the 2^-9 rounding of a
    # bf16 image hand-off to 3.5e-3 .. 6.4e-3 at the outputs even in exact arithmetic (measured with the oracle), which
    # would leave no budget for the path under test inside the 1e-2 gate.
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(77 + seed)).to(torch.bfloat16).float()
    with torch.no_grad():
        feats = bb(x)
        outs = neck(feats)
    arrays = {"x": x.numpy()}
    for i, t in enumerate(feats):
        arrays["C%d" % (i + 2)] = t.numpy()
    for i, t in enumerate(outs):
        arrays["P%d" % (i + 2)] = t.numpy()
    meta = dict(depth=depth, seed=seed, input_seed=77 + seed, bb_hash=state_hash(bb.state_dict()),
                neck_hash=state_hash(neck.state_dict()))
    np.savez(os.path.join(out_dir, "r50_gn_fpn_64x96.npz"), **arrays,
             **{"meta_" + k: np.array(v) for k, v in meta.items()})
    print("r50_gn_fpn_64x96", {k: v.shape for k, v in arrays.items()})


def main():
    assert reference_shim.available(), "needs /root/reference"
    torch.set_num_threads(1)  # fixed reduction order
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, depth, seed, shape, bnstats in CASES:
        bb, neck = reference_shim.build_pair(depth, seed=seed)
        if bnstats:
            sd = bb.state_dict()
            g = torch.Generator().manual_seed(1000 + seed)
            orc.randomize_bn_stats(sd, generator=g)
            bb.load_state_dict(sd)
        g = torch.Generator().manual_seed(77 + seed)
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            feats = bb(x)
            outs = neck(feats)
        arrays = {"x": x.numpy()}
        for i, t in enumerate(feats):
            arrays["C%d" % (i + 2)] = t.numpy()
        for i, t in enumerate(outs):
            arrays["P%d" % (i + 2)] = t.numpy()
        meta = dict(depth=depth, seed=seed, bnstats=int(bnstats), input_seed=77 + seed,
                    bb_hash=state_hash(bb.state_dict()), neck_hash=state_hash(neck.state_dict()))
        np.savez(os.path.join(out_dir, name + ".npz"),
                 **arrays, **{"meta_" + k: np.array(v) for k, v in meta.items()})
        print(name, {k: v.shape for k, v in arrays.items()})
    make_gradient_golden(out_dir)
    make_groupnorm_golden(out_dir)


if __name__ == "__main__":
    main()
