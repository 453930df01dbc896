"""Shared test helpers: build the product modules the way the parity protocol says
and obtain oracle inputs from them."""
import hashlib
import os
from collections import OrderedDict

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["r18_fpn_64x64", "r50_fpn_64x96", "r50_fpn_64x64_bnstats"]


def state_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def in_channels_for(depth):
    exp = 4 if depth >= 50 else 1
    return [64 * 2 ** i * exp for i in range(4)]


def build_product_pair(depth, seed=0, out_channels=256, num_outs=5, bnstats=False, **bb_kwargs):
    """Product ResNet+FPN built through the product's registry API with the reference's seeding
    protocol (manual_seed; build backbone; init; build neck; init).  Construction consumes the RNG
    exactly like the reference, so the weights equal the reference's for the same seed."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    from oracle import resnet_fpn_oracle as orc
    torch.manual_seed(seed)
    bb = obj_from_dict(dict(type="ResNet", depth=depth, **bb_kwargs), parent=models.backbone)
    bb.init_weights()
    bb.eval()
    neck = obj_from_dict(dict(type="FPN", in_channels=in_channels_for(depth),
                              out_channels=out_channels, num_outs=num_outs), parent=models.necks)
    neck.init_weights()
    neck.eval()
    if bnstats:
        sd = bb.state_dict()
        g = torch.Generator().manual_seed(1000 + seed)
        orc.randomize_bn_stats(sd, generator=g)
        bb.load_state_dict(sd)
    return bb, neck


def build_product_gn_pair(depth=50, seed=2, out_channels=256, num_outs=5):
    """use_gn=True backbone + neck with the seeding protocol of oracle/make_golden.py:make_groupnorm_golden."""
    from torch_detection_b200 import models
    from torch_detection_b200.utils import obj_from_dict
    from oracle import resnet_fpn_oracle as orc
    torch.manual_seed(seed)
    bb = obj_from_dict(dict(type="ResNet", depth=depth, use_gn=True), parent=models.backbone)
    bb.init_weights()
    bb.eval()
    neck = obj_from_dict(dict(type="FPN", in_channels=in_channels_for(depth), out_channels=out_channels,
                              num_outs=num_outs, normalize=dict(type="GN"), use_gn=True), parent=models.necks)
    neck.init_weights()
    neck.eval()
    g = torch.Generator().manual_seed(1000 + seed)
    for mod in (bb, neck):
        sd = mod.state_dict()
        orc.randomize_gn_affine(sd, generator=g)
        mod.load_state_dict(sd)
    return bb, neck


def cpu_state(module):
    return OrderedDict((k, v.detach().cpu().clone()) for k, v in module.state_dict().items())


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("meta_")}
    return meta, arrays
