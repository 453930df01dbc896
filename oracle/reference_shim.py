"""Import shim for the UNMODIFIED reference at /root/reference.  TEST INFRASTRUCTURE ONLY.

The reference does not import on Python >= 3.10 (``from collections import Sequence`` in
datasets/utils/misc.py:6; ``pycocotools`` in datasets/coco.py:2) -- SURVEY.md F5.  This shim
aliases the removed names and stubs pycocotools, then imports the reference's top-level packages
(``models``, ``utils``, ``datasets``) under an isolated ``sys.path`` entry.  It exists only in the
dev container: ``/root/reference`` is absent on the GPU box, so callers must handle ``None``.
"""
import collections
import collections.abc
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TDET_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "backbone", "resnet.py"))


def load():
    """Returns (models_backbone, models_necks, obj_from_dict) of the reference, or None."""
    if not available():
        return None
    for name in ("Sequence", "Mapping"):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))
    if "pycocotools" not in sys.modules:
        m = types.ModuleType("pycocotools")
        mc = types.ModuleType("pycocotools.coco")
        mc.COCO = object
        m.coco = mc
        sys.modules["pycocotools"] = m
        sys.modules["pycocotools.coco"] = mc
    for top in ("models", "utils", "datasets", "core"):
        mod = sys.modules.get(top)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            raise RuntimeError("a non-reference module named %r is already imported" % top)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.backbone as ref_backbone  # noqa: E402
    import models.necks as ref_necks  # noqa: E402
    from utils import obj_from_dict as ref_obj_from_dict  # noqa: E402
    return ref_backbone, ref_necks, ref_obj_from_dict


def build_pair(depth, seed=0, out_channels=256, num_outs=5, **bb_kwargs):
    """Reference ResNet(depth)+FPN built via the reference's own obj_from_dict and init_weights,
    in eval mode (``.eval()`` as a statement: it returns None)."""
    import torch
    ref = load()
    if ref is None:
        return None
    ref_backbone, ref_necks, obj_from_dict = ref
    torch.manual_seed(seed)
    bb = obj_from_dict(dict(type="ResNet", depth=depth, **bb_kwargs), parent=ref_backbone)
    bb.init_weights()
    bb.eval()
    exp = 4 if depth >= 50 else 1
    in_ch = [64 * 2 ** i * exp for i in range(4)]
    neck = obj_from_dict(dict(type="FPN", in_channels=in_ch, out_channels=out_channels,
                              num_outs=num_outs), parent=ref_necks)
    neck.init_weights()
    neck.eval()
    return bb, neck
