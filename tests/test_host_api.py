"""CPU tests of the host-side mirror of the reference interface and of the C-ABI library surface
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

from torch_detection_b200 import _C, models
from torch_detection_b200.models.backbone import ResNet
from torch_detection_b200.models.necks import FPN
from torch_detection_b200.registry import BACKBONES, NECKS, Registry
from torch_detection_b200.utils import obj_from_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_and_build_api():
    assert BACKBONES.module_dict["ResNet"] is ResNet and NECKS.module_dict["FPN"] is FPN
    bb = obj_from_dict(dict(type="ResNet", depth=50, frozen_stages=1, bn_eval=True, bn_frozen=True),
                       parent=models.backbone)
    neck = obj_from_dict(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256,
                              num_outs=5), parent=models.necks)
    assert isinstance(bb, ResNet) and isinstance(neck, FPN)
    assert bb.feat_dim == 2048 and bb.inplanes == 2048 and bb.res_layers == ["layer1", "layer2", "layer3", "layer4"]
    assert obj_from_dict(dict(type=ResNet, depth=18), additional_dict=dict(frozen_stages=2)).frozen_stages == 2
    with pytest.raises(TypeError):
        obj_from_dict(dict(type=3))
    r = Registry("x")
    with pytest.raises(TypeError):
        r.register_module(int)
    r.register_module(ResNet)
    with pytest.raises(KeyError):
        r.register_module(ResNet)


def test_state_dict_layout():
    sd = ResNet(50).state_dict()
    assert len(sd) == 318
    assert sd["conv1.weight"].shape == (64, 3, 7, 7)
    assert sd["layer2.0.downsample.0.weight"].shape == (512, 256, 1, 1)
    assert sd["layer4.2.conv3.weight"].shape == (2048, 512, 1, 1)
    assert "layer1.0.bn1.num_batches_tracked" in sd
    try:
        import torchvision
        tv = torchvision.models.resnet50().state_dict()
        assert set(sd.keys()) == {k for k in tv if not k.startswith("fc.")}
    except ImportError:
        pass
    nsd = FPN([256, 512, 1024, 2048], 256, 5).state_dict()
    assert len(nsd) == 16 and nsd["fpn_convs.3.conv.weight"].shape == (256, 256, 3, 3)
    assert nsd["lateral_convs.3.conv.bias"].shape == (256,)
    assert len(ResNet(18).state_dict()) == 120


def test_constructor_errors():
    with pytest.raises(KeyError):
        ResNet(42)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=5)
    with pytest.raises(AssertionError):
        ResNet(50, num_stages=2, strides=(1, 2), dilations=(1, 1), out_indices=(0, 3))
    with pytest.raises(AssertionError):
        FPN((256, 512), 256, 5)
    with pytest.raises(AssertionError):
        FPN([256, 512, 1024], 256, 2)
    with pytest.raises(AssertionError):
        FPN([256, 512, 1024], 256, 5, end_level=2)
    gn = ResNet(50, use_gn=True)          # resnet.py:215, :85-86: norm modules are named gn*
    assert isinstance(gn.gn1, torch.nn.GroupNorm) and gn.gn1.num_groups == 32 and not hasattr(gn, "bn_eval")
    assert isinstance(gn.layer1[0].gn3, torch.nn.GroupNorm) and isinstance(gn.layer1[0].downsample[1], torch.nn.GroupNorm)
    assert isinstance(FPN([256, 512], 256, 2, normalize=dict(type="GN"), use_gn=True).fpn_convs[1].norm,
                      torch.nn.GroupNorm)
    assert FPN([256, 512], 256, 2, normalize=dict(type="BN")).lateral_convs[0].with_norm
    with pytest.raises(TypeError):
        ResNet(18).init_weights(pretrained=3)


def test_train_semantics():
    """resnet.py:270-294 with the intended stage freezing and `return self` (F4)."""
    bb = ResNet(50, frozen_stages=1, bn_eval=True, bn_frozen=True)
    assert bb.bn1.training  # fresh module: BN still in training mode, like the reference
    assert bb.train() is bb and bb.eval() is bb
    bb.train()
    assert all(not m.training for m in bb.modules() if isinstance(m, nn.BatchNorm2d))
    assert not any(p.requires_grad for p in bb.conv1.parameters())
    assert not any(p.requires_grad for p in bb.layer1.parameters())
    assert not bb.layer1.training and bb.layer2.training
    trainable = [n for n, p in bb.named_parameters() if p.requires_grad]
    assert len(trainable) == 42 and all(".conv" in n or "downsample.0" in n for n in trainable)
    neck = FPN([256, 512, 1024, 2048], 256, 5)
    total = sum(p.numel() for p in bb.parameters() if p.requires_grad) + sum(p.numel() for p in neck.parameters())
    assert total == 26576896  # 58 tensors, 26.577 M elements
    bb2 = ResNet(18, bn_eval=False)
    bb2.train()
    assert bb2.bn1.training


def test_init_statistics():
    torch.manual_seed(0)
    bb = ResNet(50)
    bb.init_weights()
    w = bb.layer3[0].conv2.weight
    assert abs(float(w.std()) - (2.0 / (256 * 9)) ** 0.5) < 2e-4
    assert float(bb.bn1.weight.min()) == 1.0 and float(bb.bn1.bias.abs().max()) == 0.0
    neck = FPN([256, 512, 1024, 2048], 256, 5)
    neck.init_weights()
    bound = (6.0 / (2 * 256 * 9)) ** 0.5
    assert float(neck.fpn_convs[0].conv.weight.abs().max()) <= bound
    assert float(neck.fpn_convs[0].conv.bias.abs().max()) == 0.0


def test_checkpoint_roundtrip(tmp_path):
    from torch_detection_b200.models.utils import load_checkpoint, save_checkpoint
    torch.manual_seed(1)
    a = ResNet(18)
    a.init_weights()
    path = str(tmp_path / "ck" / "r18.pth")
    save_checkpoint(a, path, meta=dict(epoch=3))
    b = ResNet(18)
    ck = load_checkpoint(b, path, strict=True)
    assert ck["meta"]["epoch"] == 3
    assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    c = ResNet(18)
    c.init_weights(pretrained=path)
    assert torch.equal(c.conv1.weight, a.conv1.weight)
    with pytest.raises(IOError):
        load_checkpoint(b, str(tmp_path / "missing.pth"))


def test_cpu_input_is_refused_not_emulated():
    bb = ResNet(18).eval()
    with pytest.raises(NotImplementedError):
        bb(torch.randn(1, 3, 64, 64))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tdet_b200.h")).read()
    declared = set(re.findall(r"\b(tdet_[a-z0-9_]+)\s*\(", header))
    declared -= {"tdet_plan"}
    assert declared == set(_C.EXPORTS)
    lib = ctypes.CDLL(_C.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _C.lib().tdet_abi_version() == _C.ABI_VERSION
    assert ctypes.sizeof(_C.TdetOp) == 384  # static_assert(sizeof(tdet_op) == 384) in tdet_api.cu


def test_binding_constants_match_the_header():
    """Every TDET_OP_* / TDET_FLAG_* / dtype / status value the ctypes binding hard-codes equals the header's enum (a
    drifted constant would silently launch a different op)."""
    header = open(os.path.join(ROOT, "include", "tdet_b200.h")).read()
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(TDET_[A-Z0-9_]+)\s*=\s*(-?\d+)", header)}
    for prefix, strip in (("TDET_OP_", "OP_"), ("TDET_FLAG_", "FLAG_"), ("TDET_ERR_", "ERR_")):
        names = [n for n in enums if n.startswith(prefix)]
        assert names, prefix
        for n in names:
            py = strip + n[len(prefix):]
            assert hasattr(_C, py), "%s has no binding constant %s" % (n, py)
            assert getattr(_C, py) == enums[n], (n, enums[n], getattr(_C, py))
    for n, py in (("TDET_BF16", "BF16"), ("TDET_F32", "F32"), ("TDET_F16", "F16"), ("TDET_U8", "U8")):
        assert getattr(_C, py) == enums[n]
    flags = [v for n, v in enums.items() if n.startswith("TDET_FLAG_")]
    assert len(set(flags)) == len(flags) and all(v & (v - 1) == 0 for v in flags), "flags must be distinct bits"


def test_no_gpu_calls_fail_cleanly_without_device():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = _C.lib().tdet_device_supported(0)
    assert rc < 0 and len(_C.lib().tdet_last_error()) > 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torch_detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)


def test_plan_cache_is_a_bounded_lru():
    """Plans pin their activation arenas; the cache keeps at most `capacity` groups (a training forward plan and its
    backward plan form one group and leave together), least recently used first."""
    from torch_detection_b200.engine import PlanCache
    c = PlanCache(capacity=2)
    c["a"] = 1
    c.put(("bwd", "a"), 11, group="a")
    c["b"] = 2
    assert c.get("a") == 1 and c.get(("bwd", "a"), group="a") == 11 and len(c) == 3
    c["c"] = 3                      # evicts the least recently used group: b
    assert c.get("b") is None and c.get("a") == 1 and c.get("c") == 3
    c["d"] = 4                      # a (forward AND backward entry) is now the oldest
    assert c.get("a") is None and c.get(("bwd", "a"), group="a") is None and len(c) == 2
    c.clear()
    assert len(c) == 0
    assert PlanCache().capacity >= 1
    assert isinstance(ResNet(18)._plans, PlanCache) and isinstance(FPN([64, 128, 256, 512], 256, 5)._plans, PlanCache)


def test_operand_cache_force_refresh():
    """OperandCache.refresh(force=True) re-derives every entry with dependencies (what invalidate_operands() uses
    after in-place updates through .data, which do not bump tensor._version)."""
    from torch_detection_b200.engine import OperandCache
    w = nn.Parameter(torch.ones(4))
    cache = OperandCache()
    calls = []

    def make(out):
        calls.append(out is None)
        if out is None:
            return w.detach().clone()
        out.copy_(w.detach())
        return out

    v = cache.get("w", make, deps=(w,))
    const = cache.get("const", lambda out: torch.zeros(1))
    assert cache.refresh() == 0
    w.data.mul_(2.0)                         # invisible to the version check
    assert cache.refresh() == 0 and float(v[0]) == 1.0
    assert cache.refresh(force=True) == 1 and float(v[0]) == 2.0 and const is cache.value("const")
    with torch.no_grad():
        w.mul_(2.0)                          # a normal in-place update is seen
    assert cache.refresh() == 1 and float(v[0]) == 4.0
    assert calls == [True, False, False]
