"""Checkpoint interop (reference models/utils/checkpoint.py:11-169), so that weights saved by the
reference (or torchvision ResNets, whose keys equal ours minus ``fc.*``) flow through the drop-in.

Same call signatures and tolerance rules as the reference: ``load_state_dict`` copies matching
entries, reports unexpected/missing keys, and raises only when ``strict``; ``load_checkpoint``
accepts a bare ``OrderedDict`` or ``{'state_dict': ...}``, strips a ``module.`` prefix and returns
the loaded checkpoint; ``save_checkpoint`` writes ``{'meta', 'state_dict', 'optimizer'}`` with CPU
tensors.  ``modelzoo://`` is resolved through torchvision's weight enums (the reference's
``model_urls`` lookup no longer exists upstream).
"""
import os
import time
from collections import OrderedDict

import torch


def load_state_dict(module, state_dict, strict=False, logger=None):
    own = module.state_dict()
    unexpected = []
    for name, value in state_dict.items():
        if name not in own:
            unexpected.append(name)
            continue
        if isinstance(value, torch.nn.Parameter):
            value = value.data
        try:
            own[name].copy_(value)
        except Exception:
            raise RuntimeError(
                "While copying the parameter named {}, whose dimensions in the model are {} and "
                "whose dimensions in the checkpoint are {}.".format(name, own[name].size(),
                                                                    value.size()))
    missing = sorted(set(own.keys()) - set(state_dict.keys()))
    problems = []
    if unexpected:
        problems.append("unexpected key in source state_dict: {}\n".format(", ".join(unexpected)))
    if missing:
        problems.append("missing keys in source state_dict: {}\n".format(", ".join(missing)))
    if problems:
        text = "\n".join(problems)
        if strict:
            raise RuntimeError(text)
        if logger is not None:
            logger.warning(text)
        else:
            print(text)


def _from_modelzoo(name):
    import torchvision
    try:
        weights = torchvision.models.get_model_weights(name).DEFAULT
    except Exception as exc:  # unknown architecture name
        raise ValueError("Only torchvision architectures are supported in modelzoo, "
                         "{} is not: {}".format(name, exc))
    return weights.get_state_dict(progress=False)


def _trusted(flag):
    return flag or os.environ.get("TDET_TRUSTED_CHECKPOINTS", "0") != "0"


def load_checkpoint(model, filename, map_location=None, strict=False, logger=None, trusted=False):
    """Same signature as the reference plus `trusted`: files are unpickled with ``weights_only=True`` (tensors and
    plain containers, which is all save_checkpoint and the reference write for weights/meta); a checkpoint that
    carries arbitrary pickled objects is only loaded with ``trusted=True`` (or TDET_TRUSTED_CHECKPOINTS=1) --
    unpickling an untrusted file executes code."""
    if filename.startswith("modelzoo://"):
        checkpoint = _from_modelzoo(filename[len("modelzoo://"):])
    elif filename.startswith(("http://", "https://")):
        checkpoint = torch.hub.load_state_dict_from_url(filename, map_location=map_location,
                                                        weights_only=not _trusted(trusted))
    else:
        if not os.path.isfile(filename):
            raise IOError("{} is not a checkpoint file".format(filename))
        try:
            checkpoint = torch.load(filename, map_location=map_location, weights_only=True)
        except Exception as exc:
            if not _trusted(trusted):
                raise RuntimeError("{} holds more than tensors and plain containers ({}); pass trusted=True to "
                                   "unpickle it if you trust its source".format(filename, exc))
            checkpoint = torch.load(filename, map_location=map_location, weights_only=False)
    if isinstance(checkpoint, dict) and "state_dict" in checkpoint:
        state_dict = checkpoint["state_dict"]
    elif isinstance(checkpoint, dict):  # OrderedDict of tensors
        state_dict = checkpoint
    else:
        raise RuntimeError("No state_dict found in checkpoint file {}".format(filename))
    if state_dict and next(iter(state_dict)).startswith("module."):
        state_dict = OrderedDict((k[7:], v) for k, v in state_dict.items())
    target = model.module if hasattr(model, "module") else model
    load_state_dict(target, state_dict, strict, logger)
    return checkpoint


def weights_to_cpu(state_dict):
    return OrderedDict((k, v.cpu()) for k, v in state_dict.items())


def save_checkpoint(model, filename, optimizer=None, meta=None):
    if meta is None:
        meta = {}
    elif not isinstance(meta, dict):
        raise TypeError("meta must be a dict or None, but got {}".format(type(meta)))
    meta.update(time=time.asctime())
    folder = os.path.dirname(filename)
    if folder:
        os.makedirs(folder, exist_ok=True)
    target = model.module if hasattr(model, "module") else model
    payload = {"meta": meta, "state_dict": weights_to_cpu(target.state_dict())}
    if optimizer is not None:
        payload["optimizer"] = optimizer.state_dict()
    torch.save(payload, filename)
