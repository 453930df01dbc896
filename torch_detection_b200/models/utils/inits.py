"""Parameter initialisers with the reference's names and defaults (models/utils/inits.py:5-46).
They define the "random-init weights" the parity gate uses."""
import torch.nn as nn


def _set_bias(module, value):
    b = getattr(module, "bias", None)
    if b is not None:
        nn.init.constant_(b, value)


def constant_init(module, val, bias=0):
    nn.init.constant_(module.weight, val)
    _set_bias(module, bias)


def xavier_init(module, gain=1, bias=0, distribution="normal"):
    assert distribution in ["uniform", "normal"]
    fn = nn.init.xavier_uniform_ if distribution == "uniform" else nn.init.xavier_normal_
    fn(module.weight, gain=gain)
    _set_bias(module, bias)


def normal_init(module, mean=0, std=1, bias=0):
    nn.init.normal_(module.weight, mean, std)
    _set_bias(module, bias)


def uniform_init(module, a=0, b=1, bias=0):
    nn.init.uniform_(module.weight, a, b)
    _set_bias(module, bias)


def kaiming_init(module, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    assert distribution in ["uniform", "normal"]
    fn = nn.init.kaiming_uniform_ if distribution == "uniform" else nn.init.kaiming_normal_
    fn(module.weight, mode=mode, nonlinearity=nonlinearity)
    _set_bias(module, bias)
