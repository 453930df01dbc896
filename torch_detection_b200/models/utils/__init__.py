from .layers import (ConvModule, conv1x1_group, conv3x3_group, conv7x7_group, norm_layer)
from .inits import (constant_init, kaiming_init, normal_init, uniform_init, xavier_init)
from .checkpoint import load_checkpoint, load_state_dict, save_checkpoint

__all__ = [
    "ConvModule", "conv1x1_group", "conv3x3_group", "conv7x7_group", "norm_layer",
    "constant_init", "kaiming_init", "normal_init", "uniform_init", "xavier_init",
    "load_checkpoint", "load_state_dict", "save_checkpoint",
]
