"""Module registries of the build API (mirrors reference models/registry.py:4-41).

``BACKBONES`` / ``NECKS`` map class names to classes; classes self-register with the
``@X.register_module`` decorator, which returns the class unchanged.  Registering a non-``nn.Module``
raises ``TypeError``; registering the same name twice raises ``KeyError`` -- same conventions as the
reference.
"""
import torch.nn as nn


class Registry(object):
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __contains__(self, key):
        return key in self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def _register_module(self, module_class):
        if not (isinstance(module_class, type) and issubclass(module_class, nn.Module)):
            raise TypeError("module must be a child of nn.Module, but got {}".format(
                type(module_class)))
        key = module_class.__name__
        if key in self._module_dict:
            raise KeyError("{} is already registered in {}".format(key, self._name))
        self._module_dict[key] = module_class

    def register_module(self, cls):
        self._register_module(cls)
        return cls


BACKBONES = Registry("backbone")
NECKS = Registry("neck")
