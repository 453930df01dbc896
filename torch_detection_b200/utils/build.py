"""``obj_from_dict`` -- the reference's build function (utils/utils.py:5-38).

``obj_from_dict(dict(type='ResNet', depth=50), parent=models.backbone)`` resolves ``type`` as an
attribute of ``parent`` (or as a module name when ``parent`` is None, like the reference) and calls
it with the remaining items.  Differences from the reference, both deliberate:
  * ``additional_dict`` is iterated with ``.items()`` (the reference iterates keys and raises
    ``ValueError`` for any non-empty dict, utils/utils.py:36);
  * no import of the data layer (the reference pulls cv2/pycocotools in through
    ``datasets.utils.is_str``, utils/utils.py:2).
"""
import sys


def obj_from_dict(args_dict, parent=None, additional_dict=None):
    assert isinstance(args_dict, dict) and "type" in args_dict
    assert isinstance(additional_dict, dict) or additional_dict is None
    kwargs = dict(args_dict)
    target = kwargs.pop("type")
    if isinstance(target, str):
        target = getattr(parent, target) if parent is not None else sys.modules[target]
    elif not isinstance(target, type):
        raise TypeError("type must be a str or valid type, but got {}".format(type(target)))
    if additional_dict is not None:
        for key, value in additional_dict.items():
            kwargs.setdefault(key, value)
    return target(**kwargs)
