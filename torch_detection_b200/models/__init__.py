from . import backbone, necks
from ..registry import BACKBONES, NECKS

__all__ = ["backbone", "necks", "BACKBONES", "NECKS"]
