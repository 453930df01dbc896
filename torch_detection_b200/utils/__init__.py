from .build import obj_from_dict

__all__ = ["obj_from_dict"]
