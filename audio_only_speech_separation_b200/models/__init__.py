"""Model registry with the lookup contract of look2hear/models/__init__.py:29-56."""
from .base_model import BaseModel
from .sepformer import Sepformer
from .tasnet import TasNet

__all__ = ["TasNet", "Sepformer", "BaseModel"]


def register_model(custom_model):
    """Register a custom model, gettable with :func:`get` (models/__init__.py:29-38)."""
    if custom_model.__name__ in globals().keys() or custom_model.__name__.lower() in globals().keys():
        raise ValueError(f"Model {custom_model.__name__} already exists. Choose another name.")
    globals().update({custom_model.__name__: custom_model})


def get(identifier):
    """Model class from a case-insensitive name (models/__init__.py:41-56)."""
    if isinstance(identifier, str):
        to_get = {k.lower(): v for k, v in globals().items()}
        cls = to_get.get(identifier.lower())
        if cls is None:
            raise ValueError(f"Could not interpret model name : {str(identifier)}")
        return cls
    raise ValueError(f"Could not interpret model name : {str(identifier)}")
