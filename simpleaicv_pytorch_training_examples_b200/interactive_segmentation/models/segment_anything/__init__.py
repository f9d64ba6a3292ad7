"""``sam.__dict__[name](**kwargs)`` surface of SimpleAICV/interactive_segmentation/models/segment_anything."""
from . import sam  # noqa: F401
from .image_encoder import ViTImageEncoder  # noqa: F401
