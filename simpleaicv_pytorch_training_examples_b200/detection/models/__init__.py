from .detr import *  # noqa: F401,F403
