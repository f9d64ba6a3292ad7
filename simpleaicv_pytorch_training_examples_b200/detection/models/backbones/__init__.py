from .detr_resnet import *  # noqa: F401,F403
