"""DETR's ResNet body and sine position embedding with the reference's constructor surface and state_dict layout
(SimpleAICV/detection/models/backbones/detr_resnet.py:28-64 PositionEmbeddingBlock, :256-340 DetrResNetBackbone,
:343-384 constructors).  The blocks are the classification ResNet's parameter containers (same module names, so
`backbone.layer1.0.conv1.layer.0.weight` etc.); they are executed by engine.detr.DetrRT through engine.convnet."""
import math

import torch
import torch.nn as nn

from ....classification.backbones.resnet import BasicBlock, Bottleneck, ConvBnActBlock, _init_like_reference

__all__ = ['detr_resnet18backbone', 'detr_resnet34backbone', 'detr_resnet50backbone', 'detr_resnet101backbone',
           'detr_resnet152backbone']


class PositionEmbeddingBlock(nn.Module):
    """Parameter-free sine embedding of the padding masks (index arithmetic on [B, h, w] bools; runs in torch)."""

    def __init__(self, inplanes=128, temperature=10000, eps=1e-6):
        super().__init__()
        self.inplanes, self.temperature, self.eps = inplanes, temperature, eps
        self.scale = 2 * math.pi

    def forward(self, masks):
        assert masks is not None
        not_masks = ~masks
        y_embed = torch.cumsum(not_masks, 1, dtype=torch.float32)
        x_embed = torch.cumsum(not_masks, 2, dtype=torch.float32)
        y_embed = y_embed / (y_embed[:, -1:, :] + self.eps) * self.scale
        x_embed = x_embed / (x_embed[:, :, -1:] + self.eps) * self.scale
        dim_t = torch.arange(self.inplanes, dtype=torch.float32, device=masks.device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.inplanes)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class DetrResNetBackbone(nn.Module):

    def __init__(self, block, layer_nums, inplanes=64, use_gradient_checkpoint=False):
        super().__init__()
        self.block, self.layer_nums, self.inplanes = block, layer_nums, inplanes
        self.planes = [inplanes, inplanes * 2, inplanes * 4, inplanes * 8]
        self.expansion = block.expansion
        self.use_gradient_checkpoint = use_gradient_checkpoint
        self.conv1 = ConvBnActBlock(3, inplanes, kernel_size=7, stride=2, padding=3)
        self.maxpool1 = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for i, (planes, stride) in enumerate(zip(self.planes, (1, 2, 2, 2))):
            blocks = []
            for j in range(layer_nums[i]):
                blocks.append(block(self.inplanes, planes, stride if j == 0 else 1))
                self.inplanes = planes * self.expansion
            setattr(self, f'layer{i + 1}', nn.Sequential(*blocks))
        self.out_channels = [p * self.expansion for p in self.planes]
        _init_like_reference(self)

    def forward(self, x):
        raise RuntimeError('DetrResNetBackbone is executed by engine.detr.DetrRT; call the DETR model')


def _detrresnetbackbone(block, layers, inplanes, pretrained_path='', **kwargs):
    model = DetrResNetBackbone(block, layers, inplanes, **kwargs)
    if pretrained_path:
        from ....tools.utils import load_model_state
        load_model_state(model, torch.load(pretrained_path, map_location='cpu'))
    return model


def detr_resnet18backbone(pretrained_path='', **kwargs):
    return _detrresnetbackbone(BasicBlock, [2, 2, 2, 2], 64, pretrained_path=pretrained_path, **kwargs)


def detr_resnet34backbone(pretrained_path='', **kwargs):
    return _detrresnetbackbone(BasicBlock, [3, 4, 6, 3], 64, pretrained_path=pretrained_path, **kwargs)


def detr_resnet50backbone(pretrained_path='', **kwargs):
    return _detrresnetbackbone(Bottleneck, [3, 4, 6, 3], 64, pretrained_path=pretrained_path, **kwargs)


def detr_resnet101backbone(pretrained_path='', **kwargs):
    return _detrresnetbackbone(Bottleneck, [3, 4, 23, 3], 64, pretrained_path=pretrained_path, **kwargs)


def detr_resnet152backbone(pretrained_path='', **kwargs):
    return _detrresnetbackbone(Bottleneck, [3, 8, 36, 3], 64, pretrained_path=pretrained_path, **kwargs)
