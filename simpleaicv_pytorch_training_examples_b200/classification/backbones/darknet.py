"""Darknet-53 with the reference's constructor surface and state_dict layout
(SimpleAICV/classification/backbones/darknet.py:16-31 ActivationBlock, :34-65 ConvBnActBlock,
:116-144 Darknet53Block, :323-432 Darknet53), executed by engine.convnet on sm_100a kernels.
Layers whose channel counts are not multiples of 64 (conv1: 3->32, conv2: 32->64) run on
channel-padded activations inside the runtime; parameters keep the reference's shapes.

darknettiny / darknet19 (2x2 max-pools, 16/32-channel stacks, 1x1-conv classifier) are not
implemented yet and raise NotImplementedError.
"""
import torch.nn as nn

from ...engine.convnet import ACT_LEAKY, ACT_RELU, ConvBN, DarkBlockRT, PlainUnitRT, ResNetRT, run_network

__all__ = ['darknettiny', 'darknet19', 'darknet53']


class ActivationBlock(nn.Module):

    def __init__(self, act_type='leakyrelu', inplace=True):
        super().__init__()
        assert act_type in ['silu', 'relu', 'leakyrelu'], 'Unsupport activation function!'
        if act_type == 'silu':
            raise NotImplementedError('act_type silu is not implemented by the B200 runtime (relu / leakyrelu are)')
        self.act = nn.ReLU(inplace=inplace) if act_type == 'relu' else nn.LeakyReLU(0.1, inplace=inplace)


class ConvBnActBlock(nn.Module):

    def __init__(self, inplanes, planes, kernel_size, stride, padding, groups=1, has_bn=True, has_act=True,
                 act_type='leakyrelu'):
        super().__init__()
        assert groups == 1 and has_bn and has_act
        self.layer = nn.Sequential(
            nn.Conv2d(inplanes, planes, kernel_size, stride=stride, padding=padding, groups=groups, bias=False),
            nn.BatchNorm2d(planes), ActivationBlock(act_type=act_type, inplace=True))


class Darknet53Block(nn.Module):

    def __init__(self, inplanes, act_type='leakyrelu'):
        super().__init__()
        squeezed = int(inplanes // 2)
        self.conv = nn.Sequential(
            ConvBnActBlock(inplanes, squeezed, kernel_size=1, stride=1, padding=0, act_type=act_type),
            ConvBnActBlock(squeezed, inplanes, kernel_size=3, stride=1, padding=1, act_type=act_type))


class Darknet53(nn.Module):

    def __init__(self, act_type='leakyrelu', num_classes=1000):
        super().__init__()
        self.num_classes = num_classes
        self.act_type = act_type
        self.conv1 = ConvBnActBlock(3, 32, kernel_size=3, stride=1, padding=1, act_type=act_type)
        widths = [(32, 64, 1), (64, 128, 2), (128, 256, 8), (256, 512, 8), (512, 1024, 4)]
        for i, (cin, cout, nblocks) in enumerate(widths):
            setattr(self, f'conv{i + 2}', ConvBnActBlock(cin, cout, kernel_size=3, stride=2, padding=1, act_type=act_type))
            setattr(self, f'block{i + 1}', nn.Sequential(*[Darknet53Block(cout, act_type=act_type) for _ in range(nblocks)]))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(1024, num_classes)
        for m in self.modules():  # darknet.py:397-404
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _runtime(self):
        rt = self.__dict__.get('_rt')
        if rt is None:
            act = ACT_RELU if self.act_type == 'relu' else ACT_LEAKY
            stages = []
            for i in range(5):
                stages.append(PlainUnitRT(getattr(self, f'conv{i + 2}'), act))
                stages += [DarkBlockRT(b, act) for b in getattr(self, f'block{i + 1}')]
            rt = ResNetRT(self, has_maxpool=False, stem=ConvBN(self.conv1, act), blocks=stages)
            self.__dict__['_rt'] = rt
        return rt

    def grad_sink(self):
        return self._runtime().sink

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('this model runs on B200 kernels only; move the batch to the GPU '
                               '(no CPU fallback exists)')
        return run_network(self._runtime(), x.float(), self.training)


def darknettiny(**kwargs):
    raise NotImplementedError('darknettiny is not implemented by the B200 runtime yet (darknet53 is)')


def darknet19(**kwargs):
    raise NotImplementedError('darknet19 is not implemented by the B200 runtime yet (darknet53 is)')


def darknet53(**kwargs):
    return Darknet53(**kwargs)
