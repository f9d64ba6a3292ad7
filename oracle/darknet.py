"""Oracle: Darknet-53 forward as functional fp32 torch-CPU code over a state dict.

Follows SimpleAICV/classification/backbones/darknet.py:34-65 (conv -> BN -> LeakyReLU(0.1)),
:116-144 (Darknet53Block: 1x1 squeeze, 3x3 expand, shortcut added after the activation),
:323-432 (Darknet53: conv1, five stride-2 3x3 convs each followed by 1/2/8/8/4 blocks, global
average pool, fc).  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import math

import torch
import torch.nn.functional as F

from .convnets import BN_EPS, BN_MOMENTUM, _act_store, _keep, _RoundGrad, _w_operand

WIDTHS = [(32, 64, 1), (64, 128, 2), (128, 256, 8), (256, 512, 8), (512, 1024, 4)]


def _specs():
    """(prefix, cin, cout, k, stride) of every ConvBnActBlock in construction order."""
    specs = [('conv1', 3, 32, 3, 1)]
    for i, (cin, cout, nblocks) in enumerate(WIDTHS):
        specs.append((f'conv{i + 2}', cin, cout, 3, 2))
        for b in range(nblocks):
            specs.append((f'block{i + 1}.{b}.conv.0', cout, cout // 2, 1, 1))
            specs.append((f'block{i + 1}.{b}.conv.1', cout // 2, cout, 3, 1))
    return specs


def init_state(num_classes, seed):
    """Seeded initial state identical to constructing the reference Darknet53 after
    torch.manual_seed(seed) (default Conv2d/Linear inits drawn in construction order, then
    darknet.py:397-404 re-draws convs with kaiming_normal_(fan_out))."""
    torch.manual_seed(seed)
    sd = {}
    for prefix, cin, cout, k, _ in _specs():
        w = torch.empty(cout, cin, k, k)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        sd[f'{prefix}.layer.0.weight'] = w
        sd[f'{prefix}.layer.1.weight'] = torch.ones(cout)
        sd[f'{prefix}.layer.1.bias'] = torch.zeros(cout)
        sd[f'{prefix}.layer.1.running_mean'] = torch.zeros(cout)
        sd[f'{prefix}.layer.1.running_var'] = torch.ones(cout)
        sd[f'{prefix}.layer.1.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    fw = torch.empty(num_classes, 1024)
    torch.nn.init.kaiming_uniform_(fw, a=math.sqrt(5))
    bound = 1 / math.sqrt(1024)
    sd['fc.weight'], sd['fc.bias'] = fw, torch.empty(num_classes).uniform_(-bound, bound)
    for prefix, *_ in _specs():
        torch.nn.init.kaiming_normal_(sd[f'{prefix}.layer.0.weight'], mode='fan_out', nonlinearity='relu')
    return sd


def _cba(sd, prefix, x, k, stride, training, emu):
    y = F.conv2d(x, _w_operand(sd[f'{prefix}.layer.0.weight'], emu), None, stride, k // 2)
    y = _act_store(y, emu)
    bn = f'{prefix}.layer.1'
    y = F.batch_norm(y, sd[f'{bn}.running_mean'], sd[f'{bn}.running_var'], sd[f'{bn}.weight'], sd[f'{bn}.bias'],
                     training, BN_MOMENTUM, BN_EPS)
    if training:
        sd[f'{bn}.num_batches_tracked'] += 1
    return F.leaky_relu(y, 0.1)


def forward(sd, x, training=True, emulate_bf16=False, trace=None):
    emu = emulate_bf16
    if emu:
        x = x.bfloat16().float()
    x = _keep(trace, 'stem_out', _act_store(_cba(sd, 'conv1', x, 3, 1, training, emu), emu))
    idx = 0
    for i, (_, cout, nblocks) in enumerate(WIDTHS):
        x = _keep(trace, f'block{idx}_out', _act_store(_cba(sd, f'conv{i + 2}', x, 3, 2, training, emu), emu))
        idx += 1
        for b in range(nblocks):
            p = f'block{i + 1}.{b}.conv'
            t = _act_store(_cba(sd, f'{p}.0', x, 1, 1, training, emu), emu)
            x = _keep(trace, f'block{idx}_out', _act_store(_cba(sd, f'{p}.1', t, 3, 1, training, emu) + x, emu))
            idx += 1
    x = _act_store(F.adaptive_avg_pool2d(x, (1, 1)).flatten(1), emu)
    z = F.linear(x, _w_operand(sd['fc.weight'], emu))
    if emu:
        z = _RoundGrad.apply(z)
    return _keep(trace, 'logits', z + sd['fc.bias'])


def loss_and_grads(sd, x, labels, emulate_bf16=False, trace=None):
    from .convnets import param_names
    from .train_step import ce_loss
    names = param_names(sd)
    for n in names:
        sd[n].requires_grad_(True)
        sd[n].grad = None
    logits = forward(sd, x, True, emulate_bf16, trace)
    loss = ce_loss(logits, labels)
    loss.backward()
    grads = {n: sd[n].grad.detach().clone() for n in names}
    for n in names:
        sd[n].requires_grad_(False)
        sd[n].grad = None
    return logits.detach(), loss.detach(), grads
