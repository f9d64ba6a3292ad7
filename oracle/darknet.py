"""Oracle: DarknetTiny / Darknet19 / Darknet53 forward as functional fp32 torch-CPU code over a state dict.

Follows SimpleAICV/classification/backbones/darknet.py:34-65 (conv -> BN -> LeakyReLU(0.1) | ReLU | SiLU),
:68-113 (Darknet19Block: alternating 3x3 / 1x1 units, optional 2x2/2 max-pool),
:116-144 (Darknet53Block: 1x1 squeeze, 3x3 expand, shortcut added after the activation),
:147-244 (DarknetTiny: six 3x3 units with 2x2/2 pools, ZeroPad2d((0,1,0,1)) + MaxPool2d(2,1), avgpool, fc),
:247-320 (Darknet19: layer1..6, biased 1x1-conv classifier, avgpool),
:323-432 (Darknet53: conv1, five stride-2 3x3 convs each followed by 1/2/8/8/4 blocks, avgpool, fc).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import math

import torch
import torch.nn.functional as F

from .convnets import BN_EPS, BN_MOMENTUM, _act_store, _keep, _RoundGrad, _w_operand

WIDTHS53 = [(32, 64, 1), (64, 128, 2), (128, 256, 8), (256, 512, 8), (512, 1024, 4)]
TINY = [(3, 16), (16, 32), (32, 64), (64, 128), (128, 256), (256, 512)]
D19 = [(32, 64, 1, True), (64, 128, 3, True), (128, 256, 3, True), (256, 512, 5, True), (512, 1024, 5, False)]


def _specs(arch):
    """(prefix, cin, cout, k, stride) of every conv+BN unit in construction order."""
    if arch == 'darknet53':
        specs = [('conv1', 3, 32, 3, 1)]
        for i, (cin, cout, nblocks) in enumerate(WIDTHS53):
            specs.append((f'conv{i + 2}', cin, cout, 3, 2))
            for b in range(nblocks):
                specs.append((f'block{i + 1}.{b}.conv.0', cout, cout // 2, 1, 1))
                specs.append((f'block{i + 1}.{b}.conv.1', cout // 2, cout, 3, 1))
        return specs
    if arch == 'darknettiny':
        return [(f'conv{i + 1}', cin, cout, 3, 1) for i, (cin, cout) in enumerate(TINY)]
    if arch == 'darknet19':
        specs = [('layer1', 3, 32, 3, 1)]
        for i, (cin, cout, n, _) in enumerate(D19):
            for j in range(n):
                a, b, k = (cin, cout, 3) if j % 2 == 0 else (cout, cin, 1)
                specs.append((f'layer{i + 2}.Darknet19Block.{j}', a, b, k, 1))
        return specs
    raise KeyError(arch)


def _default_conv(cout, cin, k):
    w = torch.empty(cout, cin, k, k)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    return w


def init_state(num_classes, seed, arch='darknet53'):
    """Seeded initial state identical to constructing the reference model after torch.manual_seed(seed):
    default Conv2d/Linear inits are drawn in construction order (conv weight, then the bias for biased
    layers), then every conv weight is re-drawn with kaiming_normal_(fan_out) in modules() order."""
    torch.manual_seed(seed)
    sd = {}
    for prefix, cin, cout, k, _ in _specs(arch):
        sd[f'{prefix}.layer.0.weight'] = _default_conv(cout, cin, k)
        sd[f'{prefix}.layer.1.weight'] = torch.ones(cout)
        sd[f'{prefix}.layer.1.bias'] = torch.zeros(cout)
        sd[f'{prefix}.layer.1.running_mean'] = torch.zeros(cout)
        sd[f'{prefix}.layer.1.running_var'] = torch.ones(cout)
        sd[f'{prefix}.layer.1.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)
    convs = [f'{p}.layer.0.weight' for p, *_ in _specs(arch)]
    if arch == 'darknet19':
        sd['layer7.layer.0.weight'] = _default_conv(num_classes, 1024, 1)
        bound = 1 / math.sqrt(1024)
        sd['layer7.layer.0.bias'] = torch.empty(num_classes).uniform_(-bound, bound)
        convs.append('layer7.layer.0.weight')
    else:
        feat = 1024 if arch == 'darknet53' else 512
        fw = torch.empty(num_classes, feat)
        torch.nn.init.kaiming_uniform_(fw, a=math.sqrt(5))
        bound = 1 / math.sqrt(feat)
        sd['fc.weight'], sd['fc.bias'] = fw, torch.empty(num_classes).uniform_(-bound, bound)
    for name in convs:
        torch.nn.init.kaiming_normal_(sd[name], mode='fan_out', nonlinearity='relu')
    return sd


def _act(y, act_type):
    if act_type == 'leakyrelu':
        return F.leaky_relu(y, 0.1)
    return F.relu(y) if act_type == 'relu' else F.silu(y)


def _cba(sd, prefix, x, k, stride, training, emu, act_type='leakyrelu'):
    y = F.conv2d(x, _w_operand(sd[f'{prefix}.layer.0.weight'], emu), None, stride, k // 2)
    y = _act_store(y, emu)
    bn = f'{prefix}.layer.1'
    y = F.batch_norm(y, sd[f'{bn}.running_mean'], sd[f'{bn}.running_var'], sd[f'{bn}.weight'], sd[f'{bn}.bias'],
                     training, BN_MOMENTUM, BN_EPS)
    if training:
        sd[f'{bn}.num_batches_tracked'] += 1
    return _act(y, act_type)


def forward(sd, x, training=True, emulate_bf16=False, trace=None, arch='darknet53', act_type='leakyrelu'):
    emu = emulate_bf16
    if emu:
        x = x.bfloat16().float()
    st = lambda t: _act_store(t, emu)  # noqa: E731
    idx = 0

    def stage(t):
        nonlocal idx
        t = _keep(trace, f'block{idx}_out', t)
        idx += 1
        return t

    if arch == 'darknet53':
        x = _keep(trace, 'stem_out', st(_cba(sd, 'conv1', x, 3, 1, training, emu, act_type)))
        for i, (_, cout, nblocks) in enumerate(WIDTHS53):
            x = stage(st(_cba(sd, f'conv{i + 2}', x, 3, 2, training, emu, act_type)))
            for b in range(nblocks):
                p = f'block{i + 1}.{b}.conv'
                t = st(_cba(sd, f'{p}.0', x, 1, 1, training, emu, act_type))
                x = stage(st(_cba(sd, f'{p}.1', t, 3, 1, training, emu, act_type) + x))
    elif arch == 'darknettiny':
        x = _keep(trace, 'stem_out', st(_cba(sd, 'conv1', x, 3, 1, training, emu, act_type)))
        x = stage(F.max_pool2d(x, 2, 2))
        for i in range(2, 7):
            x = stage(st(_cba(sd, f'conv{i}', x, 3, 1, training, emu, act_type)))
            if i < 6:
                x = stage(F.max_pool2d(x, 2, 2))
        x = stage(F.max_pool2d(F.pad(x, (0, 1, 0, 1)), 2, 1))
    elif arch == 'darknet19':
        x = _keep(trace, 'stem_out', st(_cba(sd, 'layer1', x, 3, 1, training, emu, act_type)))
        x = stage(F.max_pool2d(x, 2, 2))
        for i, (_, _, n, pool) in enumerate(D19):
            for j in range(n):
                x = stage(st(_cba(sd, f'layer{i + 2}.Darknet19Block.{j}', x, 3 if j % 2 == 0 else 1, 1, training, emu, act_type)))
            if pool:
                x = stage(F.max_pool2d(x, 2, 2))
        z = F.conv2d(x, _w_operand(sd['layer7.layer.0.weight'], emu))
        if emu:
            z = _RoundGrad.apply(z)
        z = st(z + sd['layer7.layer.0.bias'].view(1, -1, 1, 1))
        return _keep(trace, 'logits', st(F.adaptive_avg_pool2d(z, (1, 1)).flatten(1)))
    else:
        raise KeyError(arch)
    x = st(F.adaptive_avg_pool2d(x, (1, 1)).flatten(1))
    z = F.linear(x, _w_operand(sd['fc.weight'], emu))
    if emu:
        z = _RoundGrad.apply(z)
    return _keep(trace, 'logits', z + sd['fc.bias'])


def loss_and_grads(sd, x, labels, emulate_bf16=False, trace=None, arch='darknet53', act_type='leakyrelu'):
    from .convnets import param_names
    from .train_step import ce_loss
    names = param_names(sd)
    for n in names:
        sd[n].requires_grad_(True)
        sd[n].grad = None
    logits = forward(sd, x, True, emulate_bf16, trace, arch, act_type)
    loss = ce_loss(logits, labels)
    loss.backward()
    grads = {n: sd[n].grad.detach().clone() for n in names}
    for n in names:
        sd[n].requires_grad_(False)
        sd[n].grad = None
    return logits.detach(), loss.detach(), grads
