"""Oracle: SAM ViT image encoder forward as functional fp32 torch-CPU code over a state dict.

Follows SimpleAICV/interactive_segmentation/models/segment_anything/image_encoder.py:8-29 (PatchEmbed: 16x16/16 conv,
NHWC tokens), :32-79 (window partition / unpartition with zero padding), :82-144 (get_rel_pos without interpolation,
add_decomposed_rel_pos), :147-184 (Attention: packed qkv, (q*scale) k^T + rel_h + rel_w, softmax, @v, proj),
:187-198 (MLPBlock, exact GELU), :201-239 (Block: pre-LN, windowed or global attention), :242-256 (LayerNorm2d),
:313-331 (ViTImageEncoder.forward: pos_embed add, blocks, neck).  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import math

import torch
import torch.nn.functional as F

from .convnets import _keep, _RoundBoth, _RoundGrad, _RoundValue
from .vit import _linear_default

LN_EPS = 1e-6


def init_state(seed, image_size, patch_size, embedding_planes, block_nums, head_nums, mlp_ratio=4, out_planes=256, window_size=0,
               global_attn_indexes=()):
    """Seeded initial state identical to constructing the reference ViTImageEncoder after torch.manual_seed(seed):
    default Conv2d / Linear inits drawn in construction order; pos_embed and rel_pos tables start at zero."""
    torch.manual_seed(seed)
    c, grid = embedding_planes, image_size // patch_size
    hd = c // head_nums
    sd = {'pos_embed': torch.zeros(1, grid, grid, c)}
    w = torch.empty(c, 3, patch_size, patch_size)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1 / math.sqrt(3 * patch_size * patch_size)
    sd['patch_embed.proj.weight'], sd['patch_embed.proj.bias'] = w, torch.empty(c).uniform_(-bound, bound)
    for i in range(block_nums):
        p = f'blocks.{i}'
        s = grid if (window_size == 0 or i in global_attn_indexes) else window_size
        sd[f'{p}.norm1.weight'], sd[f'{p}.norm1.bias'] = torch.ones(c), torch.zeros(c)
        sd[f'{p}.attn.rel_pos_h'], sd[f'{p}.attn.rel_pos_w'] = torch.zeros(2 * s - 1, hd), torch.zeros(2 * s - 1, hd)
        sd[f'{p}.attn.qkv.weight'], sd[f'{p}.attn.qkv.bias'] = _linear_default(3 * c, c)
        sd[f'{p}.attn.proj.weight'], sd[f'{p}.attn.proj.bias'] = _linear_default(c, c)
        sd[f'{p}.norm2.weight'], sd[f'{p}.norm2.bias'] = torch.ones(c), torch.zeros(c)
        sd[f'{p}.mlp.lin1.weight'], sd[f'{p}.mlp.lin1.bias'] = _linear_default(int(c * mlp_ratio), c)
        sd[f'{p}.mlp.lin2.weight'], sd[f'{p}.mlp.lin2.bias'] = _linear_default(c, int(c * mlp_ratio))
    w0 = torch.empty(out_planes, c, 1, 1)
    torch.nn.init.kaiming_uniform_(w0, a=math.sqrt(5))
    sd['neck.0.weight'] = w0
    sd['neck.1.weight'], sd['neck.1.bias'] = torch.ones(out_planes), torch.zeros(out_planes)
    w2 = torch.empty(out_planes, out_planes, 3, 3)
    torch.nn.init.kaiming_uniform_(w2, a=math.sqrt(5))
    sd['neck.2.weight'] = w2
    sd['neck.3.weight'], sd['neck.3.bias'] = torch.ones(out_planes), torch.zeros(out_planes)
    return sd


def window_partition(x, ws):
    b, h, w, c = x.shape
    ph, pw = (ws - h % ws) % ws, (ws - w % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    hp, wp = h + ph, w + pw
    x = x.view(b, hp // ws, ws, wp // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, c), (hp, wp)


def window_unpartition(win, ws, pad_hw, hw):
    hp, wp = pad_hw
    h, w = hw
    b = win.shape[0] // (hp * wp // ws // ws)
    x = win.view(b, hp // ws, wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(b, hp, wp, -1)
    return x[:, :h, :w, :].contiguous()


def _rel_table(size, rel_pos):
    """get_rel_pos for q_size == k_size == size and a table of 2*size-1 rows (no interpolation)."""
    assert rel_pos.shape[0] == 2 * size - 1
    idx = torch.arange(size)[:, None] - torch.arange(size)[None, :] + (size - 1)
    return rel_pos[idx]           # [size(q), size(k), hd]


def forward(sd, x, head_nums, window_size=0, global_attn_indexes=(), patch_size=16, emulate_bf16=False, trace=None):
    """x: NCHW fp32 image batch -> fp32 [B, out_planes, H, W]."""
    emu = emulate_bf16
    rb = (lambda t: _RoundBoth.apply(t)) if emu else (lambda t: t)
    rw = (lambda t: _RoundValue.apply(t)) if emu else (lambda t: t)
    rg = (lambda t: _RoundGrad.apply(t)) if emu else (lambda t: t)
    if emu:
        x = x.bfloat16().float()
    t = F.conv2d(x, rw(sd['patch_embed.proj.weight']), None, stride=patch_size)
    t = rg(t) + sd['patch_embed.proj.bias'].view(1, -1, 1, 1)
    x = t.permute(0, 2, 3, 1) + sd['pos_embed']
    x = _keep(trace, 'tokens', x)
    b, h, w, c = x.shape
    hd = c // head_nums
    scale = hd ** -0.5
    nblocks = len([k for k in sd if k.endswith('.norm1.weight')])
    for i in range(nblocks):
        p = f'blocks.{i}'
        ws = 0 if (window_size == 0 or i in global_attn_indexes) else window_size
        y = rb(F.layer_norm(x, (c,), sd[f'{p}.norm1.weight'], sd[f'{p}.norm1.bias'], LN_EPS))
        if ws > 0:
            y, pad_hw = window_partition(y, ws)
        bw, sh, sw, _ = y.shape
        qkv = rb(F.linear(y, rw(sd[f'{p}.attn.qkv.weight']), sd[f'{p}.attn.qkv.bias']))
        qkv = qkv.reshape(bw, sh * sw, 3, head_nums, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.reshape(3, bw * head_nums, sh * sw, hd).unbind(0)
        attn = rb(q * scale) @ k.transpose(-2, -1)
        rq = q.reshape(bw * head_nums, sh, sw, hd)
        rel_h = rb(torch.einsum('bhwc,hkc->bhwk', rq, rw(_rel_table(sh, sd[f'{p}.attn.rel_pos_h']))))
        rel_w = rb(torch.einsum('bhwc,wkc->bhwk', rq, rw(_rel_table(sw, sd[f'{p}.attn.rel_pos_w']))))
        attn = (attn.view(-1, sh, sw, sh, sw) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(-1, sh * sw, sh * sw)
        attn = attn.softmax(dim=-1)
        o = rb((rw(attn) @ v).view(bw, head_nums, sh, sw, hd).permute(0, 2, 3, 1, 4).reshape(bw, sh, sw, c))
        if ws > 0:
            o = window_unpartition(o, ws, pad_hw, (h, w))
        x = x + (rg(F.linear(o, rw(sd[f'{p}.attn.proj.weight']))) + sd[f'{p}.attn.proj.bias'])
        y = rb(F.layer_norm(x, (c,), sd[f'{p}.norm2.weight'], sd[f'{p}.norm2.bias'], LN_EPS))
        u = rb(F.linear(y, rw(sd[f'{p}.mlp.lin1.weight']), sd[f'{p}.mlp.lin1.bias']))
        hdn = rb(F.gelu(u))
        x = x + (rg(F.linear(hdn, rw(sd[f'{p}.mlp.lin2.weight']))) + sd[f'{p}.mlp.lin2.bias'])
        x = _keep(trace, f'block{i}_out', x)
    # neck on NHWC rows: 1x1 conv == linear, LayerNorm2d == LayerNorm over channels
    oc = sd['neck.0.weight'].shape[0]
    y1 = rg(F.linear(rw(x), rw(sd['neck.0.weight'].view(oc, c))))
    l1 = rb(F.layer_norm(y1, (oc,), sd['neck.1.weight'], sd['neck.1.bias'], LN_EPS))
    y2 = rb(F.conv2d(l1.permute(0, 3, 1, 2), rw(sd['neck.2.weight']), None, 1, 1))
    l2 = rb(F.layer_norm(y2.permute(0, 2, 3, 1), (oc,), sd['neck.3.weight'], sd['neck.3.bias'], LN_EPS))
    return _keep(trace, 'out', l2.permute(0, 3, 1, 2))


def loss_and_grads(sd, x, proj, head_nums, window_size=0, global_attn_indexes=(), patch_size=16, emulate_bf16=False, trace=None):
    """loss = mean(out * proj) with a fixed random `proj` of the output's shape (the encoder has no loss of its own:
    in the reference it feeds the mask decoder / a distillation MSE).  Returns (out, loss, grads)."""
    names = list(sd.keys())
    for n in names:
        sd[n].requires_grad_(True)
        sd[n].grad = None
    out = forward(sd, x, head_nums, window_size, global_attn_indexes, patch_size, emulate_bf16, trace)
    loss = (out.float() * proj).mean()
    loss.backward()
    grads = {n: (sd[n].grad.detach().clone() if sd[n].grad is not None else torch.zeros_like(sd[n])) for n in names}
    for n in names:
        sd[n].requires_grad_(False)
        sd[n].grad = None
    return out.detach(), loss.detach(), grads
