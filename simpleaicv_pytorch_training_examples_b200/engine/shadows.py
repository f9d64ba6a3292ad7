"""Enumerates the bf16 GEMM-operand copies a runtime keeps of its fp32 weights, for the fused optimizers
(optim.FusedSGD / FusedAdamW refresh them inside the parameter update instead of one cast / re-layout launch per weight
in the next forward's prep()).

Two layouts are fused: row-major copies with the parameter's own linear index (nn.Linear [N][K] incl. the zero-padded
class rows, 1x1-conv classifiers) and the tap-major conv operand [K][R*S*Cp] of saicv_prep_conv_weight order 0.  The
3-channel stems / patch embeddings ([K][C*R*S8] padded columns) stay on their prep kernel: the optimizer bumps those
parameters' version so prep() re-casts them."""
from .convnet import ConvBN


def _walk(obj, seen, out, depth=0):
    if id(obj) in seen or depth > 6:
        return
    seen.add(id(obj))
    if isinstance(obj, ConvBN):
        out.append(obj)
    elif hasattr(obj, 'w_bf16') and (hasattr(obj, 'mod') or hasattr(obj, 'fc') or hasattr(obj, 'conv')):
        out.append(obj)
    if isinstance(obj, (list, tuple)):
        for o in obj:
            _walk(o, seen, out, depth + 1)
        return
    d = getattr(obj, '__dict__', None)
    if d is None or type(obj).__module__.startswith('torch'):
        return
    for v in d.values():
        if isinstance(v, (list, tuple)) or (hasattr(v, '__dict__') and type(v).__module__.startswith(__package__)):
            _walk(v, seen, out, depth + 1)


def collect(rt):
    """[(parameter, bf16 shadow tensor, conv layout (c, taps, cp, kpad) or None, owner object holding it as `w_bf16`)] of the
    runtime `rt` (after rt.prep())."""
    rt.prep()
    units, res = [], []
    _walk(rt, set(), units)
    for u in units:
        if isinstance(u, ConvBN):
            if u.is_stem or u.w_bf16 is None:
                continue
            res.append((u.conv.weight, u.w_bf16, (u.c, u.r * u.s, u.cp, u.kpad), u))
            continue
        mod = getattr(u, 'mod', None) or getattr(u, 'fc', None) or getattr(u, 'conv', None)
        w = getattr(mod, 'weight', None)
        plain = w is not None and (w.dim() == 2 or (w.dim() == 4 and w.shape[2] == 1 and w.shape[3] == 1))
        if not plain or u.w_bf16 is None or u.w_bf16.dim() != 2 or u.w_bf16.shape[1] != w.shape[1] or \
                u.w_bf16.shape[0] < w.shape[0] or not u.w_bf16.is_contiguous():
            continue
        res.append((w, u.w_bf16, None, u))
    return res
