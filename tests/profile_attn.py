"""Attention kernels at the ViT-B/16 bs256 shape (and optionally others) for ncu / quick timing."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleaicv_pytorch_training_examples_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--b', type=int, default=256)
ap.add_argument('--l', type=int, default=197)
ap.add_argument('--h', type=int, default=12)
ap.add_argument('--d', type=int, default=64)
ap.add_argument('--iters', type=int, default=3)
a = ap.parse_args()
g = torch.Generator(device='cuda').manual_seed(0)
qkv = torch.randn(a.b, a.l, 3, a.h, a.d, device='cuda', generator=g).to(torch.bfloat16)
dout = torch.randn(a.b * a.l, a.h * a.d, device='cuda', generator=g).to(torch.bfloat16)
scale = a.d ** -0.5
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, a.b, a.l, a.h, a.d, scale)
    ops.attention_bwd(qkv, out, dout, lse, a.b, a.l, a.h, a.d, scale)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record()
for _ in range(a.iters):
    out, lse = ops.attention_fwd(qkv, a.b, a.l, a.h, a.d, scale)
ev[1].record()
for _ in range(a.iters):
    ops.attention_bwd(qkv, out, dout, lse, a.b, a.l, a.h, a.d, scale)
ev[2].record()
torch.cuda.synchronize()
print(f'attention B{a.b} L{a.l} H{a.h} D{a.d}: fwd {ev[0].elapsed_time(ev[1]) / a.iters * 1e3:.1f} us, '
      f'bwd {ev[1].elapsed_time(ev[2]) / a.iters * 1e3:.1f} us')
