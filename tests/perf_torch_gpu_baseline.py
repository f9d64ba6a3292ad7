"""Context number (NOT a bench.py arm): the reference's GPU path on this box, i.e. its model code
(restated by oracle/convnets.py, the same torch ops) executed by stock PyTorch CUDA kernels under
autocast(bf16) + autograd + torch.optim.SGD, with the per-step host syncs of
tools/scripts.py:141-270 (isinf/isnan checks, loss == 0 test, .item()).  This is the
"torch-DDP on the same box" number BASELINE.md 2b asks for at N = 1.

    python tests/perf_torch_gpu_baseline.py [--batch 256] [--steps 10] [--channels-last] [--benchmark]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import convnets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--benchmark', action='store_true', help='cudnn.benchmark=True (the reference ships False + deterministic)')
    ap.add_argument('--channels-last', action='store_true')
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = a.benchmark
    torch.backends.cudnn.deterministic = not a.benchmark  # tools/utils.py:106-107
    dev = torch.device('cuda')
    sd = {k: v.to(dev) for k, v in convnets.init_state('resnet50', 1000, 0).items()}
    names = convnets.param_names(sd)
    if a.channels_last:
        for n in names:
            if sd[n].ndim == 4:
                sd[n] = sd[n].contiguous(memory_format=torch.channels_last)
    for n in names:
        sd[n].requires_grad_(True)
    decay = [sd[n] for n in names if sd[n].ndim > 1]
    nodecay = [sd[n] for n in names if sd[n].ndim <= 1]
    opt = torch.optim.SGD([{'params': decay, 'weight_decay': 1e-4}, {'params': nodecay, 'weight_decay': 0.}], lr=0.1, momentum=0.9)
    x = torch.randn(a.batch, 3, 224, 224, device=dev)
    if a.channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (a.batch,), device=dev)

    def step():
        if torch.any(torch.isinf(x)) or torch.any(torch.isnan(x)):
            return
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = convnets.forward(sd, x, 'resnet50', training=True)
            loss = torch.nn.functional.cross_entropy(out.float(), y)
        if loss == 0. or torch.any(torch.isinf(loss)) or torch.any(torch.isnan(loss)):
            return
        loss.backward()
        opt.step()
        opt.zero_grad()
        return loss.item()

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.steps
    print(json.dumps({'what': 'stock torch (reference GPU path) ResNet-50 bf16 autocast', 'batch': a.batch,
                      'cudnn_benchmark': a.benchmark, 'channels_last': a.channels_last, 'ms_per_step': ms,
                      'images_per_sec': a.batch / ms * 1e3, 'torch': torch.__version__}))


if __name__ == '__main__':
    main()
