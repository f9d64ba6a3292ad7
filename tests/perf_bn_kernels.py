"""Micro-benchmark of the HBM-bound BatchNorm kernels at ResNet-50 bs256 shapes (CUDA events,
inputs far larger than L2 or rotated buffers).  Not a test; run on the GPU box:
    python tests/perf_bn_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simpleaicv_pytorch_training_examples_b200 import ops  # noqa: E402

SHAPES = [(802816, 64), (802816, 256), (200704, 128), (200704, 512), (50176, 256), (50176, 1024), (12544, 512), (12544, 2048)]


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for rows, c in SHAPES:
        nb = max(2, int(400e6 // (rows * c * 2)) + 1)  # rotate buffers so that L2 never holds the input
        ys = [torch.randn(rows, c, device='cuda').to(torch.bfloat16) for _ in range(nb)]
        douts = [torch.randn(rows, c, device='cuda').to(torch.bfloat16) for _ in range(nb)]
        out = torch.empty_like(ys[0])
        stats = torch.zeros(2, c, device='cuda')
        ss, saved = torch.empty(2, c, device='cuda'), torch.empty(2, c, device='cuda')
        gamma, beta = torch.ones(c, device='cuda'), torch.zeros(c, device='cuda')
        dg, db = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
        part = ops.bn_stats(ys[0])
        ops.bn_finalize(part, gamma, beta, None, None, ss, saved, rows, 1e-5, 0.1)
        it = [0]

        def nxt():
            it[0] = (it[0] + 1) % nb
            return it[0]

        mb = rows * c * 2 / 1e6
        res = {}
        if only in (None, 'stats'):
            res['stats'] = (timeit(lambda: ops.bn_stats(ys[nxt()])), 1)
        if only in (None, 'apply'):
            res['apply'] = (timeit(lambda: ops.bn_apply(ys[nxt()], ss, out, 1)), 2)
        if only in (None, 'bwd_reduce'):
            res['bwd_reduce'] = (timeit(lambda: ops.bn_bwd_reduce(douts[nxt()], None, ys[it[0]], saved, stats, 1, scale_shift=ss)), 2)
        if only in (None, 'bwd_apply'):
            res['bwd_apply'] = (timeit(lambda: ops.bn_bwd_apply(douts[nxt()], None, ys[it[0]], saved, gamma, stats, out, None, dg, db, 1, scale_shift=ss)), 3)
        print(f'rows {rows:7d} C {c:5d} ({mb:6.1f} MB): ' + '  '.join(f'{k} {ms * 1e3:7.1f} us {mb * n / ms / 1e3:6.2f} TB/s' for k, (ms, n) in res.items()), flush=True)


if __name__ == '__main__':
    main()
