"""Per-shape timing of the BatchNorm slab kernels on the ResNet-50 bs256 activation shapes (GB/s by algorithmic bytes).
Run on the GPU box: python tests/profile_bn.py [--once]  (--once: a single launch of each, for ncu)."""
import sys

import torch

sys.path.insert(0, '.')
from simpleaicv_pytorch_training_examples_b200 import ops  # noqa: E402

SHAPES = [(256 * 112 * 112, 64), (256 * 56 * 56, 64), (256 * 56 * 56, 256), (256 * 28 * 28, 128), (256 * 28 * 28, 512),
          (256 * 14 * 14, 256), (256 * 14 * 14, 1024), (256 * 7 * 7, 512), (256 * 7 * 7, 2048)]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    once = '--once' in sys.argv
    dev = torch.device('cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for rows, c in SHAPES if not once else SHAPES[2:3]:
        y = torch.randn(rows, c, device=dev).bfloat16()
        dout = torch.randn(rows, c, device=dev).bfloat16()
        res = torch.randn(rows, c, device=dev).bfloat16()
        out, dy, dres = torch.empty_like(y), torch.empty_like(y), torch.empty_like(y)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        ss, saved, sums = (torch.empty(2 * c, device=dev) for _ in range(3))
        dg, db = torch.empty(c, device=dev), torch.empty(c, device=dev)
        part = ops.bn_stats(y)
        ops.bn_finalize(part, gamma, beta, rm, rv, ss, saved, rows, 1e-5, 0.1)
        nb = rows * c * 2
        cases = {
            'bn_stats': (lambda: ops.bn_stats(y), 1),
            'bn_apply relu': (lambda: ops.bn_apply(y, ss, out, 1), 2),
            'bn_apply +res relu': (lambda: ops.bn_apply(y, ss, out, 1, res), 3),
            'bn_bwd_reduce recompute-mask': (lambda: ops.bn_bwd_reduce(dout, None, y, saved, sums, 1, ss), 2),
            'bn_bwd_reduce out-mask': (lambda: ops.bn_bwd_reduce(dout, out, y, saved, sums, 1), 3),
            'bn_bwd_apply recompute-mask': (lambda: ops.bn_bwd_apply(dout, None, y, saved, gamma, sums, dy, None, dg, db, 1, False, ss), 3),
            'bn_bwd_apply out-mask +dres': (lambda: ops.bn_bwd_apply(dout, out, y, saved, gamma, sums, dy, dres, dg, db, 1), 5),
            'add_bf16': (lambda: ops.add_bf16(dy, dout), 3),
        }
        for name, (fn, passes) in cases.items():
            if once:
                fn()
                continue
            ms = timeit(fn)
            print(f'rows {rows:9d} C {c:5d} {name:32s} {ms * 1e3:8.1f} us  {passes * nb / ms / 1e6:7.0f} GB/s ({passes} tensor passes, {nb / 1e6:.0f} MB each)')
        del y, dout, res, out, dy, dres
    torch.cuda.synchronize()
    del flush


if __name__ == '__main__':
    main()
