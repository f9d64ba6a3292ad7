"""Generates tests/golden/detr_loss_b3.ptl by running the REFERENCE DETRLoss (SimpleAICV/detection/losses.py:843-1095)
on seeded predictions / annotations: the 18 loss terms, the Hungarian indices and the gradients of their sum.

    python tests/golden/make_detr_loss_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from baseline import ref_import  # noqa: E402


def make_inputs(seed=0, counts=(3, 7, 1), queries=100, classes=80):
    g = torch.Generator().manual_seed(seed)
    b = len(counts)
    cls = torch.randn(6, b, queries, classes + 1, generator=g)
    reg = torch.rand(6, b, queries, 4, generator=g)
    ann = -torch.ones(b, max(counts), 5)
    for i, n in enumerate(counts):
        ann[i, :n, :2] = torch.rand(n, 2, generator=g) * 0.6 + 0.2
        ann[i, :n, 2:4] = torch.rand(n, 2, generator=g) * 0.3 + 0.05
        ann[i, :n, 4] = torch.randint(0, classes, (n,), generator=g).float()
    return cls, reg, ann


def main():
    ref = ref_import.module('SimpleAICV.detection.losses').DETRLoss()
    cls, reg, ann = make_inputs()
    cls.requires_grad_(True)
    reg.requires_grad_(True)
    losses = ref([cls, reg], ann)
    sum(losses.values()).backward()
    idx = ref.get_matched_pred_target_idxs(cls[-1].detach(), reg[-1].detach().clamp(1e-4, 1 - 1e-4), ann)
    torch.save({'losses': {k: v.detach() for k, v in losses.items()}, 'indices': idx, 'dcls_norm': cls.grad.norm(), 'dreg_norm': reg.grad.norm(),
                'dcls_head': cls.grad.flatten()[:8].clone(), 'dreg_sum': reg.grad.sum(), 'torch_version': torch.__version__},
               os.path.join(HERE, 'detr_loss_b3.ptl'))
    print({k: float(v) for k, v in list(losses.items())[:3]})


if __name__ == '__main__':
    main()
