"""Generates tests/golden/sam_loss_b3.ptl by running the REFERENCE SAMLoss (SimpleAICV/interactive_segmentation/losses.py:11-198)
on seeded mask logits / IoU predictions / targets (3 prompt iterations, batch 3, 4 masks, 96 x 96): the three weighted loss
terms and the gradients of their sum, for supervise_all_iou True and False.

    python tests/golden/make_sam_loss_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def make_inputs(seed=0, iters=3, b=3, m=4, h=96, w=96):
    g = torch.Generator().manual_seed(seed)
    targets = (torch.rand(b, 1, h, w, generator=g) > 0.55).float()
    masks = [torch.randn(b, m, h, w, generator=g) * 2 + (targets * 2 - 1) * 0.5 for _ in range(iters)]
    ious = [torch.rand(b, m, generator=g) for _ in range(iters)]
    return masks, ious, targets


def main():
    from baseline import ref_import
    ref_cls = ref_import.module('SimpleAICV.interactive_segmentation.losses').SAMLoss
    out = {}
    for supervise_all in (True, False):
        masks, ious, targets = make_inputs()
        masks = [m.requires_grad_(True) for m in masks]
        ious = [i.requires_grad_(True) for i in ious]
        losses = ref_cls(supervise_all_iou=supervise_all)((masks, ious), targets)
        sum(losses.values()).backward()
        out[supervise_all] = {'losses': {k: v.detach() for k, v in losses.items()}, 'dmask_norm': [m.grad.norm() for m in masks],
                              'diou': [i.grad.clone() for i in ious]}
        print(supervise_all, {k: float(v) for k, v in losses.items()})
    out['torch_version'] = torch.__version__
    torch.save(out, os.path.join(HERE, 'sam_loss_b3.ptl'))


if __name__ == '__main__':
    main()
