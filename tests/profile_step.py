"""One training step bracketed by cudaProfilerStart/Stop for ncu (--profile-from-start off).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches_r50.csv python tests/profile_step.py --model resnet50
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='resnet50', choices=['resnet50', 'vit_base_patch16'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--fwd-only', action='store_true')
    a = ap.parse_args()
    from simpleaicv_pytorch_training_examples_b200.classification import backbones, losses
    from simpleaicv_pytorch_training_examples_b200.tools import utils as tutils
    torch.manual_seed(0)
    x = torch.randn(a.batch, 3, 224, 224, device='cuda')
    if a.model == 'resnet50':
        model = backbones.resnet50(num_classes=1000).cuda().train()
        crit, y = losses.CELoss(), torch.randint(0, 1000, (a.batch,), device='cuda')

        class Cfg:   # the optimizer bench.py builds: fused multi-tensor SGD (optim.py), 1-D parameters without weight decay
            optimizer = ('SGD', {'lr': 0.1, 'momentum': 0.9, 'global_weight_decay': False, 'weight_decay': 1e-4,
                                 'no_weight_decay_layer_name_list': []})
        opt, _ = tutils.build_optimizer(Cfg, model)
    else:
        model = backbones.vit_base_patch16(image_size=224, num_classes=1000, drop_path_prob=0.1, global_pool=True).cuda().train()
        crit = losses.OneHotLabelCELoss()
        y = torch.nn.functional.one_hot(torch.randint(0, 1000, (a.batch,), device='cuda'), 1000).float()

        class Cfg:
            optimizer = ('AdamW', {'lr': 5e-4, 'global_weight_decay': False, 'weight_decay': 0.05,
                                   'no_weight_decay_layer_name_list': ['position_encoding', 'cls_token'],
                                   'lr_layer_decay': 0.65, 'lr_layer_decay_block': model.blocks, 'block_name': 'blocks'})
        opt, _ = tutils.build_optimizer(Cfg, model)

    def step():
        loss = crit(model(x), y)
        if not a.fwd_only:
            loss.backward()
            opt.step()
            opt.zero_grad()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main()
