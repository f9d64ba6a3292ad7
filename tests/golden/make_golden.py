"""Generates tests/golden/*.pt by running the REFERENCE itself (imported from /root/reference)
in the build container.  The fixtures are small: the seeded input batch, the logits, the loss and
per-parameter gradient digests (L2 norm + first 4 values); the weights are not stored because
the oracle reproduces the reference's seeded initialisation exactly (checked here, bit for bit).

    python tests/golden/make_golden.py            # rewrites the fixtures
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

CASES = [
    # (tag, arch, num_classes, batch shape, seed)
    ('resnet18cifar_c100_b8', 'resnet18cifar', 100, (8, 3, 32, 32), 0),  # BASELINE config #1 shape
    ('resnet50_c1000_b4_64px', 'resnet50', 1000, (4, 3, 64, 64), 0),
    ('resnet18_c10_b2_64px', 'resnet18', 10, (2, 3, 64, 64), 3),
]


def main():
    from SimpleAICV.classification import backbones
    from SimpleAICV.classification.losses import CELoss
    from oracle import convnets
    torch.set_num_threads(1)  # fixed reduction order for the recorded numbers
    for tag, arch, nc, shape, seed in CASES:
        torch.manual_seed(seed)
        model = backbones.__dict__[arch](num_classes=nc)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        osd = convnets.init_state(arch, nc, seed)
        assert all(torch.equal(sd0[k], osd[k]) for k in sd0), 'oracle init != reference init'
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.randn(*shape, generator=g)
        y = torch.randint(0, nc, (shape[0],), generator=g)
        model.train()
        logits = model(x)
        loss = CELoss()(logits, y)
        loss.backward()
        fix = {
            'arch': arch, 'num_classes': nc, 'seed': seed, 'x': x, 'y': y,
            'logits': logits.detach(), 'loss': loss.detach(),
            'grad_norm': {n: p.grad.norm().item() for n, p in model.named_parameters()},
            'grad_head': {n: p.grad.flatten()[:4].clone() for n, p in model.named_parameters()},
            'running_mean_conv1': model.state_dict()['conv1.layer.1.running_mean'].clone(),
            'torch_version': torch.__version__,
            'reference_commit': '14b1826',
        }
        model.eval()
        with torch.no_grad():
            fix['eval_logits'] = model(x).clone()
        torch.save(fix, os.path.join(HERE, tag + '.pt'))
        print(tag, 'loss', float(loss), os.path.getsize(os.path.join(HERE, tag + '.pt')), 'bytes')


if __name__ == '__main__':
    main()
