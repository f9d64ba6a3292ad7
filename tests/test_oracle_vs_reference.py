"""Live comparison oracle <-> reference; only where /root/reference exists (build container)."""
import os
import sys

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'SimpleAICV')),
                                reason='reference checkout not present (GPU box)')


def _ref_backbones():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from SimpleAICV.classification import backbones
    return backbones


@pytest.mark.parametrize('arch,nc,shape', [('resnet34cifar', 100, (2, 3, 32, 32)),
                                           ('resnet50cifar', 10, (2, 3, 32, 32)),
                                           ('resnet18', 1000, (2, 3, 64, 64))])
def test_oracle_matches_reference_bitwise(arch, nc, shape):
    from oracle import convnets, train_step
    torch.manual_seed(5)
    ref = _ref_backbones().__dict__[arch](num_classes=nc)
    sd = convnets.init_state(arch, nc, 5)
    rs = ref.state_dict()
    assert list(rs.keys()) == list(sd.keys())
    assert all(torch.equal(rs[k], sd[k]) for k in rs)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(*shape, generator=g)
    y = torch.randint(0, nc, (shape[0],), generator=g)
    ref.train()
    out = ref(x)
    loss = torch.nn.functional.cross_entropy(out.float(), y)
    loss.backward()
    lo, ls, gr = train_step.loss_and_grads(sd, x, y, arch)
    torch.testing.assert_close(lo, out.detach(), rtol=1e-5, atol=1e-5)
    for n, p in ref.named_parameters():
        torch.testing.assert_close(gr[n], p.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('arch', ['resnet18', 'resnet50', 'resnet101', 'resnet18cifar', 'resnet152cifar'])
def test_b200_constructors_match_reference_state_dict(arch):
    """Drop-in contract: same keys, shapes and seeded values as the reference."""
    from simpleaicv_pytorch_training_examples_b200.classification import backbones as mine
    torch.manual_seed(0)
    ref = _ref_backbones().__dict__[arch](num_classes=100)
    torch.manual_seed(0)
    m = mine.__dict__[arch](num_classes=100)
    rs, ms = ref.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    assert all(torch.equal(rs[k], ms[k]) for k in rs)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in m.named_parameters()]
