"""Darknet-53: oracle vs reference (when present) and constructor state_dict parity, on CPU."""
import os
import sys

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'SimpleAICV')), reason='reference checkout not present')


def test_darknet53_oracle_and_constructor_match_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from SimpleAICV.classification import backbones as refb
    from oracle import darknet
    from simpleaicv_pytorch_training_examples_b200.classification import backbones as mine
    torch.manual_seed(2)
    ref = refb.darknet53(num_classes=10)
    torch.manual_seed(2)
    m = mine.darknet53(num_classes=10)
    sd = darknet.init_state(10, 2)
    rs, ms = ref.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys()) == list(sd.keys())
    assert all(torch.equal(rs[k], ms[k]) and torch.equal(rs[k], sd[k]) for k in rs)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 64, 64, generator=g)
    y = torch.randint(0, 10, (2,), generator=g)
    ref.train()
    out = ref(x)
    torch.nn.functional.cross_entropy(out.float(), y).backward()
    lo, _, gr = darknet.loss_and_grads(sd, x, y)
    torch.testing.assert_close(lo, out.detach(), rtol=1e-5, atol=1e-5)
    for n, p in ref.named_parameters():
        torch.testing.assert_close(gr[n], p.grad, rtol=1e-4, atol=1e-6)
    with pytest.raises(NotImplementedError):
        mine.darknet19()
