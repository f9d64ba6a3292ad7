"""The oracle replays the committed reference outputs (tests/golden/*.pt, produced by
tests/golden/make_golden.py from the reference itself).  Runs anywhere: CPU only."""
import glob
import os

import pytest
import torch

from oracle import convnets, train_step

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_outputs(path):
    fix = torch.load(path, weights_only=False)
    torch.set_num_threads(1)
    sd = convnets.init_state(fix['arch'], fix['num_classes'], fix['seed'])
    logits, loss, grads = train_step.loss_and_grads(sd, fix['x'], fix['y'], fix['arch'])
    # fp32 vs fp32 on CPU: rtol 1e-5; identical torch builds give bit equality
    torch.testing.assert_close(logits, fix['logits'], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(loss, fix['loss'], rtol=1e-5, atol=1e-6)
    for n, g in grads.items():
        assert abs(g.norm().item() - fix['grad_norm'][n]) <= 1e-4 * max(1.0, fix['grad_norm'][n]), n
        torch.testing.assert_close(g.flatten()[:4], fix['grad_head'][n], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(sd['conv1.layer.1.running_mean'], fix['running_mean_conv1'], rtol=1e-5, atol=1e-6)
    with torch.no_grad():
        ev = convnets.forward(sd, fix['x'], fix['arch'], training=False)
    torch.testing.assert_close(ev, fix['eval_logits'], rtol=1e-4, atol=1e-4)


def test_golden_fixtures_present():
    assert len(GOLDEN) >= 3
