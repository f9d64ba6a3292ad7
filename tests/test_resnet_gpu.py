"""Whole-network parity on the GPU: the B200 runtime (through the C ABI) against the CPU oracle
on the same seeded weights and inputs, and against the committed reference outputs.

What is compared, and why two oracles.  The kernels store activations / activation gradients in
bf16 and accumulate in fp32 (like the reference under autocast).  On randomly initialised
ResNets that storage precision alone moves early-layer gradients by ~0.3 relative L2 — the
reference's own autocast(bf16) run is that far from its fp32 run (DESIGN.md "Parity") — so a
fixed tight tolerance against fp32 is meaningless.  Therefore:

 (1) tight check against the oracle evaluated WITH the same bf16 storage points
     (oracle.convnets.forward(emulate_bf16=True)): logits atol 2e-2 + rtol 2e-2, loss rtol 5e-3,
     every parameter gradient relative L2 <= 4e-2 and cosine >= 0.999, BN running mean rtol 1e-2;
 (2) noise-bounded check against the fp32 oracle (= the reference): our distance to fp32 may not
     exceed 1.5x the distance of the bf16-storage oracle to fp32 (+ 1e-2), for logits and for
     every parameter gradient.
"""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()


def _run_mine(arch, nc, seed, x, y):
    from simpleaicv_pytorch_training_examples_b200.classification import backbones, losses
    torch.manual_seed(seed)
    model = backbones.__dict__[arch](num_classes=nc).cuda()
    model.train()
    logits = model(x.cuda())
    loss = losses.CELoss()(logits, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters()}
    return model, logits.detach().float().cpu(), float(loss.detach()), grads


@pytest.mark.parametrize('arch,nc,shape', [('resnet18cifar', 100, (16, 3, 32, 32)),
                                           ('resnet50', 1000, (8, 3, 64, 64)),
                                           ('resnet50cifar', 100, (4, 3, 32, 32)),
                                           ('resnet18', 1000, (4, 3, 96, 96))])
def test_resnet_step_matches_oracle(arch, nc, shape):
    from oracle import convnets, train_step
    seed = 0
    g = torch.Generator().manual_seed(77)
    x = torch.randn(*shape, generator=g)
    y = torch.randint(0, nc, (shape[0],), generator=g)
    sd32 = convnets.init_state(arch, nc, seed)
    l32, ls32, g32 = train_step.loss_and_grads(sd32, x, y, arch)
    sde = convnets.init_state(arch, nc, seed)
    le, lse, ge = train_step.loss_and_grads(sde, x, y, arch, emulate_bf16=True)
    model, logits, loss, grads = _run_mine(arch, nc, seed, x, y)
    # (1) against the bf16-storage oracle
    err = (logits - le).abs()
    assert (err <= 2e-2 + 2e-2 * le.abs()).all(), f'logits max err vs bf16-storage oracle {err.max().item():.4g}'
    assert abs(loss - float(lse)) <= 5e-3 * abs(float(lse)), (loss, float(lse))
    worst = (0.0, None)
    for n, ref in ge.items():
        assert n in grads, f'missing grad {n}'
        rl, cs = _rel_l2(grads[n], ref), _cos(grads[n], ref)
        worst = max(worst, (rl, n))
        assert rl <= 4e-2 and cs >= 0.999, f'{n}: rel L2 {rl:.4g} cos {cs:.5f} vs bf16-storage oracle'
    rm = model.state_dict()['conv1.layer.1.running_mean'].cpu()
    torch.testing.assert_close(rm, sde['conv1.layer.1.running_mean'], rtol=1e-2, atol=1e-3)
    # (2) against the fp32 oracle, bounded by the bf16 storage noise itself
    assert _rel_l2(logits, l32) <= 1.5 * _rel_l2(le, l32) + 1e-2
    worst32 = (0.0, None)
    for n, ref in g32.items():
        mine, emu = _rel_l2(grads[n], ref), _rel_l2(ge[n], ref)
        worst32 = max(worst32, (mine, n))
        assert mine <= 1.5 * emu + 1e-2, f'{n}: rel L2 to fp32 {mine:.4g} vs bf16-storage noise {emu:.4g}'
    print(f'{arch}: vs bf16-storage oracle: logits max err {err.max().item():.4g}, worst grad rel L2 {worst}; '
          f'vs fp32: logits rel L2 {_rel_l2(logits, l32):.4g} (storage noise {_rel_l2(le, l32):.4g}), worst grad {worst32}')


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), 'golden', '*.pt')))


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_resnet_matches_reference_golden(path):
    """Against outputs recorded from the reference itself (tests/golden/make_golden.py)."""
    fix = torch.load(path, weights_only=False)
    model, logits, loss, grads = _run_mine(fix['arch'], fix['num_classes'], fix['seed'], fix['x'], fix['y'])
    # the recorded reference numbers are fp32: bound our distance by the bf16 storage noise
    from oracle import convnets, train_step
    sde = convnets.init_state(fix['arch'], fix['num_classes'], fix['seed'])
    le, lse, ge = train_step.loss_and_grads(sde, fix['x'], fix['y'], fix['arch'], emulate_bf16=True)
    assert _rel_l2(logits, fix['logits']) <= 1.5 * _rel_l2(le, fix['logits']) + 1e-2
    assert abs(loss - float(fix['loss'])) <= 1.5 * abs(float(lse) - float(fix['loss'])) + 1e-2 * abs(float(fix['loss']))
    for n, gn in fix['grad_norm'].items():
        noise = abs(ge[n].norm().item() - gn)
        assert abs(grads[n].norm().item() - gn) <= 1.5 * noise + 3e-2 * max(gn, 1e-6), (n, grads[n].norm().item(), gn)
    # eval mode (running statistics) through the same kernels
    model.eval()
    with torch.no_grad():
        ev = model(fix['x'].cuda()).float().cpu()
    e2 = (ev - fix['eval_logits']).abs()
    assert (e2 <= 5e-2 + 5e-2 * fix['eval_logits'].abs()).all(), f'eval logits max err {e2.max().item():.4g}'


def test_two_sgd_steps_track_oracle():
    """Loss trajectory over optimizer steps (parameters updated by torch.optim.SGD as the
    reference does) stays within 2% of the oracle's."""
    from oracle import train_step
    from simpleaicv_pytorch_training_examples_b200.classification import backbones, losses
    arch, nc = 'resnet18cifar', 100
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(32, 3, 32, 32, generator=g), torch.randint(0, nc, (32,), generator=g)) for _ in range(3)]
    _, ref_losses = train_step.train_steps(arch, nc, 0, batches, lr=0.05)
    torch.manual_seed(0)
    model = backbones.__dict__[arch](num_classes=nc).cuda().train()
    decay = [p for p in model.parameters() if p.ndim > 1]
    no_decay = [p for p in model.parameters() if p.ndim == 1]
    opt = torch.optim.SGD([{'params': decay, 'weight_decay': 1e-4}, {'params': no_decay, 'weight_decay': 0.0}],
                          lr=0.05, momentum=0.9)
    crit = losses.CELoss()
    got = []
    for x, y in batches:
        loss = crit(model(x.cuda()), y.cuda())
        loss.backward()
        opt.step()
        opt.zero_grad()
        got.append(float(loss))
    for a, b in zip(got, ref_losses):
        assert abs(a - b) <= 2e-2 * abs(b), (got, ref_losses)


def test_cpu_input_raises():
    from simpleaicv_pytorch_training_examples_b200.classification import backbones
    m = backbones.resnet18cifar(num_classes=10)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 3, 32, 32))
