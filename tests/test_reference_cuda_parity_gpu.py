"""Parity at the BENCHMARKED shape against the reference's own CUDA path.

The unmodified reference modules (baseline/_ref) run on the same GPU under ``torch.autocast(bf16)`` — what
tools/scripts.py:153-170 does — on the same seeded weights and the same batch (256 x 3 x 224 x 224) as the
B200 runtime.  Compared: logits, loss, BN running statistics and every parameter gradient.

Tolerances: logits / loss 2e-2; gradients relative L2 <= 2e-2 and cosine >= 0.999 on
well-conditioned tensors.  Which tensors are well conditioned is MEASURED, not assumed: the reference is also
run in fp32, and ``noise = relL2(ref_bf16, ref_fp32)`` is the error the reference's own bf16 path makes on that
tensor.  A tensor with noise <= 1e-2 is well conditioned and must meet the tight bound against the reference's
bf16 result; the others (early layers of a randomly initialised network, listed in the output) must satisfy
relL2(ours, ref_fp32) <= 2 * noise + 2e-2, i.e. be no further from the exact gradient than the reference itself.
"""
import pytest
import torch

from baseline import ref_import

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason='reference not installed (baseline/install_ref.sh)')]


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _ref_run(model, crit, x, y, autocast):
    model.zero_grad(set_to_none=True)
    with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
        out = model(x)
        loss = crit(out, y)
    loss.backward()
    torch.cuda.synchronize()
    return out.detach().float(), loss.detach().float(), {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}


CASES = [('resnet50', {}, 'CELoss', 256), ('vit_base_patch16', {'image_size': 224, 'global_pool': True}, 'OneHotLabelCELoss', 256)]


@pytest.mark.parametrize('arch,kw,loss_name,batch', CASES, ids=[c[0] for c in CASES])
def test_bs256_matches_reference_cuda_autocast(arch, kw, loss_name, batch):
    from simpleaicv_pytorch_training_examples_b200.classification import backbones as mine_b, losses as mine_l
    ref_b = ref_import.backbones()
    ref_l = ref_import.module('SimpleAICV.classification.losses')
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(batch, 3, 224, 224, generator=g).cuda()
    lab = torch.randint(0, 1000, (batch,), generator=g)
    if loss_name == 'CELoss':
        y = lab.cuda()
    else:
        oh = torch.nn.functional.one_hot(lab, 1000).float() * 0.9 + 0.1 / 1000
        y = (0.5 * oh + 0.5 * oh.roll(1, 0)).cuda()
    torch.manual_seed(0)
    ref = ref_b.__dict__[arch](num_classes=1000, **kw).cuda().train()
    torch.manual_seed(0)
    mine = mine_b.__dict__[arch](num_classes=1000, **kw).cuda().train()
    assert all(torch.equal(a, b) for a, b in zip(ref.state_dict().values(), mine.state_dict().values()))
    sd0 = {k: v.clone() for k, v in ref.state_dict().items()}
    rcrit, mcrit = ref_l.__dict__[loss_name]().cuda(), mine_l.__dict__[loss_name]().cuda()

    lo32, ls32, g32 = _ref_run(ref, rcrit, x, y, autocast=False)
    ref.load_state_dict(sd0)                                   # undo the BN running-stat update
    lo16, ls16, g16 = _ref_run(ref, rcrit, x, y, autocast=True)
    ref_stats = {k: v.clone() for k, v in ref.state_dict().items() if 'running_' in k}

    out = mine(x)
    loss = mcrit(out, y)
    loss.backward()
    torch.cuda.synchronize()
    lom, lsm = out.detach().float(), loss.detach().float()
    gm = {n: p.grad.detach().float() for n, p in mine.named_parameters() if p.grad is not None}

    # ---- logits / loss: 2e-2, measured against the reference's bf16-autocast CUDA output
    scale = lo16.abs().max().item()
    err = (lom - lo16).abs().max().item()
    noise_logits = (lo16 - lo32).abs().max().item()
    print(f'{arch}: logits max|ours-ref_bf16| {err:.4g} (ref_bf16 vs ref_fp32 {noise_logits:.4g}, max|logit| {scale:.3g}); '
          f'loss ours {lsm.item():.6f} ref_bf16 {ls16.item():.6f} ref_fp32 {ls32.item():.6f}')
    assert err <= 2e-2 + 2e-2 * scale + 2 * noise_logits
    assert abs(lsm.item() - ls16.item()) <= 2e-2 * abs(ls16.item()) + 2e-2
    # ---- BN running statistics: rtol 1e-3 against the fp32 run's statistics is too tight for bf16 inputs of
    #      either implementation; both bf16 paths are held to 1e-2 of the tensor scale
    ms = mine.state_dict()
    for k, v in ref_stats.items():
        assert (ms[k] - v).abs().max().item() <= 1e-2 * v.abs().max().item() + 1e-3, k
    # ---- gradients
    assert set(g16) <= set(gm)
    assert all(gm[n].abs().max().item() == 0 for n in set(gm) - set(g16))   # params the reference gives no gradient
    well, ill, failures = 0, [], []
    for n in g16:
        noise = _rel(g16[n], g32[n])
        if noise <= 1e-2:
            well += 1
            r, c = _rel(gm[n], g16[n]), _cos(gm[n], g16[n])
            if not (r <= 2e-2 and c >= 0.999):
                failures.append(f'{n}: relL2 {r:.3g} cos {c:.5f} vs ref_bf16 (well conditioned, noise {noise:.2g})')
        else:
            r = _rel(gm[n], g32[n])
            ill.append((n, noise, r))
            if not r <= 2 * noise + 2e-2:
                failures.append(f'{n}: relL2 to fp32 {r:.3g} > 2 x reference noise {noise:.3g} + 2e-2')
    cat = lambda d: torch.cat([d[n].flatten() for n in g16])
    whole = _rel(cat(gm), cat(g16))
    print(f'{arch}: {well} well-conditioned gradient tensors within 2e-2 / cos 0.999 of the reference bf16 run; '
          f'{len(ill)} ill-conditioned (reference bf16-vs-fp32 noise > 1e-2), worst 5: '
          + ', '.join(f'{n} noise {a:.2g} ours {b:.2g}' for n, a, b in sorted(ill, key=lambda t: -t[1])[:5])
          + f'; whole-gradient relL2 vs ref_bf16 {whole:.4g} (ref noise {_rel(cat(g16), cat(g32)):.4g})')
    assert not failures, f'{len(failures)} gradient tensors out of tolerance: ' + '; '.join(failures[:10])
