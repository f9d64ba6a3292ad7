"""Host-side logic of the tools layer: optimizer parameter groups and LR schedule.

Runs on CPU.  Where /root/reference exists the results are compared with the reference's own
tools/utils.py (its `calflops` import is stubbed — the hot path never calls it)."""
import math
import os
import sys
import types

import pytest
import torch

from simpleaicv_pytorch_training_examples_b200.classification import backbones
from simpleaicv_pytorch_training_examples_b200.tools import utils as my_utils

REF = '/root/reference'


class _Cfg:
    optimizer = ('SGD', {'lr': 0.1, 'momentum': 0.9, 'global_weight_decay': False, 'weight_decay': 1e-4,
                         'no_weight_decay_layer_name_list': []})
    scheduler = ('MultiStepLR', {'warm_up_epochs': 1, 'gamma': 0.1, 'milestones': [30, 60, 90]})
    epochs = 100


def test_one_dim_params_are_not_decayed():
    m = backbones.resnet18cifar(num_classes=10)
    opt, info = my_utils.build_optimizer(_Cfg, m)
    for g in opt.param_groups:
        nd = {p.ndim for p in g['params']}
        assert (g['weight_decay'] == 0.) == (nd == {1}), (g['weight_decay'], nd)
    assert sum(len(g['params']) for g in opt.param_groups) == len(list(m.parameters()))


@pytest.mark.parametrize('name,params', [('MultiStepLR', {'warm_up_epochs': 2, 'gamma': 0.1, 'milestones': [3, 6]}),
                                         ('CosineLR', {'warm_up_epochs': 1, 'min_lr': 1e-6}),
                                         ('PolyLR', {'warm_up_epochs': 0, 'power': 0.9})])
def test_scheduler_closed_forms(name, params):
    class C(_Cfg):
        scheduler = (name, params)
        epochs = 10
    m = torch.nn.Linear(4, 4)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    s = my_utils.Scheduler(C, opt)
    for e in [0.0, 0.5, 1.0, 2.5, 3.0, 6.1, 9.99]:
        s.step(opt, e)
        w = params['warm_up_epochs']
        if e < w:
            exp = e / w * 0.1
        elif name == 'MultiStepLR':
            exp = 0.1 * 0.1 ** len([x for x in params['milestones'] if x <= e])
        elif name == 'CosineLR':
            exp = 0.5 * (math.cos((e - w) / (10 - w) * math.pi) + 1) * (0.1 - 1e-6) + 1e-6
        else:
            exp = (1 - (e - w) / (10 - w)) ** 0.9 * 0.1
        assert abs(opt.param_groups[0]['lr'] - exp) < 1e-12 and abs(s.current_lr - exp) < 1e-12


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'tools')), reason='reference checkout not present')
def test_groups_and_schedule_match_reference_utils():
    if 'calflops' not in sys.modules:
        stub = types.ModuleType('calflops')
        stub.calculate_flops = lambda *a, **k: None
        sys.modules['calflops'] = stub
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from tools import utils as ref_utils
    m = backbones.resnet50cifar(num_classes=10)
    ro, _ = ref_utils.build_optimizer(_Cfg, m)
    mo, _ = my_utils.build_optimizer(_Cfg, m)

    def canon(o):
        return sorted((g['weight_decay'], g['lr'], sorted(id(p) for p in g['params'])) for g in o.param_groups)

    assert canon(ro) == canon(mo)
    rs, ms = ref_utils.Scheduler(_Cfg, ro), my_utils.Scheduler(_Cfg, mo)
    for e in [0.0, 0.5, 0.99, 1.0, 29.9, 30.0, 61.2, 95.0]:
        rs.step(ro, e)
        ms.step(mo, e)
        assert abs(rs.current_lr - ms.current_lr) < 1e-12
        assert sorted(g['lr'] for g in ro.param_groups) == sorted(g['lr'] for g in mo.param_groups)
