"""Loss tails of knowledge distillation with the reference's names and arithmetic
(SimpleAICV/distillation/losses.py:14-113): fp32 torch on [B, classes] logits."""
import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ['CELoss', 'OneHotLabelCELoss', 'KDLoss', 'DMLLoss', 'L2Loss']


class CELoss(nn.Module):

    def __init__(self):
        super().__init__()
        self.loss = nn.CrossEntropyLoss(reduction='mean')

    def forward(self, pred, label):
        return self.loss(pred.float(), label)


class OneHotLabelCELoss(nn.Module):

    def forward(self, pred, label):
        return torch.sum(-label * F.log_softmax(pred.float(), dim=-1), dim=-1).mean()


def _clamped_softmax(x, t):
    return torch.clamp(F.softmax(x.float() / t, dim=1), min=1e-4, max=1. - 1e-4)


class KDLoss(nn.Module):

    def __init__(self, T):
        super().__init__()
        self.t = T

    def forward(self, stu_preds, tea_preds):
        s = torch.log(_clamped_softmax(stu_preds, self.t))
        t = _clamped_softmax(tea_preds, self.t)
        return F.kl_div(s, t, reduction='batchmean') * (self.t ** 2)


class DMLLoss(nn.Module):

    def __init__(self, T):
        super().__init__()
        self.t = T

    def forward(self, stu_preds, tea_preds):
        # the probabilities and the log-probabilities come from separate softmax evaluations, as in the reference (:75-98):
        # same autograd graph, same accumulation order of the student's gradient
        stu, tea = _clamped_softmax(stu_preds, self.t), _clamped_softmax(tea_preds, self.t)
        stu_log = torch.log(_clamped_softmax(stu_preds, self.t))
        tea_log = torch.log(_clamped_softmax(tea_preds, self.t))
        return (F.kl_div(stu_log, tea, reduction='batchmean') * (self.t ** 2) +
                F.kl_div(tea_log, stu, reduction='batchmean') * (self.t ** 2)) / 2.0


class L2Loss(nn.Module):

    def __init__(self):
        super().__init__()
        self.loss = nn.MSELoss(reduction='mean')

    def forward(self, stu_preds, tea_preds):
        return self.loss(stu_preds.float(), tea_preds.float())
