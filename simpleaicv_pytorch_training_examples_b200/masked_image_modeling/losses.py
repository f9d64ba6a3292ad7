"""Loss tails of MAE pre-training with the reference's names and arithmetic
(SimpleAICV/masked_image_modeling/losses.py:11-48): fp32 torch on the [B, L, p*p*3] predictions."""
import torch
import torch.nn as nn

__all__ = ['MSELoss', 'L1Loss']


class MSELoss(nn.Module):
    """Mean squared error per patch, averaged over the REMOVED patches (mask == 1)."""

    def forward(self, pred, label, mask):
        pred, label, mask = pred.float(), label.float(), mask.float()
        loss = ((pred - label) ** 2).mean(dim=-1)
        return (loss * mask).sum() / (mask.sum() + 1e-4)


class L1Loss(nn.Module):

    def forward(self, pred, label, mask):
        pred, label, mask = pred.float(), label.float(), mask.float()
        loss = torch.abs(pred - label)
        return (loss * mask).sum() / (mask.sum() + 1e-4)
