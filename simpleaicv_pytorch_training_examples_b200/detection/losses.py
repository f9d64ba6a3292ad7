"""DETR criterion with the reference's interface (SimpleAICV/detection/losses.py:843-1095 DETRLoss): Hungarian
matching of the last decoder layer's predictions to the ground-truth boxes (class probability, L1 and GIoU costs;
SciPy's linear_sum_assignment on the host, integer indices), then per decoder layer a weighted cross-entropy over all
queries (unmatched queries target the no-object class), an L1 and a GIoU loss over the matched pairs.  fp32 torch on the
device: the criterion consumes [6, B, 100, 81] / [6, B, 100, 4] tensors, negligible next to the model.

annotations: fp32 [B, max_boxes, 5] rows (cx, cy, w, h, class) normalised by the image size, class -1 = padding row.
Returns the reference's dict: layer_{i}_cls_loss / layer_{i}_box_l1_loss / layer_{i}_box_iou_loss, i = 0..5.
"""
import numpy as np
import scipy.optimize
import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ['DETRLoss']


def cxcywh_to_xyxy(b):
    half = 0.5 * b[:, 2:4]
    return torch.cat((b[:, 0:2] - half, b[:, 0:2] + half), dim=1)


def pairwise_giou(a, b):
    """GIoU of every box of a [N, 4] with every box of b [M, 4] (xyxy) -> [N, M]; the reference's clamps are kept
    (areas and intersections >= 0, union and enclosing area >= 1e-4)."""
    area_a = ((a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])).clamp(min=0)
    area_b = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).clamp(min=0)
    wh = (torch.min(a[:, None, 2:], b[:, 2:]) - torch.max(a[:, None, :2], b[:, :2])).clamp(min=0)
    inter = (wh[..., 0] * wh[..., 1]).clamp(min=0)
    union = (area_a[:, None] + area_b - inter).clamp(min=1e-4)
    ewh = (torch.max(a[:, None, 2:], b[:, 2:]) - torch.min(a[:, None, :2], b[:, :2])).clamp(min=0)
    enclose = (ewh[..., 0] * ewh[..., 1]).clamp(min=1e-4)
    return inter / union - (enclose - union) / enclose


def assignment_with_inf(cost):
    """scipy.optimize.linear_sum_assignment on a cost matrix that may hold NaN (-> 1e5) or infinities of one sign
    (-> a finite value beyond every other entry, scaled so that it still dominates any assignment)."""
    cost = np.asarray(cost)
    if np.isnan(cost).any():
        cost[np.isnan(cost)] = 1e5
    neg, pos = np.isneginf(cost).any(), np.isposinf(cost).any()
    if neg and pos:
        raise ValueError('matrix contains both inf and -inf')
    if neg or pos:
        finite = cost[~np.isinf(cost)]
        lo, hi = finite.min(), finite.max()
        m = min(cost.shape)
        margin = m * (hi - lo + np.abs(hi) + np.abs(lo) + 1)
        cost[np.isinf(cost)] = (hi + (m - 1) * (hi - lo)) + margin if pos else (lo + (m - 1) * (lo - hi)) - margin
    return scipy.optimize.linear_sum_assignment(cost)


class DETRLoss(nn.Module):

    def __init__(self, cls_match_cost=1.0, box_match_cost=5.0, giou_match_cost=2.0, cls_loss_weight=1.0, box_l1_loss_weight=5.0,
                 iou_loss_weight=2.0, no_object_cls_weight=0.1, num_classes=80):
        super().__init__()
        assert cls_match_cost != 0 or box_match_cost != 0 or giou_match_cost != 0, 'all costs cant be 0'
        self.cls_match_cost, self.box_match_cost, self.giou_match_cost = cls_match_cost, box_match_cost, giou_match_cost
        self.cls_loss_weight, self.box_l1_loss_weight, self.iou_loss_weight = cls_loss_weight, box_l1_loss_weight, iou_loss_weight
        self.no_object_cls_weight = no_object_cls_weight
        self.num_classes = num_classes

    @torch.no_grad()
    def get_matched_pred_target_idxs(self, cls_preds, reg_preds, annotations):
        """Per image (query indices, ground-truth indices), int64 CPU tensors, minimising
        cls_match_cost * (-p[class]) + box_match_cost * L1 + giou_match_cost * (-GIoU)."""
        B, Q = cls_preds.shape[:2]
        prob = F.softmax(cls_preds.flatten(0, 1), dim=-1).clamp(min=1e-4, max=1. - 1e-4)
        boxes = reg_preds.flatten(0, 1)
        gts = [a[a[:, 4] >= 0] for a in annotations]
        counts = [g.shape[0] for g in gts]
        gt = torch.cat(gts, dim=0)
        cost = self.cls_match_cost * (-prob[:, gt[:, 4].long()]) + self.box_match_cost * torch.cdist(boxes, gt[:, 0:4], p=1) \
            + self.giou_match_cost * (-pairwise_giou(cxcywh_to_xyxy(boxes), cxcywh_to_xyxy(gt[:, 0:4])))
        cost = cost.view(B, Q, -1)
        out = []
        for i, block in enumerate(cost.split(counts, -1)):
            rows, cols = assignment_with_inf(block[i].cpu().numpy())
            out.append((torch.as_tensor(rows, dtype=torch.int64), torch.as_tensor(cols, dtype=torch.int64)))
        return out

    @staticmethod
    def _matched(indices):
        batch = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(indices)])
        return batch, torch.cat([src for src, _ in indices])

    def compute_batch_cls_loss(self, cls_preds, annotations, indices):
        B, Q = cls_preds.shape[:2]
        dev = cls_preds.device
        labels = torch.cat([a[a[:, 4] >= 0][:, 4][j] for a, (_, j) in zip(annotations, indices)])
        target = torch.full((B, Q), self.num_classes, dtype=torch.long, device=dev)
        target[self._matched(indices)] = labels.long()
        weight = torch.ones(self.num_classes + 1, device=dev)
        weight[-1] = self.no_object_cls_weight
        return F.cross_entropy(cls_preds.transpose(1, 2), target, weight)

    def compute_batch_l1_iou_loss(self, reg_preds, annotations, indices):
        pred = reg_preds[self._matched(indices)]
        gts = [a[a[:, 4] >= 0][:, 0:4] for a in annotations]
        n = sum(g.shape[0] for g in gts)
        tgt = torch.cat([g[j] for g, (_, j) in zip(gts, indices)], dim=0)
        l1 = F.l1_loss(pred, tgt, reduction='none').sum() / n
        giou = torch.diag(pairwise_giou(cxcywh_to_xyxy(pred), cxcywh_to_xyxy(tgt)))
        return l1, (1 - giou).sum() / n

    def forward(self, preds, annotations):
        cls_preds, reg_preds = preds
        reg_preds = torch.clamp(reg_preds, min=1e-4, max=1. - 1e-4).float()
        cls_preds = cls_preds.float()
        annotations = annotations.float()
        indices = self.get_matched_pred_target_idxs(cls_preds[-1], reg_preds[-1], annotations)
        losses = {}
        for i, (c, r) in enumerate(zip(cls_preds, reg_preds)):
            l1, iou = self.compute_batch_l1_iou_loss(r, annotations, indices)
            losses[f'layer_{i}_cls_loss'] = self.cls_loss_weight * self.compute_batch_cls_loss(c, annotations, indices)
            losses[f'layer_{i}_box_l1_loss'] = self.box_l1_loss_weight * l1
            losses[f'layer_{i}_box_iou_loss'] = self.iou_loss_weight * iou
        return losses
