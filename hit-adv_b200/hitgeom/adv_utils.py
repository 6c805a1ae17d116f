"""Adversarial objectives on the victim's logits (util/adv_utils.py), device-agnostic: the one-hot mask is built
on the logits' device instead of `torch.zeros(B, K).cuda()` (adv_utils.py:29,59)."""
import torch
import torch.nn as nn
import torch.nn.functional as F_t


def _split(logits, targets):
    one_hot = torch.zeros_like(logits).scatter_(1, targets.view(-1, 1).long(), 1.0)
    real = torch.sum(one_hot * logits, dim=1)
    other = torch.max((1. - one_hot) * logits - one_hot * 10000., dim=1)[0]
    return real, other


class LogitsAdvLoss(nn.Module):
    """util/adv_utils.py:6-35: targeted margin loss, mean over the batch."""

    def __init__(self, kappa=0.):
        super().__init__()
        self.kappa = kappa

    def forward(self, logits, targets):
        real, other = _split(logits, targets)
        return torch.clamp(other - real + self.kappa, min=0.).mean()


class UntargetedLogitsAdvLoss(nn.Module):
    """util/adv_utils.py:38-67."""

    def __init__(self, kappa=0.):
        super().__init__()
        self.kappa = kappa

    def forward(self, logits, targets):
        real, other = _split(logits, targets)
        return torch.clamp(real - other + self.kappa, min=0.).mean()


class CrossEntropyAdvLoss(nn.Module):
    """util/adv_utils.py:70-85."""

    def forward(self, logits, targets):
        return F_t.cross_entropy(logits, targets)
