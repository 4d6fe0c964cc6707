"""Training losses with the reference's names and call conventions (reference model/criterion.py:8-259).  They are
image-space elementwise reductions outside the four hot-path modules (SURVEY.md 8f "next" #1) and stay plain PyTorch."""
import math

import torch
import torch.nn.functional as F
from torch import nn


def temporal_weight_func(T):
    """w_t = exp(t * ln(T)/(T-1)), t = 0..T-1 (reference :8-13)."""
    t = torch.linspace(0, T - 1, T)
    return torch.exp(math.log(T) / (T - 1) * t)


def _apply_temporal_weight(e, w):
    if w is None:
        return e
    w = w.to(e.device)
    return e * w.view(1, -1, *([1] * (e.dim() - 2)))


class _PixelLoss(nn.Module):
    def __init__(self, temporal_weight=None, norm_dim=None):
        super().__init__()
        self.temporal_weight, self.norm_dim = temporal_weight, norm_dim

    def _err(self, d):
        raise NotImplementedError

    def __call__(self, gt, pred):
        if self.norm_dim is not None:
            gt = F.normalize(gt, p=2, dim=self.norm_dim)
            pred = F.normalize(pred, p=2, dim=self.norm_dim)
        return _apply_temporal_weight(self._err(pred - gt), self.temporal_weight).mean()


class L1Loss(_PixelLoss):
    def _err(self, d):
        return d.abs()


class MSELoss(_PixelLoss):
    def _err(self, d):
        return d.square()


class GDL(nn.Module):
    """Gradient-difference loss (reference :134-204): mean | |dy gt| - |dy pred| |^a + mean | |dx gt| - |dx pred| |^a."""

    def __init__(self, alpha=1, temporal_weight=None):
        super().__init__()
        self.alpha, self.temporal_weight = alpha, temporal_weight

    def __call__(self, gt, pred):
        lead = gt.shape[:-3]
        g, p = gt.flatten(0, -4), pred.flatten(0, -4)
        terms = []
        for dim in (2, 3):
            n = g.shape[dim]
            dg = (g.narrow(dim, 1, n - 1) - g.narrow(dim, 0, n - 1)).abs()
            dp = (p.narrow(dim, 1, n - 1) - p.narrow(dim, 0, n - 1)).abs()
            e = (dg - dp).abs()
            if self.alpha != 1:
                e = e.pow(self.alpha)
            if self.temporal_weight is not None:
                assert self.temporal_weight.shape[0] == lead[1], "Mismatch between temporal_weight and predicted sequence length"
                e = _apply_temporal_weight(e.reshape(*lead, *e.shape[1:]), self.temporal_weight)
            terms.append(e.mean())
        return terms[0] + terms[1]


class BiPatchNCE(nn.Module):
    """Bidirectional patch-wise contrastive loss (reference :206-259): positives = same (frame, position); gradients
    flow to the prediction only through the positive pairs."""

    def __init__(self, N, T, h, w, temperature=0.07):
        super().__init__()
        self.register_buffer('mask', torch.eye(h * w).long().unsqueeze(0).repeat(N * T, 1, 1))
        self.temperature = temperature

    def forward(self, gt_f, pred_f):
        mask = self.mask
        gt = gt_f.permute(0, 1, 3, 4, 2).flatten(0, 1).flatten(1, 2)       # (N*T, h*w, C)
        pr = pred_f.permute(0, 1, 3, 4, 2).flatten(0, 1).flatten(1, 2)
        neg = 1.0 - mask

        def scores(a, b):
            return (torch.matmul(a, b.transpose(1, 2)) * mask + torch.matmul(a, b.detach().transpose(1, 2)) * neg) / self.temperature

        target = torch.arange(mask.shape[1], device=gt.device).repeat(mask.shape[0])
        l1 = F.cross_entropy(scores(gt, pr).flatten(0, 1), target)
        l2 = F.cross_entropy(scores(pr, gt).flatten(0, 1), target)
        return (l1 + l2) * 0.5


class GANLoss(nn.Module):
    """GAN objectives 'lsgan' | 'vanilla' | 'wgangp' (reference :15-74)."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer('real_label', torch.tensor(target_real_label))
        self.register_buffer('fake_label', torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        if gan_mode == 'lsgan':
            self.loss = nn.MSELoss()
        elif gan_mode == 'vanilla':
            self.loss = nn.BCEWithLogitsLoss()
        elif gan_mode == 'wgangp':
            self.loss = None
        else:
            raise NotImplementedError('gan mode %s not implemented' % gan_mode)

    def get_target_tensor(self, prediction, target_is_real):
        return (self.real_label if target_is_real else self.fake_label).expand_as(prediction)

    def __call__(self, prediction, target_is_real):
        if self.gan_mode == 'wgangp':
            return -prediction.mean() if target_is_real else prediction.mean()
        return self.loss(prediction, self.get_target_tensor(prediction, target_is_real))
