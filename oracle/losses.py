"""ORACLE (test infrastructure, never shipped or measured as the product).

torch-CPU restatement of the loss value of the reference's
``TranscriptionModel.compute_loss`` for the CNN-RNN models
(models/transcription_model.py:110-163 single head, :165-190 three heads,
:192-217 per-head helper): BCE-with-logits against the piano roll, linear time
interpolation when the logits' frame count differs (:140-142), masking to
``lengths`` with the ``max(valid * 88, 1)`` denominator (:148-163), and the
0.5 / 0.25 / 0.25 head weighting with onset / offset targets taken from the
roll's positive / negative time differences (:176-189).

Pinned by tests/golden/loss_reference.npz, produced by running the reference's
own method (oracle/make_golden.py gen_losses).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def single_head(logits: torch.Tensor, targets: torch.Tensor, lengths=None) -> torch.Tensor:
    logits, targets = logits.float(), targets.float()
    if logits.shape[-1] != targets.shape[-1]:
        logits = F.interpolate(logits, size=targets.shape[-1], mode="linear", align_corners=False)
    per = F.binary_cross_entropy_with_logits(logits, targets, reduction="none")
    if lengths is None:
        return per.mean()
    B, P, T = logits.shape
    mask = (torch.arange(T).unsqueeze(0) < torch.as_tensor(lengths).unsqueeze(1)).unsqueeze(1)
    return (per * mask).sum() / (mask.sum() * P).clamp_min(1)


def onset_offset_targets(targets: torch.Tensor):
    on, off = torch.zeros_like(targets), torch.zeros_like(targets)
    if targets.shape[-1] > 1:
        on[:, :, 1:] = torch.clamp(targets[:, :, 1:] - targets[:, :, :-1], min=0)
        off[:, :, :-1] = torch.clamp(targets[:, :, :-1] - targets[:, :, 1:], min=0)
    return on, off


def compute_loss(logits, targets, lengths=None) -> torch.Tensor:
    if not isinstance(logits, dict):
        return single_head(logits, targets, lengths)
    targets = targets.float()
    on, off = onset_offset_targets(targets)
    return (0.5 * single_head(logits["frame"], targets, lengths) + 0.25 * single_head(logits["onset"], on, lengths)
            + 0.25 * single_head(logits["offset"], off, lengths))
