"""Synthetic detection task used where the reference would use a dataset (there is no network for datasets or checkpoints):
colour-coded rectangles and ellipses on a smooth noisy background.  Class = shape (rectangle / ellipse) x colour (4 hues) = 8
classes.  Vectorised torch, device-agnostic, fully determined by the generator's seed.

    images, targets = synth_batch(B, S, gen, device)      images (B,3,S,S) float in [0,1] quantised to k/255,
                                                           targets (n,6) rows [image, cls, cx, cy, w, h] normalised (ultralytics batch layout)
"""
from __future__ import annotations

import torch

NC = 8
_HUES = torch.tensor([[0.9, 0.15, 0.15], [0.15, 0.8, 0.2], [0.2, 0.3, 0.95], [0.95, 0.85, 0.1]])
KMAX = 4


def synth_batch(B: int, S: int, gen: torch.Generator, device="cpu"):
    r = lambda *s: torch.rand(*s, generator=gen, device=device)
    low = r(B, 3, S // 32, S // 32) * 0.5 + 0.2
    img = torch.nn.functional.interpolate(low, size=(S, S), mode="bilinear", align_corners=False)
    img = img + 0.08 * (r(B, 3, S, S) - 0.5)
    n_obj = 1 + (r(B) * KMAX).long().clamp_(max=KMAX - 1)                      # 1..KMAX objects
    cls = (r(B, KMAX) * NC).long().clamp_(max=NC - 1)
    w = (0.12 + 0.33 * r(B, KMAX)) * S
    h = (0.12 + 0.33 * r(B, KMAX)) * S
    cx = w / 2 + r(B, KMAX) * (S - w)
    cy = h / 2 + r(B, KMAX) * (S - h)
    bright = 0.75 + 0.25 * r(B, KMAX)
    ys = torch.arange(S, device=device, dtype=torch.float32).view(1, S, 1) + 0.5
    xs = torch.arange(S, device=device, dtype=torch.float32).view(1, 1, S) + 0.5
    hues = _HUES.to(device)
    for k in range(KMAX):
        dx = (xs - cx[:, k].view(B, 1, 1)) / (w[:, k].view(B, 1, 1) / 2)
        dy = (ys - cy[:, k].view(B, 1, 1)) / (h[:, k].view(B, 1, 1) / 2)
        rect = (dx.abs() <= 1) & (dy.abs() <= 1)
        ell = dx * dx + dy * dy <= 1
        is_ell = (cls[:, k] >= 4).view(B, 1, 1)
        mask = torch.where(is_ell, ell, rect) & (k < n_obj).view(B, 1, 1)
        colour = hues[cls[:, k] % 4] * bright[:, k].view(B, 1)                   # (B,3)
        img = torch.where(mask.unsqueeze(1), colour.view(B, 3, 1, 1).expand(B, 3, S, S), img)
    img = (img.clamp_(0, 1) * 255).round() / 255                                # exactly representable as uint8
    valid = torch.arange(KMAX, device=device).view(1, KMAX) < n_obj.view(B, 1)
    bi = torch.arange(B, device=device).view(B, 1).expand(B, KMAX)
    t = torch.stack([bi.float(), cls.float(), cx / S, cy / S, w / S, h / S], -1)[valid]
    return img, t


def labels_xyxy(targets: torch.Tensor, B: int, S: int):
    """targets (n,6) -> per-image numpy arrays (m,5) [cls, x1, y1, x2, y2] in pixels (what oracle/metrics_ref.evaluate takes)."""
    out = []
    t = targets.cpu()
    for b in range(B):
        r = t[t[:, 0] == b]
        xy, wh = r[:, 2:4] * S, r[:, 4:6] * S
        out.append(torch.cat([r[:, 1:2], xy - wh / 2, xy + wh / 2], 1).numpy())
    return out
