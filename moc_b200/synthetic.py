"""Seeded synthetic cohorts: CONCH-shaped 512-d bags with a planted class signal.

No CONCH checkpoint or slide data is available offline, so every config in
BASELINE.json runs on these.  The recipe gives bags whose
zero-shot AUC is well away from 0.5, so accuracy / AUC parity is meaningful:

* prompt matrix ``W_all = normalise(randn(C+4, 512))``; ``W = W_all[:C].T`` and
  ``W_ext = W_all.T`` - unit-norm columns as ``zero_shot_classifier`` produces
  (reference ``utils/zeroshot_utils.py:38-50``), with ``W_ext[:, :C] == W`` as in
  every shipped prompt file;
* a bag is ``randn(N,512)/sqrt(512)`` plus, for the first 4 % of rows, a pull
  towards a class prompt (the bag label with probability ``purity``), for the
  rest a weaker pull towards one of the four normal-tissue prompts, plus a common
  offset; rows are L2-normalised and shuffled.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

D = 512
N_NORMAL = 4  # Stroma / Inflammation / Vascular / Necrosis columns of the *_w4normal prompt files


def prompt_matrices(n_classes: int, seed: int = 1234, device="cpu", n_normal: int = N_NORMAL
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (W [512,C], W_ext [512,C+n_normal]), fp32, contiguous, unit-norm columns."""
    g = torch.Generator().manual_seed(seed)
    w_all = torch.randn(n_classes + n_normal, D, generator=g)
    w_all = w_all / w_all.norm(dim=1, keepdim=True)
    w = w_all[:n_classes].t().contiguous()
    w_ext = w_all.t().contiguous()
    return w.to(device), w_ext.to(device)


def prompt_bank(n_classes: int, prompts_per_class: int, seed: int = 4321, device="cpu"
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """An un-collapsed bank [512, C*P] of unit vectors and the collapsed [512, C]
    matrix the reference would score against (mean over prompts, renormalised:
    ``utils/zeroshot_utils.py:41-44``)."""
    g = torch.Generator().manual_seed(seed)
    centre = torch.randn(n_classes, 1, D, generator=g)
    bank = centre + 0.5 * torch.randn(n_classes, prompts_per_class, D, generator=g)
    bank = bank / bank.norm(dim=2, keepdim=True)
    collapsed = bank.mean(dim=1)
    collapsed = collapsed / collapsed.norm(dim=1, keepdim=True)
    return (bank.reshape(n_classes * prompts_per_class, D).t().contiguous().to(device),
            collapsed.t().contiguous().to(device))


def make_bag(n_patches: int, label: int, w_ext: torch.Tensor, n_classes: int, seed: int,
             device="cpu", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One bag [N,512] fp32.  ``w_ext`` must live on ``device``.  With ``out`` the
    bag is written into that (pre-allocated, contiguous) slice of a ragged store."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    n = int(n_patches)
    w_all = w_ext.t()  # [C+4, 512]
    x = torch.randn(n, D, generator=g, device=dev) * (1.0 / math.sqrt(D))
    n_tum = int(0.04 * n)
    u = torch.rand(n, 4, generator=g, device=dev)
    purity = 0.55 + 0.40 * float(torch.rand(1, generator=g, device=dev))
    if n_tum > 0:
        amp = 0.12 * (0.3 + 0.7 * u[:n_tum, 0])
        other = (label + 1 + (u[:n_tum, 1] * max(n_classes - 1, 1)).long().clamp_(max=max(n_classes - 2, 0))) % n_classes
        cls = torch.where(u[:n_tum, 2] < purity, torch.full_like(other, label), other)
        x[:n_tum] += amp.unsqueeze(1) * w_all[cls]
    if n > n_tum:
        n_bg = w_all.size(0) - n_classes
        bg = n_classes + (u[n_tum:, 1] * n_bg).long().clamp_(max=n_bg - 1)
        x[n_tum:] += (0.06 * u[n_tum:, 0]).unsqueeze(1) * w_all[bg]
    x += 0.15 * w_all.mean(dim=0, keepdim=True)
    x /= x.norm(dim=1, keepdim=True)
    perm = torch.randperm(n, generator=g, device=dev)
    if out is not None:
        torch.index_select(x, 0, perm, out=out)
        return out
    return x[perm].contiguous()


def slide_seed(cohort_seed: int, i: int) -> int:
    return cohort_seed * 100003 + i


def make_cohort(n_slides: int, n_patches, n_classes: int, cohort_seed: int = 0, device="cpu",
                w_ext: Optional[torch.Tensor] = None) -> Tuple[List[torch.Tensor], List[int]]:
    """List of bags and labels (label = i mod C).  ``n_patches`` is an int or a per-slide sequence."""
    if w_ext is None:
        _, w_ext = prompt_matrices(n_classes, device=device)
    sizes = [int(n_patches)] * n_slides if isinstance(n_patches, int) else [int(v) for v in n_patches]
    bags, labels = [], []
    for i in range(n_slides):
        y = i % n_classes
        bags.append(make_bag(sizes[i], y, w_ext, n_classes, slide_seed(cohort_seed, i), device))
        labels.append(y)
    return bags, labels


def log_uniform_sizes(n_slides: int, lo: int = 1000, hi: int = 100000, seed: int = 7) -> List[int]:
    """Bag sizes for the throughput sweep (BASELINE.json configs[4])."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n_slides, generator=g)
    return [int(round(math.exp(math.log(lo) + float(v) * (math.log(hi) - math.log(lo))))) for v in u]
