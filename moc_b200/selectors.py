"""The four index selectors with the reference's signatures (utils/patch_selection_classifier_index.py).

Each returns a LongTensor ``[maxj, C]`` of row indices sorted by decreasing key, like ``Tensor.topk``;
ties are broken towards the lower row index (torch leaves tie order unspecified).  Inputs are the per-patch
logits ``[N, C]`` exactly as the reference passes them; the per-row transforms (softmax, |top1-top2|,
background sum) run in a small CUDA kernel of ours, the selection in the radix-select kernel.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import MocError


def _maxj(topj, n_rows: int) -> int:
    return min(max(topj), n_rows)  # _index.py:24


def _check(logits: torch.Tensor) -> torch.Tensor:
    if not isinstance(logits, torch.Tensor) or logits.dim() != 2:
        raise MocError(_lib.E_SHAPE, "logits must be an [N,C] tensor")
    if not logits.is_cuda:
        raise MocError(_lib.E_ARG, "logits must live on a CUDA device: moc_b200 has no CPU path")
    return logits.float().contiguous()


def index_topj_classifier(logits, topj, **kwargs):
    """Per class, the rows with the maxj largest logits (_index.py:17-26)."""
    logits = _check(logits)
    return ops.topj_sorted(logits, _maxj(topj, logits.size(0)), largest=True)


def index_delta_softmax_classifier(logits, topj, **kwargs):
    """Per class, the rows with the maxj largest row-softmax probabilities (_index.py:28-36)."""
    logits = _check(logits)
    keys = ops.row_keys(logits, logits.size(1))
    c = logits.size(1)
    return ops.topj_sorted(keys[c:2 * c].t(), _maxj(topj, logits.size(0)), largest=True)


def index_delta_diff_classifier(logits, topj, **kwargs):
    """Rows with the largest |top1 - top2| margin, replicated over the C columns (_index.py:38-51)."""
    logits = _check(logits)
    c = logits.size(1)
    if c < 2:
        raise MocError(_lib.E_SHAPE, "delta_diff needs at least two classes (the reference's topk(2, dim=1) raises)")
    keys = ops.row_keys(logits, c)
    idx = ops.topj_sorted(keys[2 * c], _maxj(topj, logits.size(0)), largest=True)
    return idx.unsqueeze(1).expand(-1, c).contiguous()


def index_bottomk_irrel_classifier(logits, topj, n_classes, bottomk=None, detection=False, **kwargs):
    """Rows least like the background prompts, re-ranked per foreground class (_index.py:53-87)."""
    assert n_classes is not None, "coords_list should be provided"
    logits = _check(logits)
    assert logits.size(1) > n_classes, "logits should have more bg classes"
    if detection:
        raise MocError(_lib.E_SHAPE, "detection=True is unused by MOC and not built (SURVEY.md section 8a, row a6)")
    maxj = _maxj(topj, logits.size(0))
    if bottomk is None:
        bottomk = maxj
    bottomk = min(bottomk, logits.size(0))
    keys = ops.row_keys(logits, n_classes)
    bg_idx = ops.topj_sorted(keys[2 * n_classes + 1], bottomk, largest=False)       # smallest background sum
    fg = ops.take_rows(logits, bg_idx, n_classes)                                     # [bottomk, C]
    fg_idx = ops.topj_sorted(fg, min(maxj, bottomk), largest=True)                    # [maxj, C]
    return bg_idx[fg_idx]
