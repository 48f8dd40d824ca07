"""Pooling classifiers with the reference's signatures (utils/patch_selection_classifier.py).

``topj_pooling`` is the one MOC's train/eval loops use (main_moc.py:405,:493); the delta_softmax / delta_diff /
bottomk_irrel variants are the ones ``zs_evaluation(pooling_func=...)`` accepts (main_moc.py:12-13,:429-432).
Each returns ``(preds {j: argmax}, pooled {j: [1,C]}[, indices [maxj,C]])`` for every j in ``topj``.
"""
from __future__ import annotations

import torch

from . import _lib, ops
from ._lib import MocError
from .selectors import _check, _maxj


def _finish(values: torch.Tensor, indices: torch.Tensor, topj, maxj: int, return_indices: bool):
    pooled = {j: ops.col_prefix_mean(values, min(j, maxj, values.size(0))) for j in topj}
    preds = {j: v.argmax(dim=1) for j, v in pooled.items()}
    return (preds, pooled, indices) if return_indices else (preds, pooled)


def topj_pooling(logits, topj, return_indices=False, **kwargs):
    """Mean of the min(j, N) largest logits of every class column (:18-32)."""
    logits = _check(logits)
    maxj = _maxj(topj, logits.size(0))
    idx, vals = ops.topj_sorted(logits, maxj, largest=True, want_values=True)
    return _finish(vals, idx, topj, maxj, return_indices)


def delta_softmax_classifier_pooling(logits, topj, return_indices=False, **kwargs):
    """Rows chosen by softmax probability, their raw logits pooled (:35-53)."""
    logits = _check(logits)
    c = logits.size(1)
    maxj = _maxj(topj, logits.size(0))
    keys = ops.row_keys(logits, c)
    idx = ops.topj_sorted(keys[c:2 * c].t(), maxj, largest=True)
    vals = torch.stack([ops.take_rows(logits[:, i:i + 1], idx[:, i], 1)[:, 0] for i in range(c)], dim=1)
    return _finish(vals, idx, topj, maxj, return_indices)


def delta_diff_classifier_pooling(logits, topj, return_indices=False, **kwargs):
    """Rows chosen by |top1 - top2|, their whole logit rows pooled (:56-78)."""
    logits = _check(logits)
    c = logits.size(1)
    if c < 2:
        raise MocError(_lib.E_SHAPE, "delta_diff needs at least two classes")
    maxj = _maxj(topj, logits.size(0))
    keys = ops.row_keys(logits, c)
    idx1 = ops.topj_sorted(keys[2 * c], maxj, largest=True)
    vals = ops.take_rows(logits, idx1, c)
    return _finish(vals, idx1.unsqueeze(1).expand(-1, c).contiguous(), topj, maxj, return_indices)


def bottomk_irrel_classifier_pooling(logits, topj, return_indices=False, coords_list=None, bottomk=None,
                                     detection=False, **kwargs):
    """Foreground logits pooled over the rows least like background (:127-171)."""
    assert coords_list is not None, "coords_list should be provided"
    logits = _check(logits)
    if type(coords_list) == int:
        assert logits.size(1) > coords_list, "logits should have more bg classes"
        n_fg = coords_list
    elif type(coords_list) == list:
        assert logits.size(1) > len(coords_list), "logits should have more bg classes"
        n_fg = len(coords_list)
    else:
        raise ValueError("coords_list should be int or list")
    if detection:
        raise MocError(_lib.E_SHAPE, "detection=True is unused by MOC and not built")
    maxj = _maxj(topj, logits.size(0))
    if bottomk is None:
        bottomk = maxj
    keys = ops.row_keys(logits, n_fg)
    bg_idx = ops.topj_sorted(keys[2 * n_fg + 1], bottomk, largest=False)
    fg = ops.take_rows(logits, bg_idx, n_fg)
    fg_idx, fg_vals = ops.topj_sorted(fg, min(maxj, bg_idx.numel()), largest=True, want_values=True)
    return _finish(fg_vals, bg_idx[fg_idx], topj, maxj, return_indices)
