"""``slide_process`` with the reference's signature and return value (main_moc.py:322-375).

One call = score every patch of the bag against the two prompt matrices, make the four top-J selections,
take their union in ascending order and return the selected features with their four score planes.  All of
it runs in the CUDA kernels of libmoc_b200; the only host round trip is reading the selected count (the
reference syncs four times, once per ``.tolist()``, main_moc.py:343-352).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib, ops
from ._lib import MocError

_PROMPT_CACHE: Dict[tuple, ops.Prompts] = {}


def prompts_for(w: torch.Tensor, w_ext: torch.Tensor) -> ops.Prompts:
    """Packed prompt matrices, cached on the identity + version of the two tensors (they are run constants)."""
    key = (w.data_ptr(), w_ext.data_ptr(), w._version, w_ext._version, tuple(w.shape), tuple(w_ext.shape))
    p = _PROMPT_CACHE.get(key)
    if p is None:
        if len(_PROMPT_CACHE) > 16:
            _PROMPT_CACHE.clear()
        p = ops.Prompts.pack(w, w_ext)
        _PROMPT_CACHE[key] = p
    return p


def slide_process(feat, zeroshot_weights, zeroshot_weights_ext, n_classes, topj=10, random_mask=False,
                  discard_classifiers=[], mask: Optional[torch.Tensor] = None) -> dict:
    """Drop-in for main_moc.py:322-375.  ``mask`` (bool [N]) is an addition that replaces the random draw.

    With ``random_mask=True`` the half mask is drawn exactly like the reference - ``torch.rand(N) > 0.5`` on
    the CPU default generator (main_moc.py:330) - but the bag is not copied: the mask goes to the selection
    kernel and ``selected_index`` still indexes the *masked* bag, as in the reference.
    """
    device = zeroshot_weights.device
    if device.type != "cuda":
        raise MocError(_lib.E_ARG, "zeroshot_weights must live on a CUDA device: moc_b200 has no CPU path")
    if zeroshot_weights.size(1) != n_classes:
        raise MocError(_lib.E_SHAPE, "n_classes=%d but zeroshot_weights has %d columns" % (n_classes, zeroshot_weights.size(1)))
    feat = feat.to(device=device, dtype=torch.float32).contiguous()
    zeroshot_weights_ext = zeroshot_weights_ext.to(device)
    n = feat.size(0)
    if mask is None and random_mask:
        mask = torch.rand(n) > 0.5
    mask_d = mask.to(device) if mask is not None else None

    prompts = prompts_for(zeroshot_weights, zeroshot_weights_ext)
    keys = ops.score_keys(feat, prompts, check_domain=True)   # |x| >= 65504 with a wide prompt set: fp32 kernel instead
    offs_h = [0, n]
    offs = torch.tensor(offs_h, dtype=torch.int64, device=device)
    sel = ops.select_union(keys, offs, offs_h, n_classes, int(topj), _lib.discard_bits(discard_classifiers), mask_d)
    count = int(sel.sel_count[0])
    sel_feat, planes = ops.gather_selected(feat, keys, n_classes, sel.sel_rows, count)
    return {
        "selected_index": sel.sel_local[:count].tolist(),
        "selected_feat": sel_feat,
        "logits_top_classifier": planes[0],
        "logits_delta_softmax_classifier": planes[1],
        "logits_delta_diff_classifier": planes[2],
        "logits_bottomk_irrel_classifier": planes[3],
    }
