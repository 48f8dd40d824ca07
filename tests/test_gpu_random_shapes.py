"""Randomised shapes through the whole pass against the CPU oracle: class counts on both sides of every kernel switch
(register-resident / tensor-core scoring, 1..many row tiles of the gate kernel), ragged bags from one patch to a few
thousand, J below / at / above the bag size, K above the selected count, discarded classifiers, half masks.

Per case: bag logits of every slide within 1e-3 relative of `oracle.slide_eval_logits`, the selected index sets
identical up to rank-J ties, and for one masked slide the loss and all 33 092 gradients of a training step."""
import random

import numpy as np
import pytest
import torch

from moc_b200 import RaggedBagStore, _lib, ops, synthetic
from moc_b200.engine import MocEngine
from oracle import moc_oracle as O
from tests.helpers import assert_union_set, close

pytestmark = pytest.mark.gpu
DEV = "cuda"
DISCARDS = [(), ("delta_diff",), ("delta_softmax",), ("topk", "bottomk"), ("bottomk",)]


def _okeys(x, w, we, c):
    k = O.selection_keys(x, w, we, c)
    return np.concatenate([k["logit"].T, k["softmax"].T, k["delta"][None], k["bg_sum"][None], k["bg_max"][None]], 0)


@pytest.mark.parametrize("seed", range(16))
def test_random_cohort_matches_oracle(seed):
    rng = random.Random(1000 + seed)
    c = rng.choice([2, 2, 3, 4, 5, 8, 9, 12, 30])
    n_slides = rng.randint(1, 7)
    sizes = [rng.choice([1, 2, 7, 31, 127, 128, 129, 400, 401, rng.randint(3, 2500)]) for _ in range(n_slides)]
    j = rng.choice([1, 5, 64, 400])
    k = rng.choice([1, 3, 10, 16])
    disc = rng.choice(DISCARDS)
    w, we = synthetic.prompt_matrices(c, seed=seed + 5)
    bags, labels = synthetic.make_cohort(n_slides, sizes, c, cohort_seed=seed, w_ext=we)
    oprm = O.SenetParams.init(seed)
    prm = ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))
    store = RaggedBagStore.from_bags(bags, labels, DEV)
    eng = MocEngine(w.to(DEV), we.to(DEV), j, k, discard_classifiers=disc)
    got = eng.eval_logits(store, prm, check_domain=True)
    assert not eng.is_wide(store)
    for i, x in enumerate(bags):
        ref = O.slide_eval_logits(oprm, x, w, we, c, j, k, discard_classifiers=disc)
        close(got[i:i + 1], ref, rtol=1e-3, atol=2e-6)

    # one half-masked training step on the largest slide (selection inside the masked bag, CE, backward)
    i = max(range(n_slides), key=lambda t: sizes[t])
    n = sizes[i]
    mask = torch.rand(n, generator=torch.Generator().manual_seed(seed)) > 0.5
    if int(mask.sum()) == 0:
        mask[0] = True
    slide = O.slide_process(bags[i], w, we, c, j, discard_classifiers=disc, mask=mask)
    act = O.active_classifiers(disc, "train")
    loss_ref, logits_ref, grads_ref = O.head_forward_backward(oprm, slide, labels[i], k, act)
    flat = torch.empty(ops.NUM_PARAMS, device=DEV)
    out = eng.train_step(store, i, store.labels[i:i + 1], prm, mask.to(DEV), flat)
    lo, hi = store.offsets_h[i], store.offsets_h[i + 1]
    keys = eng.keys_for(store, i, i + 1)
    sel = ops.select_union(keys, torch.tensor([0, n], device=DEV), [0, n], c, j, _lib.discard_bits(disc), mask.to(DEV))
    got_idx = sel.sel_local[:int(sel.sel_count[0])].cpu().tolist()
    assert_union_set(got_idx, slide["selected_index"], _okeys(bags[i], w, we, c), c, j, disc, mask=mask.numpy())
    if got_idx == slide["selected_index"]:
        close(out.bag_logits, logits_ref, rtol=1e-3, atol=2e-6)
        assert abs(float(out.loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref)) + 1e-6
        gref = torch.cat([t.flatten() for t in grads_ref])
        assert (flat.cpu() - gref).abs().max().item() <= 1e-3 * gref.abs().max().item() + 1e-9
