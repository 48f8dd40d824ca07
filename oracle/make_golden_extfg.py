"""Golden for prompt files whose W_ext[:, :C] differs from W  --  TEST INFRASTRUCTURE.

    python oracle/make_golden_extfg.py        (build container only: needs /root/reference)

The reference builds `zeroshot_weights` and `zeroshot_weights_ext` from two different prompt files
(main_moc.py:163-197); in the shipped files the class columns coincide, but the code treats the tensors as independent.
`zs_evaluation(pooling_func=bottomk_irrel_classifier_pooling)` pools (feats @ W_ext)[:, :C] (main_moc.py:428-432,
utils/patch_selection_classifier.py:127-171) while everything else scores classes against W.  This fixture runs the
reference's own `zs_evaluation` (all four pooling functions) and `evaluation` with class columns of W_ext that are
NOT W, and stores per-slide pooled logits as well.  Writes tests/golden/zs_extfg.npz.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from moc_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402


def main():
    assert ref_loader.available()
    torch.set_num_threads(1)
    ref = ref_loader.load()
    c, j, k, seed = 3, 48, 10, 21
    w, w_ext = synthetic.prompt_matrices(c)
    g = torch.Generator().manual_seed(seed)
    fg = w + 0.35 * torch.randn(512, c, generator=g) / 512 ** 0.5      # the "other prompt file": near W, not W
    w_ext = w_ext.clone()
    w_ext[:, :c] = fg / fg.norm(dim=0, keepdim=True)
    ref.set_weights(w, w_ext)
    sizes = [150, 90, 210, 120, 170, 60]
    bags, labels = synthetic.make_cohort(len(sizes), sizes, c, cohort_seed=seed)
    bags = [b.half().float() for b in bags]
    args = types.SimpleNamespace(disable_tqdm=True, n_classes=c, topj=j, topk=k, discard_classifiers=[],
                                 pretrain="conch", ablation_study="none")
    loader = ref_loader.RefLoader(ref_loader.RefDataset(bags, labels))
    ev = lambda d: np.asarray([d["loss"], d["acc"], d["auc"]], dtype=np.float64)
    out = {"C": c, "J": j, "K": k, "W": w.numpy(), "W_ext": w_ext.numpy(), "labels": np.asarray(labels),
           "n_slides": len(bags)}
    for i, x in enumerate(bags):
        out["feat_%d" % i] = x.half().numpy()
    pools = {"topj": ref.pool.topj_pooling, "dsoftmax": ref.pool.delta_softmax_classifier_pooling,
             "ddiff": ref.pool.delta_diff_classifier_pooling, "bottomk": ref.pool.bottomk_irrel_classifier_pooling}
    for name, fn in pools.items():
        out["zs_" + name] = ev(ref.zs_evaluation(loader, "cpu", args, pooling_func=fn))
    # per-slide pooled logits of the bottomk_irrel mode, as zs_evaluation computes them (main_moc.py:427-432)
    out["bottomk_logits"] = np.concatenate(
        [ref.pool.bottomk_irrel_classifier_pooling(x @ w_ext, [k], coords_list=c)[1][k].numpy() for x in bags], 0)
    torch.manual_seed(seed + 1)
    model = ref.senet(512, 4)
    for kk, v in model.state_dict().items():
        out["sd_" + kk.replace(".", "_")] = v.numpy().copy()
    out["eval"] = ev(ref.evaluation(model, loader, "cpu", args))
    path = os.path.join(ROOT, "tests", "golden", "zs_extfg.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {n: out["zs_" + n].tolist() for n in pools})


if __name__ == "__main__":
    main()
