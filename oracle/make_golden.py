"""Generate tests/golden/*.npz from the UNMODIFIED reference  --  TEST INFRASTRUCTURE.

Run in the build container only (needs the read-only checkout at /root/reference):

    python oracle/make_golden.py

It executes the reference's own selector / pooling modules and the functions
lifted from ``main_moc.py`` (see ``oracle/ref_loader.py``) on small seeded
inputs and stores inputs *and* outputs, so the fixtures are self-contained on
the GPU box where the checkout does not exist.  The reference has no golden
vectors of its own; these files are the parity pin for ``oracle/moc_oracle.py``
and, through it, for the CUDA path.
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

OUT = os.path.join(ROOT, "tests", "golden")


class _RecordingF:
    """``torch.nn.functional`` with every cross_entropy value recorded."""

    def __init__(self):
        import torch.nn.functional as F
        self._F = F
        self.losses = []

    def __getattr__(self, k):
        return getattr(self._F, k)

    def cross_entropy(self, *a, **kw):
        v = self._F.cross_entropy(*a, **kw)
        self.losses.append(float(v.detach()))
        return v


def _args(c, j, k, discard=()):
    return types.SimpleNamespace(disable_tqdm=True, n_classes=c, topj=j, topk=k,
                                 discard_classifiers=list(discard), pretrain="conch", ablation_study="none")


def _fp16_exact(cohort):
    """Inputs are rounded to fp16-representable fp32 values *before* the reference sees them, so the
    fixtures can store them in half the bytes without changing what was computed."""
    bags, labels = cohort
    return [b.half().float() for b in bags], labels


def _sd_np(sd, prefix):
    return {prefix + k.replace(".", "_"): v.detach().numpy().copy() for k, v in sd.items()}


def slide_case(ref, name, c, sizes, j, k, seed, discards=((),)):
    """Per-slide outputs of selectors, slide_process, poolers, senet gate and bag logits."""
    w, w_ext = synthetic.prompt_matrices(c)
    bags, labels = _fp16_exact(synthetic.make_cohort(len(sizes), sizes, c, cohort_seed=seed))
    torch.manual_seed(seed + 101)
    model = ref.senet(512, 4)
    out = {"C": c, "J": j, "K": k, "W": w.numpy(), "W_ext": w_ext.numpy(), "n_slides": len(bags),
           "labels": np.asarray(labels), "discards": np.asarray(["|".join(d) for d in discards])}
    out.update(_sd_np(model.state_dict(), "sd_"))
    with torch.no_grad():
        for i, x in enumerate(bags):
            p = "s%d_" % i
            out[p + "feat"] = x.half().numpy()
            lo, le = x @ w, x @ w_ext
            out[p + "L"], out[p + "Le"] = lo.numpy(), le.numpy()
            out[p + "idx_topj"] = ref.index.index_topj_classifier(lo, [j]).numpy()
            out[p + "idx_dsoftmax"] = ref.index.index_delta_softmax_classifier(lo, [j]).numpy()
            out[p + "idx_ddiff"] = ref.index.index_delta_diff_classifier(lo, [j]).numpy()
            out[p + "idx_bottomk"] = ref.index.index_bottomk_irrel_classifier(le, [j], c).numpy()
            out[p + "pool_topj"] = ref.pool.topj_pooling(lo, [k])[1][k].numpy()
            out[p + "pool_dsoftmax"] = ref.pool.delta_softmax_classifier_pooling(lo, [k])[1][k].numpy()
            out[p + "pool_ddiff"] = ref.pool.delta_diff_classifier_pooling(lo, [k])[1][k].numpy()
            out[p + "pool_bottomk"] = ref.pool.bottomk_irrel_classifier_pooling(le, [k], coords_list=c)[1][k].numpy()
            for di, disc in enumerate(discards):
                q = p + "d%d_" % di
                r = ref.slide_process(x, w, w_ext, n_classes=c, topj=j, discard_classifiers=list(disc))
                out[q + "selected_index"] = np.asarray(r["selected_index"], dtype=np.int64)
                out[q + "plane_top"] = r["logits_top_classifier"].numpy()
                out[q + "plane_dsoftmax"] = r["logits_delta_softmax_classifier"].numpy()
                out[q + "plane_ddiff"] = r["logits_delta_diff_classifier"].numpy()
                out[q + "plane_bottomk"] = r["logits_bottomk_irrel_classifier"].numpy()
                gate = model(r["selected_feat"])
                out[q + "gate"] = gate.numpy()
                # eval-mode combination exactly as main_moc.py:482-493
                f = gate[:, 0:1] * r["logits_top_classifier"]
                if "delta_softmax" not in disc:
                    f = f + gate[:, 1:2] * r["logits_delta_softmax_classifier"]
                if "delta_diff" not in disc:
                    f = f + gate[:, 2:3] * r["logits_delta_diff_classifier"]
                f = f + gate[:, 3:4] * r["logits_bottomk_irrel_classifier"]
                out[q + "final"] = f.numpy()
                out[q + "bag_logits"] = ref.pool.topj_pooling(f, [k])[1][k].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "slides", len(bags))


def loop_case(ref, name, c, train_sizes, val_sizes, j, k, seed, repeat_num, epochs=2, discard=()):
    """reference train() / evaluation() / zs_evaluation() / ablation_evaluation() on a tiny cohort."""
    import torch.nn as nn  # noqa: F401
    w, w_ext = synthetic.prompt_matrices(c)
    ref.set_weights(w, w_ext)
    tr_bags, tr_lab = _fp16_exact(synthetic.make_cohort(len(train_sizes), train_sizes, c, cohort_seed=seed))
    va_bags, va_lab = _fp16_exact(synthetic.make_cohort(len(val_sizes), val_sizes, c, cohort_seed=seed + 1))
    args = _args(c, j, k, discard)
    recF = _RecordingF()
    for fn in (ref.train, ref.evaluation, ref.zs_evaluation, ref.ablation_evaluation):
        fn.__globals__["F"] = recF

    torch.manual_seed(seed + 11)
    model = ref.senet(512, 4)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    out = {"C": c, "J": j, "K": k, "W": w.numpy(), "W_ext": w_ext.numpy(), "repeat_num": repeat_num,
           "epochs": epochs, "discard": "|".join(discard),
           "train_labels": np.asarray(tr_lab), "val_labels": np.asarray(va_lab),
           "n_train": len(tr_bags), "n_val": len(va_bags)}
    for i, x in enumerate(tr_bags):
        out["train_feat_%d" % i] = x.half().numpy()
    for i, x in enumerate(va_bags):
        out["val_feat_%d" % i] = x.half().numpy()
    out.update(_sd_np(model.state_dict(), "sd0_"))

    tr_loader = ref_loader.RefLoader(ref_loader.RefDataset(tr_bags, tr_lab, repeat_num=repeat_num))
    va_loader = ref_loader.RefLoader(ref_loader.RefDataset(va_bags, va_lab, repeat_num=None))

    def ev(d):
        return np.asarray([d["loss"], d["acc"], d["auc"]], dtype=np.float64)

    recF.losses.clear()
    out["zs_train"] = ev(ref.zs_evaluation(tr_loader, "cpu", args))
    out["zs_val"] = ev(ref.zs_evaluation(va_loader, "cpu", args))
    out["zs_val_dsoftmax"] = ev(ref.zs_evaluation(va_loader, "cpu", args,
                                                  pooling_func=ref.pool.delta_softmax_classifier_pooling))
    out["zs_val_ddiff"] = ev(ref.zs_evaluation(va_loader, "cpu", args,
                                               pooling_func=ref.pool.delta_diff_classifier_pooling))
    out["zs_val_bottomk"] = ev(ref.zs_evaluation(va_loader, "cpu", args,
                                                 pooling_func=ref.pool.bottomk_irrel_classifier_pooling))
    for how in ("avg", "sum", "max"):
        args.ablation_study = how
        out["ablation_val_" + how] = ev(ref.ablation_evaluation(va_loader, "cpu", args))
    args.ablation_study = "none"

    # the masks train() is about to draw: same generator, same order (RefLoader draws nothing)
    mask_seed = seed + 12
    torch.manual_seed(mask_seed)
    masks = []
    for _ in range(epochs):
        for kstep in range(repeat_num):
            n = tr_bags[kstep % len(tr_bags)].size(0)
            masks.append((torch.rand(n) > 0.5).numpy())
    torch.manual_seed(mask_seed)
    step = 0
    for e in range(epochs):
        recF.losses.clear()
        ref.train(model, tr_loader, opt, "cpu", args)
        out["train_losses_e%d" % e] = np.asarray(recF.losses, dtype=np.float64)
        out.update(_sd_np(model.state_dict(), "sd_e%d_" % e))
        state = torch.get_rng_state()  # evaluation must not disturb the mask stream
        out["eval_train_e%d" % e] = ev(ref.evaluation(model, tr_loader, "cpu", args))
        recF.losses.clear()
        out["eval_val_e%d" % e] = ev(ref.evaluation(model, va_loader, "cpu", args))
        out["eval_val_losses_e%d" % e] = np.asarray(recF.losses, dtype=np.float64)
        torch.set_rng_state(state)
        step += repeat_num
    for i, m in enumerate(masks):
        out["mask_%d" % i] = m
    out["n_masks"] = len(masks)
    for gi, p_ in enumerate(model.parameters()):
        out["adam_m_%d" % gi] = opt.state[p_]["exp_avg"].numpy().copy()
        out["adam_v_%d" % gi] = opt.state[p_]["exp_avg_sq"].numpy().copy()
    out["adam_step"] = int(opt.state[next(model.parameters())]["step"])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def main():
    assert ref_loader.available(), "reference checkout not found at %s" % ref_loader.REFERENCE_ROOT
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = ref_loader.load()
    # NSCLC-shaped (C=2, C_ext=6): N<J, N~J, N>J; the shipped J=400 and a smaller J
    slide_case(ref, "slide_c2", 2, [37, 420, 640], 400, 10, seed=3,
               discards=((), ("delta_diff",), ("topk", "bottomk")))
    slide_case(ref, "slide_c2_j64", 2, [5, 600], 64, 10, seed=4)
    # RCC-shaped (C=3, C_ext=7)
    slide_case(ref, "slide_c3", 3, [50, 500], 100, 10, seed=5, discards=((), ("delta_softmax",)))
    # EBRAINS-30-shaped (C=30, C_ext=34)
    slide_case(ref, "slide_c30", 30, [400], 40, 10, seed=6)
    # few-shot loops
    loop_case(ref, "loop_c2", 2, [180, 150, 200, 170], [120, 140, 130, 110, 150, 160], 60, 10, seed=7,
              repeat_num=6, epochs=2)
    loop_case(ref, "loop_c3_discard", 3, [120, 140, 160], [100, 110, 90, 120, 130, 105], 40, 10, seed=8,
              repeat_num=3, epochs=1, discard=("delta_diff",))


if __name__ == "__main__":
    main()
