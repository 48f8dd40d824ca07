"""CPU oracle for the MOC per-slide hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, in plain fp32 torch-CPU / numpy, the algorithm that
xmed-lab/MOC runs per slide (score -> four top-J patch selections -> union ->
meta-learner gate -> classifier-bank combination -> top-K pooling -> CE ->
backward -> Adam).  It exists so that the CUDA path in ``moc_b200`` can be
checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package never does (``tests/test_no_oracle_in_product.py`` enforces it).

Parity pin: the reference ships no golden vectors or tests of its own
(SURVEY.md section 4), so this oracle is pinned against outputs of the
reference's *own code* executed in the build container:
``oracle/make_golden.py`` imports ``utils/patch_selection_classifier*.py``
unmodified and AST-lifts ``senet/slide_process/train/evaluation/zs_evaluation``
out of ``main_moc.py``, runs them on seeded inputs and stores the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them through this
file.  The arithmetic underneath both is the same ATen CPU build (torch
2.11.0+cu128), so agreement is expected to be bit-exact for everything except
the hand-written backward/Adam, which is compared at 1e-6.

All ``file:line`` citations are relative to the reference checkout.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

TEMPERATURE_CONCH = 56.3477  # main_moc.py:175, :443, :505
CLASSIFIERS = ("topk", "delta_softmax", "delta_diff", "bottomk")  # main_moc.py:39


# --------------------------------------------------------------------------
# a2: scoring                                            main_moc.py:336-337
# --------------------------------------------------------------------------
def collapse_prompt_bank(bank: torch.Tensor, prompts_per_class: Sequence[int]) -> torch.Tensor:
    """utils/zeroshot_utils.py:29-50 from the text embeddings on: per class F.normalize every prompt embedding
    (:38), mean over classnames and templates (:43), divide by the norm (:44), stack as columns (:50)."""
    cols, p0 = [], 0
    for n in prompts_per_class:
        e = F.normalize(bank[p0:p0 + n].float(), dim=-1)
        m = e.mean(dim=0)
        cols.append(m / m.norm())
        p0 += n
    return torch.stack(cols, dim=1)


def score(feat: torch.Tensor, w: torch.Tensor, w_ext: torch.Tensor):
    """``L = feat @ W`` [N,C] and ``Le = feat @ W_ext`` [N,C_ext]; no normalisation."""
    return feat @ w, feat @ w_ext


# --------------------------------------------------------------------------
# a3..a6: the four index selectors      utils/patch_selection_classifier_index.py
# --------------------------------------------------------------------------
def _maxj(topj: Sequence[int], n_rows: int) -> int:
    return min(max(topj), n_rows)  # _index.py:24


def row_top1_minus_top2(logits: torch.Tensor) -> torch.Tensor:
    """|largest - second largest| of every row (needs C >= 2).  _index.py:46-48"""
    top2 = torch.topk(logits, 2, dim=1)[0]
    return torch.abs(top2[:, 0] - top2[:, 1])


def index_topj(logits, topj):
    """Per class column, rows of the maxj largest logits, descending.  _index.py:17-26"""
    return logits.topk(_maxj(topj, logits.size(0)), 0, True, True)[1]


def index_delta_softmax(logits, topj):
    """Same, on the row-softmax probabilities.  _index.py:28-36"""
    p = torch.softmax(logits, dim=1)
    return p.topk(_maxj(topj, logits.size(0)), 0, True, True)[1]


def index_delta_diff(logits, topj):
    """Rows with the largest |top1-top2| margin, replicated to C columns.  _index.py:38-51"""
    d = row_top1_minus_top2(logits)
    d = torch.stack([d] * logits.size(1), dim=1)
    return d.topk(_maxj(topj, logits.size(0)), 0, True, True)[1]


def index_bottomk_irrel(logits_ext, topj, n_classes, bottomk=None):
    """Rows least similar to the background prompts.  _index.py:53-87 (detection=False).

    The bottom-``bottomk`` rows of the summed background logits are taken, then
    re-ordered per foreground class; as a *set* this is just the bottom-maxj of
    the background sum.
    """
    assert n_classes is not None
    assert logits_ext.size(1) > n_classes
    maxj = _maxj(topj, logits_ext.size(0))
    if bottomk is None:
        bottomk = maxj
    bg = logits_ext[:, n_classes:].sum(dim=1)
    bottomk = min(bottomk, bg.size(0))
    bg_idx = bg.topk(bottomk, 0, False, True)[1]
    fg = logits_ext[:, :n_classes][bg_idx]
    fg_idx = fg.topk(maxj, 0, True, True)[1]
    return bg_idx[fg_idx]


# --------------------------------------------------------------------------
# a11: pooling                         utils/patch_selection_classifier.py
# --------------------------------------------------------------------------
def topj_pooling(logits, topj, return_indices=False):
    """Mean of the min(j, rows) largest values of every column.  :18-32"""
    maxj = _maxj(topj, logits.size(0))
    values, indices = logits.topk(maxj, 0, True, True)
    pooled = {j: values[: min(j, maxj)].mean(dim=0, keepdim=True) for j in topj}
    preds = {j: v.argmax(dim=1) for j, v in pooled.items()}
    return (preds, pooled, indices) if return_indices else (preds, pooled)


def delta_softmax_pooling(logits, topj, return_indices=False):
    """Select rows by softmax probability, pool their raw logits.  :35-53"""
    maxj = _maxj(topj, logits.size(0))
    idx = torch.softmax(logits, dim=1).topk(maxj, 0, True, True)[1]
    values = torch.gather(logits, 0, idx)
    pooled = {j: values[: min(j, maxj)].mean(dim=0, keepdim=True) for j in topj}
    preds = {j: v.argmax(dim=1) for j, v in pooled.items()}
    return (preds, pooled, idx) if return_indices else (preds, pooled)


def delta_diff_pooling(logits, topj, return_indices=False):
    """Select rows by |top1-top2|, pool their raw logit rows.  :56-78"""
    maxj = _maxj(topj, logits.size(0))
    d = row_top1_minus_top2(logits)
    d = torch.stack([d] * logits.size(1), dim=1)
    idx = d.topk(maxj, 0, True, True)[1]
    values = logits[idx[:, 0]]
    pooled = {j: values[: min(j, maxj)].mean(dim=0, keepdim=True) for j in topj}
    preds = {j: v.argmax(dim=1) for j, v in pooled.items()}
    return (preds, pooled, idx) if return_indices else (preds, pooled)


def bottomk_irrel_pooling(logits_ext, topj, coords_list, return_indices=False, bottomk=None):
    """Pool foreground logits over the rows least like background.  :127-171"""
    n_fg = coords_list if isinstance(coords_list, int) else len(coords_list)
    assert logits_ext.size(1) > n_fg
    maxj = _maxj(topj, logits_ext.size(0))
    if bottomk is None:
        bottomk = maxj
    bg = logits_ext[:, n_fg:].sum(dim=1)
    bg_idx = bg.topk(bottomk, 0, False, True)[1]
    fg_values, fg_idx = logits_ext[:, :n_fg][bg_idx].topk(maxj, 0, True, True)
    pooled = {j: fg_values[: min(j, maxj)].mean(dim=0, keepdim=True) for j in topj}
    preds = {j: v.argmax(dim=1) for j, v in pooled.items()}
    idx = bg_idx[fg_idx]
    return (preds, pooled, idx) if return_indices else (preds, pooled)


# --------------------------------------------------------------------------
# a7, a8, a13: slide_process                              main_moc.py:322-375
# --------------------------------------------------------------------------
def slide_process(feat, w, w_ext, n_classes, topj=10, random_mask=False,
                  discard_classifiers=(), mask: Optional[torch.Tensor] = None) -> dict:
    """Select the informative patches of one bag and build the four score planes.

    ``mask`` (bool [N]) is an addition: when given it replaces the reference's
    ``torch.rand(N) > 0.5`` draw (:329-331) so both sides of a parity test can be
    fed the same rows.  With ``random_mask=True`` and no ``mask`` the draw is
    made exactly as the reference does, on the CPU default generator.
    """
    if mask is None and random_mask:
        mask = torch.rand(feat.size(0)) > 0.5
    if mask is not None:
        feat = feat[mask]
    logits, logits_ext = score(feat, w, w_ext)
    tj = [topj]
    chosen = set()
    if "topk" not in discard_classifiers:
        chosen.update(index_topj(logits, tj).flatten().tolist())
    if "delta_softmax" not in discard_classifiers:
        chosen.update(index_delta_softmax(logits, tj).flatten().tolist())
    if "delta_diff" not in discard_classifiers:
        chosen.update(index_delta_diff(logits, tj).flatten().tolist())
    if "bottomk" not in discard_classifiers:
        chosen.update(index_bottomk_irrel(logits_ext, tj, n_classes).flatten().tolist())
    sel = sorted(chosen)
    sel_feat = feat[sel]
    sel_logits, sel_logits_ext = score(sel_feat, w, w_ext)
    c = sel_logits.size(1)
    diff = row_top1_minus_top2(sel_logits)
    bg = sel_logits_ext[:, n_classes:].max(dim=1)[0]  # note: max here, sum in the selector
    return {
        "selected_index": sel,
        "selected_feat": sel_feat,
        "logits_top_classifier": sel_logits,
        "logits_delta_softmax_classifier": sel_logits.softmax(dim=1),
        "logits_delta_diff_classifier": torch.stack([diff] * c, dim=1),
        "logits_bottomk_irrel_classifier": torch.stack([bg] * c, dim=1),
    }


# --------------------------------------------------------------------------
# a9: the meta-learner "senet"                            main_moc.py:299-312
# --------------------------------------------------------------------------
@dataclass
class SenetParams:
    """512 -> 64 -> 4 MLP; names follow the reference state_dict keys."""
    w1: torch.Tensor  # model.0.weight [64,512]
    b1: torch.Tensor  # model.0.bias   [64]
    w2: torch.Tensor  # model.2.weight [4,64]
    b2: torch.Tensor  # model.2.bias   [4]

    def tensors(self) -> List[torch.Tensor]:
        return [self.w1, self.b1, self.w2, self.b2]

    def clone(self) -> "SenetParams":
        return SenetParams(*[t.clone() for t in self.tensors()])

    @staticmethod
    def from_state_dict(sd) -> "SenetParams":
        return SenetParams(sd["model.0.weight"].detach().clone().float(),
                           sd["model.0.bias"].detach().clone().float(),
                           sd["model.2.weight"].detach().clone().float(),
                           sd["model.2.bias"].detach().clone().float())

    def state_dict(self):
        return {"model.0.weight": self.w1, "model.0.bias": self.b1,
                "model.2.weight": self.w2, "model.2.bias": self.b2}

    @staticmethod
    def init(seed: int, in_dim: int = 512, hidden: int = 64, out_dim: int = 4) -> "SenetParams":
        """Seeded ``nn.Linear`` default init (kaiming-uniform a=sqrt(5) == U(+-1/sqrt(fan_in)))."""
        g = torch.Generator().manual_seed(seed)

        def u(shape, fan_in):
            bound = 1.0 / math.sqrt(fan_in)
            return (torch.rand(shape, generator=g) * 2 - 1) * bound

        return SenetParams(u((hidden, in_dim), in_dim), u((hidden,), in_dim),
                           u((out_dim, hidden), hidden), u((out_dim,), hidden))


def senet_forward(p: SenetParams, x: torch.Tensor):
    """Returns (gate [S,4] in (0,1), hidden [S,64] after ReLU)."""
    h = torch.relu(x @ p.w1.t() + p.b1)
    g = torch.sigmoid(h @ p.w2.t() + p.b2)
    return g, h


# --------------------------------------------------------------------------
# a10: classifier-bank combination        main_moc.py:391-403 (train), :482-492 (eval)
# --------------------------------------------------------------------------
_PLANES = ("logits_top_classifier", "logits_delta_softmax_classifier",
           "logits_delta_diff_classifier", "logits_bottomk_irrel_classifier")


def active_classifiers(discard_classifiers=(), mode="train") -> Tuple[bool, bool, bool, bool]:
    """Which of the four gated planes enter the sum.

    train (:396-403) honours all four discard names.  evaluation (:486-492)
    always keeps the top-k plane and tests the never-matching name
    "delta_bottomk", so the bottom-k plane is always kept as well.
    """
    d = set(discard_classifiers)
    if mode == "train":
        return ("topk" not in d, "delta_softmax" not in d, "delta_diff" not in d, "bottomk" not in d)
    return (True, "delta_softmax" not in d, "delta_diff" not in d, "delta_bottomk" not in d)


def combine(gate, slide, active=(True, True, True, True)):
    out = torch.zeros_like(slide[_PLANES[0]])
    for m in range(4):
        if active[m]:
            out = out + gate[:, m:m + 1] * slide[_PLANES[m]]
    return out


def bag_logits(final_logits, topk: int):
    """[S,C] -> [1,C]: per-class mean of the min(K,S) largest.  main_moc.py:405,:493"""
    return topj_pooling(final_logits, [topk])[1][topk]


def cross_entropy(logits, label: int) -> torch.Tensor:
    """CE of one [1,C] row, no temperature.  main_moc.py:406"""
    ls = torch.log_softmax(logits, dim=1)
    return -ls[0, int(label)]


# --------------------------------------------------------------------------
# a12: loss, hand-written backward, Adam                 main_moc.py:406-410, :316
# --------------------------------------------------------------------------
@dataclass
class AdamState:
    """torch.optim.Adam(lr=1e-3, weight_decay=1e-4): L2 folded into the gradient."""
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 1e-4
    step: int = 0
    m: List[torch.Tensor] = field(default_factory=list)
    v: List[torch.Tensor] = field(default_factory=list)


def head_forward_backward(p: SenetParams, slide: dict, label: int, topk: int,
                          active=(True, True, True, True)):
    """Forward + closed-form backward of one training step.  Returns (loss, logits, grads).

    d loss/d F[s,c] = (softmax(logit)_c - [c==y]) / k_eff  for rows in column c's
    top-k_eff, zero elsewhere; everything upstream of the gate is constant.
    """
    x = slide["selected_feat"]
    z1 = x @ p.w1.t() + p.b1
    h = torch.relu(z1)
    g = torch.sigmoid(h @ p.w2.t() + p.b2)
    f = combine(g, slide, active)
    k_eff = min(topk, f.size(0))
    vals, idx = f.topk(k_eff, 0, True, True)
    logits = vals.mean(dim=0, keepdim=True)
    loss = cross_entropy(logits, label)
    dlogit = torch.softmax(logits, dim=1)[0].clone()
    dlogit[int(label)] -= 1.0
    df = torch.zeros_like(f)
    cols = torch.arange(f.size(1)).unsqueeze(0).expand_as(idx)
    df[idx, cols] = (dlogit / k_eff).unsqueeze(0).expand_as(idx)
    dg = torch.zeros_like(g)
    for m in range(4):
        if active[m]:
            dg[:, m] = (df * slide[_PLANES[m]]).sum(dim=1)
    dz2 = dg * g * (1.0 - g)
    dw2 = dz2.t() @ h
    db2 = dz2.sum(dim=0)
    dz1 = (dz2 @ p.w2) * (z1 > 0).to(x.dtype)
    dw1 = dz1.t() @ x
    db1 = dz1.sum(dim=0)
    return loss, logits, [dw1, db1, dw2, db2]


def adam_step(p: SenetParams, grads: List[torch.Tensor], st: AdamState) -> None:
    """In-place single-tensor Adam exactly as torch/optim/adam.py (no amsgrad, no maximize)."""
    if not st.m:
        st.m = [torch.zeros_like(t) for t in p.tensors()]
        st.v = [torch.zeros_like(t) for t in p.tensors()]
    st.step += 1
    bc1 = 1.0 - st.beta1 ** st.step
    bc2 = 1.0 - st.beta2 ** st.step
    step_size = st.lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for t, g, m, v in zip(p.tensors(), grads, st.m, st.v):
        g = g + st.weight_decay * t
        m.mul_(st.beta1).add_(g, alpha=1.0 - st.beta1)
        v.mul_(st.beta2).addcmul_(g, g, value=1.0 - st.beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(st.eps)
        t.addcdiv_(m, denom, value=-step_size)


# --------------------------------------------------------------------------
# a1 surface + a14 loops            datasets/dataset_generic.py:380-433; main_moc.py:378-520
# --------------------------------------------------------------------------
class BagList:
    """The slice of ``Generic_Split`` the loops use: real_len(), repeat_num, len, items."""

    def __init__(self, bags: List[torch.Tensor], labels: Sequence[int], repeat_num: Optional[int] = None):
        self.bags = bags
        self.labels = [int(v) for v in labels]
        self.repeat_num = repeat_num

    def real_len(self) -> int:
        return len(self.bags)

    def __len__(self) -> int:
        return self.repeat_num if self.repeat_num else len(self.bags)

    def __getitem__(self, idx: int):
        if idx >= len(self):
            raise IndexError
        i = idx % len(self.bags)
        return self.bags[i], self.labels[i]


def train_epoch(p: SenetParams, st: AdamState, data: BagList, w, w_ext, n_classes, topj, topk,
                discard_classifiers=(), masks: Optional[Iterable[torch.Tensor]] = None,
                dp_microbatch: Optional[int] = None) -> List[float]:
    """One pass of main_moc.py:378-410: one Adam step per (virtual) slide, half-masked.

    ``dp_microbatch=G`` is NOT the reference's semantics: it is the checker for the data-parallel training mode
     - the gradients of G consecutive slides, all taken at the same parameters, are summed in
    slide order and applied with one Adam step (a last, shorter micro-batch when G does not divide the epoch)."""
    losses = []
    masks = iter(masks) if masks is not None else None
    act = active_classifiers(discard_classifiers, "train")
    g = int(dp_microbatch) if dp_microbatch else 1
    acc, pending = None, 0
    for i in range(len(data)):
        feat, lbl = data[i]
        mk = next(masks) if masks is not None else None
        slide = slide_process(feat, w, w_ext, n_classes, topj, random_mask=True,
                              discard_classifiers=discard_classifiers, mask=mk)
        loss, _, grads = head_forward_backward(p, slide, lbl, topk, act)
        losses.append(float(loss))
        acc = grads if acc is None else [a + b for a, b in zip(acc, grads)]
        pending += 1
        if pending == g or i == len(data) - 1:
            adam_step(p, acc, st)
            acc, pending = None, 0
    return losses


def _metrics(logits_all: torch.Tensor, labels: Sequence[int], loss_sum: float, loss_div: int, real_len: int):
    """acc / AUC exactly as main_moc.py:499-520 (softmax at T=56.3477; binary uses column 1)."""
    from sklearn.metrics import roc_auc_score
    y = np.asarray(labels)
    correct = int((logits_all.argmax(dim=1).numpy() == y).sum())
    probs = torch.softmax(logits_all * TEMPERATURE_CONCH, dim=1)
    if probs.shape[1] == 2:
        auc = roc_auc_score(y, probs[:, 1].numpy())
    else:
        auc = roc_auc_score(y, probs.numpy(), multi_class="ovo", average="macro")
    return {"loss": loss_sum / loss_div, "acc": correct / real_len, "auc": auc}


def slide_eval_logits(p: SenetParams, feat, w, w_ext, n_classes, topj, topk, discard_classifiers=()):
    slide = slide_process(feat, w, w_ext, n_classes, topj, discard_classifiers=discard_classifiers)
    g, _ = senet_forward(p, slide["selected_feat"])
    f = combine(g, slide, active_classifiers(discard_classifiers, "eval"))
    return bag_logits(f, topk)


def evaluation(p: SenetParams, data: BagList, w, w_ext, n_classes, topj, topk, discard_classifiers=(),
               return_logits=False):
    """main_moc.py:462-520: every slide once; loss divided by len(dataset) *after* repeat_num is restored."""
    rows, labels, loss_sum = [], [], 0.0
    for i in range(data.real_len()):
        feat, lbl = data.bags[i], data.labels[i]
        lg = slide_eval_logits(p, feat, w, w_ext, n_classes, topj, topk, discard_classifiers)
        loss_sum += float(cross_entropy(lg, lbl))
        rows.append(lg)
        labels.append(lbl)
    out = _metrics(torch.cat(rows, 0), labels, loss_sum, len(data), data.real_len())
    return (out, torch.cat(rows, 0)) if return_logits else out


def zs_evaluation(data: BagList, w, w_ext, n_classes, topk, pooling="topj", return_logits=False):
    """main_moc.py:412-460 with pooling_func in {topj, delta_softmax, delta_diff, bottomk_irrel}."""
    rows, labels, loss_sum = [], [], 0.0
    for i in range(data.real_len()):
        feat, lbl = data.bags[i], data.labels[i]
        lo, le = score(feat, w, w_ext)
        if pooling == "topj":
            lg = topj_pooling(lo, [topk])[1][topk]
        elif pooling == "delta_softmax":
            lg = delta_softmax_pooling(lo, [topk])[1][topk]
        elif pooling == "delta_diff":
            lg = delta_diff_pooling(lo, [topk])[1][topk]
        elif pooling == "bottomk_irrel":
            lg = bottomk_irrel_pooling(le, [topk], coords_list=n_classes)[1][topk]
        else:
            raise ValueError(pooling)
        loss_sum += float(cross_entropy(lg, lbl))
        rows.append(lg)
        labels.append(lbl)
    out = _metrics(torch.cat(rows, 0), labels, loss_sum, len(data), data.real_len())
    return (out, torch.cat(rows, 0)) if return_logits else out


def ablation_logits(feat, w, w_ext, n_classes, topj, topk, how: str):
    """main_moc.py:537-555: un-gated avg / sum / max of the four planes."""
    s = slide_process(feat, w, w_ext, n_classes, topj)
    planes = [s[k] for k in _PLANES]
    if how == "avg":
        f = 0.25 * planes[0] + 0.25 * planes[1] + 0.25 * planes[2] + 0.25 * planes[3]
    elif how == "sum":
        f = planes[0] + planes[1] + planes[2] + planes[3]
    elif how == "max":
        f = torch.stack(planes, dim=0).max(dim=0)[0]
    else:
        raise ValueError(how)
    return bag_logits(f, topk)


# --------------------------------------------------------------------------
# Order-free set semantics (float64 numpy) used to judge ties in parity tests
# --------------------------------------------------------------------------
def selection_keys(feat: torch.Tensor, w, w_ext, n_classes: int) -> Dict[str, np.ndarray]:
    """The per-row keys all selections are made on, in float64 from fp32 scores."""
    lo, le = score(feat, w, w_ext)
    lo64 = lo.double().numpy()
    srt = np.sort(lo64, axis=1)
    return {
        "logit": lo.numpy(),
        "softmax": torch.softmax(lo, dim=1).numpy(),
        "delta": np.abs(srt[:, -1] - srt[:, -2]).astype(np.float32),
        "bg_sum": le[:, n_classes:].sum(dim=1).numpy(),
        "bg_max": le[:, n_classes:].max(dim=1)[0].numpy(),
    }


def rank_threshold(values: np.ndarray, j: int, largest: bool = True) -> float:
    """The value at rank j (1-based) - rows strictly beyond it are in every valid top-j set."""
    v = np.sort(values)
    return float(v[-j] if largest else v[j - 1])
