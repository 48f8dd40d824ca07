"""train / evaluation / zs_evaluation / ablation_evaluation / main with the reference's signatures
(main_moc.py:378-644), driving the CUDA kernels over GPU-resident bags.

As in the reference the two prompt matrices are module globals (``zeroshot_weights``,
``zeroshot_weights_ext``; main_moc.py:386,:427-428,:478,:537) - set them with :func:`set_prompts`.
``loader`` is anything with a ``.dataset`` exposing ``real_len()``, ``repeat_num`` and ``len()``.  When the
dataset is a :class:`moc_b200.bag_store.BagDataset` the whole split is processed in a few launches; any other
loader yielding the reference's ``(feats, lbl, coords, full_path)`` tuples is consumed slide by slide.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib, ops
from .bag_store import BagDataset, BagLoader, RaggedBagStore
from .engine import MocEngine
from .pooling import (bottomk_irrel_classifier_pooling, delta_diff_classifier_pooling,
                      delta_softmax_classifier_pooling, topj_pooling)

TEMPERATURE = {"conch": 56.3477}  # main_moc.py:443,:505

zeroshot_weights: Optional[torch.Tensor] = None
zeroshot_weights_ext: Optional[torch.Tensor] = None
_ENGINES = {}
_POOL_NAMES = {topj_pooling: "topj", delta_softmax_classifier_pooling: "delta_softmax",
               delta_diff_classifier_pooling: "delta_diff", bottomk_irrel_classifier_pooling: "bottomk_irrel"}


def set_prompts(w: torch.Tensor, w_ext: torch.Tensor) -> None:
    global zeroshot_weights, zeroshot_weights_ext
    zeroshot_weights, zeroshot_weights_ext = w, w_ext
    _ENGINES.clear()


def engine_for(args) -> MocEngine:
    if zeroshot_weights is None or zeroshot_weights_ext is None:
        raise _lib.MocError(_lib.E_ARG, "call moc_b200.loops.set_prompts(zeroshot_weights, zeroshot_weights_ext) first")
    key = (zeroshot_weights.data_ptr(), zeroshot_weights_ext.data_ptr(), zeroshot_weights._version,
           zeroshot_weights_ext._version, int(args.topj), int(args.topk),
           tuple(getattr(args, "discard_classifiers", ()) or ()), bool(getattr(args, "cache_scores", False)))
    eng = _ENGINES.get(key)
    if eng is None:
        eng = MocEngine(zeroshot_weights, zeroshot_weights_ext, args.topj, args.topk,
                        getattr(args, "discard_classifiers", ()), cache_scores=getattr(args, "cache_scores", False))
        _ENGINES[key] = eng
    return eng


def _loader_iter_draw() -> None:
    """What ``iter(DataLoader)`` does to the CPU default generator: one int64 draw for the iterator's base seed
    (torch/utils/data/dataloader.py, _BaseDataLoaderIter.__init__; the same for num_workers 0 and 1).  Every loop of the
    reference starts one iterator (``for ... in tqdm(loader)``, main_moc.py:380,:421,:472,:531), and train()'s half masks
    come from the same generator (:330) - so a SEEDED reference run and a seeded run of these loops only draw the same
    masks if the iterator's draw is mirrored, once per call, in the same place."""
    torch.empty((), dtype=torch.int64).random_()


def _store_of(loader, device) -> RaggedBagStore:
    """The split as a resident store: directly, or by draining a generic loader once (cached on the dataset).
    Consumes the CPU generator exactly as one ``iter(DataLoader)`` of the reference does."""
    ds = loader.dataset
    st = getattr(ds, "store", None)
    if st is not None:
        _loader_iter_draw()
        return st
    st = getattr(ds, "_moc_b200_store", None)
    if st is not None:
        _loader_iter_draw()
    if st is None:          # iterating the caller's own loader below makes the draw itself (if it is a DataLoader)
        keep = ds.repeat_num
        ds.repeat_num = ds.real_len()
        bags, labels, ids = [], [], []
        for feats, lbl, coords, full_path in loader:
            bags.append(feats.squeeze(0))
            labels.append(int(lbl))
            ids.append(full_path[0])
        ds.repeat_num = keep
        st = RaggedBagStore.from_bags(bags, labels, device, ids)
        ds._moc_b200_store = st
    return st


def _flat_params(model):
    return model.parameters_in_order() if hasattr(model, "parameters_in_order") else list(model.parameters())


def _is_plain_adam(optimizer) -> bool:
    """torch.optim.Adam exactly as main_moc.py:316 builds it (L2 weight decay in the gradient, no amsgrad / maximize /
    capturable / fused / decoupled decay, python-float lr, no step hooks): what our Adam kernels implement."""
    return (type(optimizer) is torch.optim.Adam and all(
        not g.get("amsgrad") and not g.get("maximize") and not g.get("capturable") and not g.get("fused")
        and not g.get("differentiable") and not g.get("decoupled_weight_decay")
        and not isinstance(g["lr"], torch.Tensor) for g in optimizer.param_groups)
        and not getattr(optimizer, "_optimizer_step_pre_hooks", None)
        and not getattr(optimizer, "_optimizer_step_post_hooks", None))


def _adam_dev(model, optimizer):
    """The optimizer's device-side mirror for graph-captured steps (one per optimizer, rebuilt if it was stepped
    behind our back or its hyper-parameters changed)."""
    from .engine import AdamDev
    params = _flat_params(model)
    ad = getattr(optimizer, "_moc_adam_dev", None)
    if ad is None or not ad.in_sync() or [id(p) for p in ad.params] != [id(p) for p in params]:
        ad = optimizer._moc_adam_dev = AdamDev(optimizer, params)
    return ad


def _optimizer_step(model, optimizer, flat_grads: torch.Tensor) -> None:
    """optimizer.step() on the flat gradient.  A plain torch.optim.Adam (what main_moc.py:316 builds) is
    advanced by our Adam kernel directly on its own state tensors; anything else gets .grad and its own step()."""
    params = _flat_params(model)
    views = ops.split_grads(flat_grads)
    plain_adam = _is_plain_adam(optimizer)
    if not plain_adam:
        for p, g in zip(params, views):
            p.grad = g.clone()
        optimizer.step()
        return
    group_of = {id(p): g for g in optimizer.param_groups for p in g["params"]}
    for p, g in zip(params, views):
        grp = group_of[id(p)]
        st = optimizer.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st["step"] += 1
        b1, b2 = grp["betas"]
        ops.adam_step(p.data, g, st["exp_avg"], st["exp_avg_sq"], int(st["step"]), grp["lr"], b1, b2, grp["eps"],
                      grp["weight_decay"])


def train(model, train_loader, optimizer, device, args, masks: Optional[Iterable[torch.Tensor]] = None,
          dp_microbatch: Optional[int] = None, group=None):
    """One epoch of main_moc.py:378-410: one Adam step per (virtual) slide on a random half of its patches.

    ``masks`` (an iterable of bool [N_i] tensors) is an addition for reproducible tests; without it the mask
    of every step is drawn as the reference does, ``torch.rand(N) > 0.5`` on the CPU default generator.

    ``dp_microbatch=G`` (also ``args.dp_microbatch``) switches to the data-parallel mode north_star names: the epoch's
    virtual slides are cut into micro-batches of G consecutive ones; slide j of a micro-batch is computed by rank
    ``j % world`` at the micro-batch's common parameters, each rank sums its gradients, ONE all-reduce (NCCL, 33 092
    floats + the G losses = 132 KB) sums them over the ranks and every rank applies the same single Adam step.  This
    changes the optimisation trajectory (G slides per step instead of one), so it is never the default and is checked
    against the oracle's ``train_epoch(dp_microbatch=G)``, not against the reference's loop.  Every rank draws every
    mask, so the CPU generators stay in step.  Returns the per-slide losses in slide order."""
    model.train()
    eng = engine_for(args)
    store = _store_of(train_loader, device)
    ds = train_loader.dataset
    masks = iter(masks) if masks is not None else None
    eng.ensure_domain(store, model.head_params())   # once per store: features outside the fast kernels' range?
    g = dp_microbatch if dp_microbatch is not None else getattr(args, "dp_microbatch", None)
    n_steps = len(ds)
    if not g or int(g) <= 1:
        flat = torch.empty(ops.NUM_PARAMS, dtype=torch.float32, device=store.device)
        # a plain Adam lets the whole step - forward, backward, update - replay from one CUDA graph per few-shot slide
        use_graph = bool(getattr(args, "cuda_graph", True)) and _is_plain_adam(optimizer) and \
            os.environ.get("MOC_TRAIN_GRAPH", "1") != "0"
        adam = _adam_dev(model, optimizer) if use_graph else None
        params = model.head_params()
        losses = []
        for k in range(n_steps):
            i = k % ds.real_len()
            n = store.n_rows(i)
            mask = next(masks) if masks is not None else (torch.rand(n) > 0.5)
            if use_graph:
                out = eng.train_step_graph(store, i, params, adam, mask)
                losses.append(out.loss.clone())
                continue
            out = eng.train_step(store, i, store.labels[i:i + 1], model.head_params(), mask.to(store.device), flat)
            _optimizer_step(model, optimizer, flat)
            losses.append(out.loss)
        return torch.cat(losses) if losses else torch.empty(0, device=store.device)

    import torch.distributed as dist
    g = int(g)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    one = torch.empty(ops.NUM_PARAMS, dtype=torch.float32, device=store.device)
    losses = []
    for k0 in range(0, n_steps, g):
        size = min(g, n_steps - k0)
        # [ summed gradient | loss of each slide of the micro-batch ]: one buffer, one all-reduce
        acc = torch.zeros(ops.NUM_PARAMS + size, dtype=torch.float32, device=store.device)
        params = model.head_params()
        for j in range(size):
            i = (k0 + j) % ds.real_len()
            n = store.n_rows(i)
            mask = next(masks) if masks is not None else (torch.rand(n) > 0.5)
            if j % world != rank:
                continue
            out = eng.train_step(store, i, store.labels[i:i + 1], params, mask.to(store.device), one)
            ops.accumulate_(acc[:ops.NUM_PARAMS], one)
            acc[ops.NUM_PARAMS + j] = out.loss[0]
        from .dist import allreduce_sum
        allreduce_sum(acc, group)
        _optimizer_step(model, optimizer, acc[:ops.NUM_PARAMS])
        losses.append(acc[ops.NUM_PARAMS:])
    return torch.cat(losses) if losses else torch.empty(0, device=store.device)


def _metrics(logits_all: torch.Tensor, labels: torch.Tensor, loss_vec: torch.Tensor, pred: torch.Tensor,
             loss_div: int, real_len: int, args) -> dict:
    """loss / acc / auc exactly as main_moc.py:436-460 and :499-520."""
    from sklearn.metrics import roc_auc_score
    pretrain = getattr(args, "pretrain", "conch")
    if pretrain not in TEMPERATURE:
        raise NotImplementedError
    test_loss = 0.0
    for v in loss_vec.cpu().tolist():  # the reference accumulates .item() values in a Python float
        test_loss += v
    correct = int((pred.long() == labels).sum().item())
    test_loss /= loss_div
    lg = logits_all.detach().cpu()
    y = labels.cpu().numpy()
    probs = torch.softmax(lg * TEMPERATURE[pretrain], dim=1)
    if probs.shape[1] == 2:
        auc = roc_auc_score(y, probs[:, 1].numpy())
    else:
        auc = roc_auc_score(y, probs.numpy(), multi_class="ovo", average="macro")
    return {"loss": test_loss, "acc": correct / real_len, "auc": auc}


def _gathered(ds, logits, labels):
    """Slide-sharded evaluation: every rank contributes its shard's rows, all ranks see the full split."""
    shard = getattr(ds, "shard", None)
    if shard is None:
        return logits, labels
    return shard.gather(logits, labels)


def zs_evaluation(loader, device, args, pooling_func=topj_pooling):
    """main_moc.py:412-460."""
    eng = engine_for(args)
    ds = loader.dataset
    store = _store_of(loader, device)
    real_len, set_len = ds.real_len(), ds.repeat_num
    ds.repeat_num = real_len
    name = _POOL_NAMES.get(pooling_func)
    if name is None:
        raise _lib.MocError(_lib.E_ARG, "pooling_func must be one of the four pooling functions of moc_b200.pooling")
    logits = eng.zero_shot_logits(store, name, check_domain=True)
    ds.repeat_num = set_len
    logits, labels = _gathered(ds, logits, store.labels)
    loss_vec, _, pred = ops.cross_entropy(logits, labels, want_pred=True)
    return _metrics(logits, labels, loss_vec, pred, _global_len(ds), _global_real_len(ds), args)


def evaluation(model, loader, device, args):
    """main_moc.py:462-520."""
    model.eval()
    eng = engine_for(args)
    ds = loader.dataset
    store = _store_of(loader, device)
    real_len, set_len = ds.real_len(), len(ds)
    ds.repeat_num = real_len
    logits = eng.eval_logits(store, model.head_params(), "eval", check_domain=True)
    ds.repeat_num = set_len
    logits, labels = _gathered(ds, logits, store.labels)
    loss_vec, _, pred = ops.cross_entropy(logits, labels, want_pred=True)
    return _metrics(logits, labels, loss_vec, pred, _global_len(ds), _global_real_len(ds), args)


def ablation_evaluation(loader, device, args):
    """main_moc.py:523-582 (``--ablation_study avg|sum|max``)."""
    eng = engine_for(args)
    ds = loader.dataset
    store = _store_of(loader, device)
    real_len, set_len = ds.real_len(), len(ds)
    ds.repeat_num = real_len
    logits = eng.ablation_logits(store, args.ablation_study, check_domain=True)
    ds.repeat_num = set_len
    logits, labels = _gathered(ds, logits, store.labels)
    loss_vec, _, pred = ops.cross_entropy(logits, labels, want_pred=True)
    return _metrics(logits, labels, loss_vec, pred, _global_len(ds), _global_real_len(ds), args)


def _global_len(ds) -> int:
    shard = getattr(ds, "shard", None)
    return len(ds) if shard is None else shard.global_len(ds)


def _global_real_len(ds) -> int:
    shard = getattr(ds, "shard", None)
    return ds.real_len() if shard is None else shard.n_global


def main(args, model, optimizer, train_loader, val_loader, test_loader, device, num_epoch: int = 25,
         is_main: bool = True):
    """main_moc.py:586-644.  The loaders / model / optimizer are arguments instead of script globals; outputs
    (prints, zs_results_*.json, best_results_*.json, best_model_*.pt) keep the reference's names and schema."""
    def dump(name, obj):
        if is_main:
            with open(os.path.join(args.result_dir, name), "w") as f:
                json.dump(obj, f, indent=4)

    if is_main:
        os.makedirs(args.result_dir, exist_ok=True)
    if args.ablation_study != "none":
        d = ablation_evaluation(test_loader, device, args)
        print(f"Ablation Study: {args.ablation_study}, Test: {d}")
        dump(f"ablation_results_{args.ablation_study}_shot_{args.shot}_fold_{args.fold}.json", d)
        return d
    zs_train, zs_val, zs_test = -1, -1, -1
    if args.check_zeroshot:
        zs_train = zs_evaluation(train_loader, device, args)
        zs_val = zs_evaluation(val_loader, device, args)
        zs_test = zs_evaluation(test_loader, device, args)
        print(f"Zero-shot Train: {zs_train}, Val: {zs_val}, Test: {zs_test}")
        dump(f"zs_results_shot_{args.shot}_fold_{args.fold}.json",
             {"zs_train": zs_train, "zs_val": zs_val, "zs_test": zs_test})
    best_val = 0
    test_at_best_val = 0
    test_acc_at_best_val = 0
    best_epoch = 0
    model_path = os.path.join(args.result_dir, f"best_model_shot_{args.shot}_fold_{args.fold}.pt")
    for epoch in range(num_epoch):
        print("Epoch: ", epoch)
        train(model, train_loader, optimizer, device, args)
        train_eval = evaluation(model, train_loader, device, args)
        val_eval = evaluation(model, val_loader, device, args)
        if val_eval["auc"] > best_val:
            test_eval = evaluation(model, test_loader, device, args)
            print(f"Epoch: {epoch}, Train: {train_eval}, Val: {val_eval}, Test: {test_eval}")
            best_val = val_eval["auc"]
            test_at_best_val = test_eval["auc"]
            test_acc_at_best_val = test_eval["acc"]
            best_epoch = epoch
            if is_main:
                torch.save(model.state_dict(), model_path)
        else:
            print(f"Epoch: {epoch}, Train: {train_eval}, Val: {val_eval}")
    print(f"Zero-shot Train: {zs_train}, Val: {zs_val}, Test: {zs_test}")
    print(f"Best Val: {best_val}, Test at Best Val: {test_at_best_val}, Test acc: {test_acc_at_best_val}, "
          f"Best Epoch: {best_epoch}")
    results = {
        "zero_shot_train": zs_train, "zero_shot_val": zs_val, "zero_shot_test": zs_test,
        "best_val": best_val, "test_at_best_val": test_at_best_val, "test_acc_at_best_val": test_acc_at_best_val,
        "best_epoch": best_epoch, "best_model_path": model_path,
    }
    dump(f"best_results_shot_{args.shot}_fold_{args.fold}.json", results)
    print("\nEnd training.")
    return results
