"""Batched driver of the hot path over a :class:`RaggedBagStore`.

One ``MocEngine`` holds the packed prompt matrices and the run's J / K / discard settings and exposes the
three passes the MOC loops are made of, each as a handful of kernel launches over *all* slides at once:

* ``zero_shot_logits``  - score + pooled top-K                      (zs_evaluation, main_moc.py:412-460)
* ``eval_logits``       - score + select + gate/combine + pooling    (evaluation,    main_moc.py:462-520)
* ``train_step``        - the same on one half-masked slide, then CE, backward, Adam (train, main_moc.py:378-410)

Scores depend only on the bag and the frozen prompts, never on the trained gate, so with
``cache_scores=True`` the key planes of a store are computed once and reused by every later pass; results are
bit-identical either way.  The default is off: every pass streams the bags
again, like the reference.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops
from .bag_store import RaggedBagStore

def _pooling_planes(pooling: str, c: int):
    """(select plane0, select step, value plane0, value step, smallest, needs the full layout) of a zero-shot pooling
    mode in the key layout of C classes.  The softmax selection indexes the C softmax planes, which the compact layout
    of wide class sets does not store: that mode pools over ``ops.expand_keys``."""
    from ._lib import PLANE_BG_SUM, PLANE_DIFF
    if pooling == "topj":
        return 0, 1, 0, 1, False, False
    if pooling == "delta_softmax":
        return c, 1, 0, 1, False, True
    if pooling == "delta_diff":
        return ops.key_plane(c, PLANE_DIFF), 0, 0, 1, False, False
    if pooling == "bottomk_irrel":
        return ops.key_plane(c, PLANE_BG_SUM), 0, 0, 1, True, False
    raise KeyError(pooling)


POOLINGS = ("topj", "delta_softmax", "delta_diff", "bottomk_irrel")


@dataclass
class StepOut:
    loss: torch.Tensor        # [1] device
    bag_logits: torch.Tensor  # [1,C] device
    n_selected: torch.Tensor  # int32 [1] device


class _TrainGraph:
    """A captured training step of one slide: static input (mask), static outputs, and what keeps them alive."""
    graph = mask_pinned = mask_dev = grads = head_ws = out = copied = None


class AdamDev:
    """torch.optim.Adam's state for the graph-captured training step: the optimizer's own exp_avg / exp_avg_sq tensors,
    its hyper-parameters, and the step count mirrored in device memory (ops.adam_prepare_dev advances it inside the
    graph).  ``count_host_step`` keeps the optimizer's host-side ``state['step']`` in line, so ``state_dict()`` and a
    later eager ``optimizer.step()`` see the right count."""

    def __init__(self, optimizer, params: List[torch.Tensor]):
        self.optimizer, self.params = optimizer, params
        group_of = {id(p): g for g in optimizer.param_groups for p in g["params"]}
        self.groups = [group_of[id(p)] for p in params]
        steps = set()
        for p in params:
            st = optimizer.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps.add(int(st["step"]))
        if len(steps) != 1:
            raise _lib.MocError(_lib.E_ARG, "the gate parameters have been stepped a different number of times")
        g0 = self.groups[0]
        if any((g["lr"], g["betas"]) != (g0["lr"], g0["betas"]) for g in self.groups):
            raise _lib.MocError(_lib.E_ARG, "graph-captured Adam needs one lr / betas for all gate parameters")
        dev = params[0].device
        self.step_dev = torch.tensor([steps.pop()], dtype=torch.int64, device=dev)
        self.scalars = torch.zeros(2, dtype=torch.float32, device=dev)

    def key(self):
        return (id(self.optimizer), self.step_dev.data_ptr()) + tuple(
            (g["lr"], g["betas"], g["eps"], g["weight_decay"]) for g in self.groups)

    def host_step(self) -> int:
        return int(self.optimizer.state[self.params[0]]["step"])

    def in_sync(self) -> bool:
        """False once someone else stepped the optimizer (the device counter no longer matches)."""
        return all(len(self.optimizer.state[p]) and int(self.optimizer.state[p]["step"]) == self._expected
                   for p in self.params) if hasattr(self, "_expected") else True

    def count_host_step(self) -> None:
        for p in self.params:
            self.optimizer.state[p]["step"] += 1
        self._expected = self.host_step()

    def apply(self, grad_views) -> None:
        """prepare + one apply per tensor, on the current stream (captured into the step's graph)."""
        g0 = self.groups[0]
        ops.adam_prepare_dev(self.step_dev, self.scalars, g0["lr"], g0["betas"][0], g0["betas"][1])
        for p, gv, grp in zip(self.params, grad_views, self.groups):
            st = self.optimizer.state[p]
            ops.adam_apply_dev(p.data, gv, st["exp_avg"], st["exp_avg_sq"], self.scalars, grp["betas"][0],
                               grp["betas"][1], grp["eps"], grp["weight_decay"])


class MocEngine:
    def __init__(self, zeroshot_weights: torch.Tensor, zeroshot_weights_ext: torch.Tensor, topj: int = 10,
                 topk: int = 10, discard_classifiers: Sequence[str] = (), normalize: bool = False,
                 cache_scores: bool = False, max_wave_rows: int = 48 * 1024 * 1024,
                 prompt_bank: Optional[tuple] = None):
        """``prompt_bank=(bank [n_prompts,512], prompts_per_class)`` keeps the class prompts UN-COLLAPSED on the scoring
        path (ops.BankPrompts): ``zeroshot_weights`` must then be the matrix the bank collapses to (it is checked) and
        every pass gives what it gives without the bank, within the scoring tolerance - the dense stress configuration
        of BASELINE.json configs[2], not something the reference does at run time (it collapses offline)."""
        if prompt_bank is None:
            self.prompts = ops.Prompts.pack(zeroshot_weights, zeroshot_weights_ext)
        else:
            bank, counts = prompt_bank
            self.prompts = ops.BankPrompts.pack(bank.to(zeroshot_weights.device), counts, zeroshot_weights_ext)
            w_from_bank = self.prompts.collapsed.packed[:len(counts)].t()
            if (w_from_bank - zeroshot_weights.float()).abs().max().item() > 1e-5:
                raise _lib.MocError(_lib.E_ARG, "zeroshot_weights is not what the prompt bank collapses to")
        self.n_classes = self.prompts.n_classes
        # slide_process never reads W_ext[:, :C] (only the background columns matter to it), but zs_evaluation with
        # bottomk_irrel_classifier_pooling pools (feats @ W_ext)[:, :C] (main_moc.py:428-432,
        # patch_selection_classifier.py:152-160).  In every shipped prompt file those columns equal W; when they do
        # not, that pooling mode scores against a second packing whose class columns are W_ext[:, :C].
        c = self.n_classes
        self._prompts_ext_fg = None
        if not torch.equal(zeroshot_weights_ext[:, :c].float(), zeroshot_weights.float()):
            self._prompts_ext_fg = ops.Prompts.pack(zeroshot_weights_ext[:, :c].contiguous(), zeroshot_weights_ext)
        self.topj, self.topk = int(topj), int(topk)
        self.discard = tuple(discard_classifiers or ())
        self.normalize = bool(normalize)
        self.cache_scores = bool(cache_scores)
        self.max_wave_rows = int(max_wave_rows)
        # per-store state lives exactly as long as the store (or HostBags) object: weak keys, so a new store that
        # happens to reuse a dead one's address can never see its key planes or selection layout
        self._key_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()
        self._layout_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()
        self.score_events = None  # when a list: (start, stop) CUDA events around every scoring launch
        # stores whose features left the fast kernels' range (|x| >= 4094 for the FP16x3 gate MLP, >= 65504 for the
        # tensor-core scoring of wide prompt sets): found by a flag-checked pass, they are served by the range-free
        # kernels from then on (3xTF32 gate, fp32 CUDA-core scoring) - the reference is finite for any finite feature
        self._wide_stores: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()
        self._graphs: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()   # store -> captured training steps

    # ---- scoring -------------------------------------------------------------------------------
    def keys_for(self, store: RaggedBagStore, lo: int = 0, hi: Optional[int] = None, wide: bool = False) -> torch.Tensor:
        """Key planes [2C+3, rows] for slides [lo, hi) of the store (whole store when cached).  ``wide``: score on the
        fp32 CUDA-core kernel (only differs for wide prompt sets, whose tensor-core kernel has an |x| limit)."""
        hi = len(store) if hi is None else hi
        wide = bool(wide and self.prompts.tc is not None)
        if self.cache_scores:
            hit = self._key_cache.get(store)
            if hit is None or hit[0].size(1) != store.total_rows or (wide and not hit[1]):
                k = ops.score_keys(store.feat, self.prompts, self.normalize, wide=True) if wide else self._score(store.feat)
                hit = self._key_cache[store] = (k, wide)
            return hit[0][:, store.offsets_h[lo]:store.offsets_h[hi]]
        r0, r1 = store.offsets_h[lo], store.offsets_h[hi]
        if wide:
            return ops.score_keys(store.feat[r0:r1], self.prompts, self.normalize, wide=True)
        return self._score(store.feat[r0:r1])

    # ---- out-of-range features ---------------------------------------------------------------------
    def _arm_flags(self, device):
        ws = ops.head_workspace(device)
        ws.clear_flag()
        if self.prompts.tc_flag is not None:
            self.prompts.tc_flag.zero_()
        return ws

    def _flags_raised(self, ws) -> bool:
        """One or two 4-byte reads (synchronises): did a fast kernel meet a feature outside its range?"""
        bad = ws.overflowed()
        if self.prompts.tc_flag is not None:
            bad = bool(int(self.prompts.tc_flag.item()) != 0) or bad
        return bad

    def is_wide(self, store) -> bool:
        return bool(self._wide_stores.get(store, False))

    # The streaming kernel's best CTA count is a property of the individual GPU (132 on most B200s, 136 on some, DESIGN
    # section 4).  The first large scoring call of a process times a few candidates on the caller's own data - the
    # results are identical whatever the count - and every later call on that device uses the winner.
    _TUNED_CTAS: Dict[int, int] = {}
    _TUNE_CANDIDATES = (132, 128, 136, 124, 140)
    _TUNE_MIN_ROWS = 4_000_000

    def _tuned_ctas(self, feat: torch.Tensor, scratch: torch.Tensor) -> int:
        import os
        dev = feat.device.index if feat.device.index is not None else torch.cuda.current_device()
        got = MocEngine._TUNED_CTAS.get(dev)
        if got is not None:
            return got
        if (self.prompts.tc is not None or feat.size(0) < self._TUNE_MIN_ROWS
                or os.environ.get("MOC_SCORE_AUTOTUNE", "1") == "0"
                or torch.cuda.get_device_properties(dev).multi_processor_count != 148):
            return 0                                   # the kernel's built-in default; try again on a larger call
        best = {c: float("inf") for c in self._TUNE_CANDIDATES}
        for _ in range(3):                             # interleaved rounds, minimum per candidate
            for c in self._TUNE_CANDIDATES:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.score_keys(feat, self.prompts, self.normalize, out=scratch, max_ctas=c)
                b.record()
                b.synchronize()
                best[c] = min(best[c], a.elapsed_time(b))
        winner = min(best, key=best.get)
        if best[winner] > 0.995 * best[132]:           # not clearly better than the default: keep the default
            winner = 132
        MocEngine._TUNED_CTAS[dev] = winner
        return winner

    def _score(self, feat: torch.Tensor, out: Optional[torch.Tensor] = None, max_ctas: int = 0) -> torch.Tensor:
        if max_ctas == 0:
            if out is None:     # the calibration launches write the same keys into the buffer the real launch fills
                out = ops.alloc_keys(self.n_classes, feat.size(0), feat.device)
            max_ctas = self._tuned_ctas(feat, out)
        if self.score_events is None:
            return ops.score_keys(feat, self.prompts, self.normalize, out=out, max_ctas=max_ctas)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        k = ops.score_keys(feat, self.prompts, self.normalize, out=out, max_ctas=max_ctas)
        b.record()
        self.score_events.append((a, b, feat.size(0)))
        return k

    def drop_cache(self, store: Optional[RaggedBagStore] = None) -> None:
        if store is None:
            self._key_cache.clear()
        else:
            self._key_cache.pop(store, None)

    def _waves(self, store: RaggedBagStore):
        """Split the store into runs of whole slides of at most max_wave_rows rows (key-plane memory bound)."""
        lo, n = 0, len(store)
        while lo < n:
            hi = lo + 1
            while hi < n and store.offsets_h[hi + 1] - store.offsets_h[lo] <= self.max_wave_rows:
                hi += 1
            yield lo, hi
            lo = hi

    def _layout(self, store: RaggedBagStore, lo: int, hi: int):
        per_store = self._layout_cache.get(store)
        if per_store is None:
            per_store = self._layout_cache[store] = {}
        key = (lo, hi, self.topj)
        hit = per_store.get(key)
        r0 = store.offsets_h[lo]
        if hit is None or hit[1][-1] != store.offsets_h[hi] - r0:
            offs_h = [v - r0 for v in store.offsets_h[lo:hi + 1]]
            offs = (store.offsets[lo:hi + 1] - r0).contiguous()
            base_h = ops.selection_layout(offs_h, self.n_classes, self.topj)
            base = torch.tensor(base_h, dtype=torch.int64, device=store.device)
            hit = (offs, offs_h, base, base_h)
            per_store[key] = hit
        return hit

    # ---- zero-shot -----------------------------------------------------------------------------
    def zero_shot_logits(self, store: RaggedBagStore, pooling: str = "topj", check_domain: bool = False) -> torch.Tensor:
        """``check_domain``: see :meth:`eval_logits` (here only the tensor-core scoring of wide prompt sets can overflow)."""
        c = self.n_classes
        sp0, ss, vp0, vs, small, want_full = _pooling_planes(pooling, c)
        want_full = want_full and ops.num_key_planes(c) != 2 * c + 3
        out = torch.empty(len(store), c, dtype=torch.float32, device=store.device)
        ext_fg = self._prompts_ext_fg if pooling == "bottomk_irrel" else None
        wide = self.is_wide(store)
        armed = check_domain and not wide and self.prompts.tc is not None
        if armed:
            ws = self._arm_flags(store.device)
            if ext_fg is not None and ext_fg.tc_flag is not None:
                ext_fg.tc_flag.zero_()
        for lo, hi in self._waves(store):
            if ext_fg is None:
                keys = self.keys_for(store, lo, hi, wide)
            else:   # class planes of this pass = feats @ W_ext[:, :C]; the background sum is the same either way
                keys = ops.score_keys(store.feat[store.offsets_h[lo]:store.offsets_h[hi]], ext_fg, self.normalize, wide=wide)
            offs, _, _, _ = self._layout(store, lo, hi)
            if want_full:
                keys = ops.expand_keys(keys, c)
            out[lo:hi] = ops.pool_topk(keys, offs, hi - lo, c, self.topk, sp0, ss, vp0, vs, small)
        if armed:
            bad = self._flags_raised(ws) or (ext_fg is not None and ext_fg.tc_flag is not None and int(ext_fg.tc_flag.item()) != 0)
            if bad:
                self._wide_stores[store] = True
                return self.zero_shot_logits(store, pooling)
        return out

    # ---- evaluation ----------------------------------------------------------------------------
    def eval_logits(self, store: RaggedBagStore, params: ops.HeadParams, mode: str = "eval",
                    out: Optional[torch.Tensor] = None, check_domain: bool = False) -> torch.Tensor:
        """Bag logits [n_slides, C] of every slide in the store under the current gate parameters.

        ``check_domain=True`` (what the loops pass; they read the logits back right afterwards): the fast kernels'
        overflow flags are cleared before the pass and read after it (a host sync); if one is raised the store is
        marked and the pass repeated on the range-free kernels, which then serve every later pass over that store."""
        c = self.n_classes
        if out is None:
            out = torch.empty(len(store), c, dtype=torch.float32, device=store.device)
        disc = _lib.discard_bits(self.discard)
        act = _lib.active_bits(self.discard, mode)
        wide = self.is_wide(store)
        armed = check_domain and not wide
        if armed:
            ws = self._arm_flags(store.device)
        for lo, hi in self._waves(store):
            keys = self.keys_for(store, lo, hi, wide)
            offs, offs_h, base, base_h = self._layout(store, lo, hi)
            feat = store.feat[store.offsets_h[lo]:store.offsets_h[hi]]
            sel = ops.select_union(keys, offs, offs_h, c, self.topj, disc, None, base, base_h)
            ho = ops.head_forward(feat, keys, c, sel, params, act, self.topk, wide=wide)
            out[lo:hi] = ho.bag_logits
        if armed:
            if self._flags_raised(ws):
                self._wide_stores[store] = True
                return self.eval_logits(store, params, mode, out)
            self._wide_stores.setdefault(store, False)
        return out

    def ablation_logits(self, store: RaggedBagStore, how: str, check_domain: bool = False) -> torch.Tensor:
        """Un-gated avg / sum / max of the four planes (ablation_evaluation, main_moc.py:523-582)."""
        c = self.n_classes
        out = torch.empty(len(store), c, dtype=torch.float32, device=store.device)
        wide = self.is_wide(store)
        armed = check_domain and not wide and self.prompts.tc is not None
        if armed:
            ws = self._arm_flags(store.device)
        for lo, hi in self._waves(store):
            keys = self.keys_for(store, lo, hi, wide)
            offs, offs_h, base, base_h = self._layout(store, lo, hi)
            feat = store.feat[store.offsets_h[lo]:store.offsets_h[hi]]
            sel = ops.select_union(keys, offs, offs_h, c, self.topj, 0, None, base, base_h)
            out[lo:hi] = ops.ablation_pool(keys, c, sel, how, self.topk)
            del feat
        if armed and self._flags_raised(ws):
            self._wide_stores[store] = True
            return self.ablation_logits(store, how)
        return out

    def ensure_domain(self, store: RaggedBagStore, params: ops.HeadParams) -> bool:
        """Decide, once per store, whether its features fit the fast kernels' range: one flag-checked evaluation pass
        over it (the few-shot training bags are a few dozen slides).  Every unmasked row is gated by that pass, and a
        masked training step only ever gates a subset of them.  Returns True when the range-free kernels are needed."""
        if store not in self._wide_stores:
            self.eval_logits(store, params, "train", check_domain=True)
        return self.is_wide(store)

    # ---- training ------------------------------------------------------------------------------
    def train_step(self, store: RaggedBagStore, slide: int, label_dev: torch.Tensor, params: ops.HeadParams,
                   row_mask: Optional[torch.Tensor], grads_out: torch.Tensor, head_ws=None) -> StepOut:
        """Forward + CE + backward of one (masked) slide; fills ``grads_out`` (flat, 33 092 floats)."""
        c = self.n_classes
        wide = self.is_wide(store)      # decided once per store by ensure_domain(): a step has no sync to poll a flag at
        keys = self.keys_for(store, slide, slide + 1, wide)
        offs, offs_h, base, base_h = self._layout(store, slide, slide + 1)
        feat = store.bag(slide)
        sel = ops.select_union(keys, offs, offs_h, c, self.topj, _lib.discard_bits(self.discard), row_mask, base, base_h)
        act = _lib.active_bits(self.discard, "train")
        ho = ops.head_forward(feat, keys, c, sel, params, act, self.topk, wide=wide, ws=head_ws)
        loss, dl, _ = ops.cross_entropy(ho.bag_logits, label_dev, want_grad=True)
        ops.head_backward(feat, keys, c, sel, params, act, self.topk, ho.pool_pos, dl, out=grads_out)
        return StepOut(loss, ho.bag_logits, sel.sel_count)

    # ---- training step as a CUDA graph -------------------------------------------------------------
    def train_step_graph(self, store: RaggedBagStore, slide: int, params: ops.HeadParams, adam: "AdamDev",
                         row_mask_host: torch.Tensor) -> StepOut:
        """One masked training step INCLUDING the Adam update, replayed from a CUDA graph.

        The eager step is ~16 kernel launches, a dozen small allocations and a pageable mask copy: ~0.3 ms of host
        time around ~0.1 ms of GPU work (main_moc.py:380-410 is one such step per few-shot slide, 25 epochs).  The
        launch sequence of a slide depends only on its size, so it is captured once per (store, slide) - score,
        select with the mask read from a device buffer, gate / combine / pool, CE, backward, Adam with its step counter
        in device memory - and every later visit is one pinned H2D copy of the fresh half mask plus one graph launch.
        Same kernels, same order, same arguments as :meth:`train_step` + ``ops.adam_step``: bit-identical results.
        The returned tensors are the graph's static outputs: copy what must outlive the next replay."""
        per_store = self._graphs.get(store)
        if per_store is None:
            per_store = self._graphs[store] = {}
        key = (slide, params.w1.data_ptr(), params.b1.data_ptr(), params.w2.data_ptr(), params.b2.data_ptr(),
               adam.key(), self.is_wide(store))
        g = per_store.get(key)
        if g is None:
            g = per_store[key] = self._capture_train_step(store, slide, params, adam)
        if g.copied is not None:
            g.copied.synchronize()      # the host runs ahead of the GPU: the previous visit's mask must have left the pinned buffer
        g.mask_pinned.copy_(row_mask_host.view(torch.uint8) if row_mask_host.dtype == torch.bool else row_mask_host)
        g.mask_dev.copy_(g.mask_pinned, non_blocking=True)
        if g.copied is None:
            g.copied = torch.cuda.Event()
        g.copied.record()
        g.graph.replay()
        adam.count_host_step()
        return g.out

    def _capture_train_step(self, store, slide, params, adam):
        dev = store.device
        n = store.n_rows(slide)
        self._layout(store, slide, slide + 1)           # host-built layout tensors: not inside the capture
        if self.cache_scores:
            self.keys_for(store, slide, slide + 1, self.is_wide(store))
        g = _TrainGraph()
        g.mask_pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        g.mask_dev = torch.ones(n, dtype=torch.uint8, device=dev)
        g.grads = torch.empty(ops.NUM_PARAMS, dtype=torch.float32, device=dev)
        g.head_ws = ops.HeadWorkspace(dev)
        label = store.labels[slide:slide + 1]
        views = ops.split_grads(g.grads)
        torch.cuda.current_stream().synchronize()
        g.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g.graph):
            out = self.train_step(store, slide, label, params, g.mask_dev, g.grads, head_ws=g.head_ws)
            adam.apply(views)
        g.out = out
        return g

    # ---- host-resident bags (end-to-end path) ---------------------------------------------------
    def eval_logits_host(self, host: "HostBags", params: ops.HeadParams, out: Optional[torch.Tensor] = None
                         ) -> torch.Tensor:
        """evaluation over bags that live in pinned HOST memory: chunks are copied to the GPU on a copy stream
        into two staging buffers while the previous chunk is scored (what a caller holding numpy / h5 data
        pays end to end).  Returns device logits [n_slides, C]."""
        c = self.n_classes
        dev = host.device
        if out is None:
            out = torch.empty(host.n_slides, c, dtype=torch.float32, device=dev)
        disc = _lib.discard_bits(self.discard)
        act = _lib.active_bits(self.discard, "eval")
        cur = torch.cuda.current_stream()
        copy = host.copy_stream
        copy.wait_stream(cur)
        lo = 0
        for ci, ch in enumerate(host.chunks):
            b = ci % 2
            buf = host.staging[b][:ch.rows]
            with torch.cuda.stream(copy):
                if host.free[b] is not None:
                    copy.wait_event(host.free[b])
                buf.copy_(ch.feat, non_blocking=True)
                host.ready[b].record(copy)
            cur.wait_event(host.ready[b])
            per_host = self._layout_cache.get(host)
            if per_host is None:
                per_host = self._layout_cache[host] = {}
            lay = per_host.get((ci, self.topj))
            if lay is None:
                base_h = ops.selection_layout(ch.offsets_h, c, self.topj)
                lay = (torch.tensor(base_h, dtype=torch.int64, device=dev), base_h)
                per_host[(ci, self.topj)] = lay
            keys = self._score(buf)
            sel = ops.select_union(keys, ch.offsets, ch.offsets_h, c, self.topj, disc, None, lay[0], lay[1])
            ho = ops.head_forward(buf, keys, c, sel, params, act, self.topk)
            n = len(ch.offsets_h) - 1
            out[lo:lo + n] = ho.bag_logits
            lo += n
            host.free[b] = torch.cuda.Event()
            host.free[b].record(cur)
        return out
