"""Tensor-level wrappers over the C ABI: torch tensors in, torch tensors out, current CUDA stream.

torch is used here only for device memory, streams and (elsewhere) torch.distributed; every kernel that
runs is one of ours from libmoc_b200.so.  Nothing in this module falls back to torch math.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import os

import torch

from . import _lib
from ._lib import MocError, check

LAUNCHES = 0  # kernels of ours launched through this module (bench.py reports the count of its timed region)


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


REGW_MAX_COLS = 8   # up to here moc_score_keys keeps the prompts in registers and is HBM-bound on CUDA cores
# developer switch: "simt" forces the CUDA-core kernel for wide prompt sets, "tc" forces tensor cores for all
SCORE_IMPL = os.environ.get("MOC_SCORE_IMPL", "auto")

D = 512
HIDDEN = 64
GATES = 4
NUM_PARAMS = _lib.NUM_PARAMS


def _stream() -> int:
    """Raw handle of torch's current CUDA stream (every kernel is enqueued there).  The two C accessors are ~20x
    cheaper than building a torch.cuda.Stream object, which showed up as a quarter of a training step's host time."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:  # private accessors moved: fall back to the public object
        return torch.cuda.current_stream().cuda_stream


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise MocError(_lib.E_ARG, "%s must be a CUDA tensor (moc_b200 has no CPU path)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ---------------------------------------------------------------------------------------------
@dataclass
class Prompts:
    """The two zero-shot classifier matrices in the packed layout the scoring kernel reads."""
    packed: torch.Tensor      # [cols_pad, 512]
    n_classes: int
    n_ext: int
    tc: Optional[torch.Tensor] = None   # uint8 image for the tensor-core scoring kernel (wide prompt sets)
    tc_flag: Optional[torch.Tensor] = None  # int32 [1] view into ``tc``: set when a score came out non-finite

    @staticmethod
    def pack(w: torch.Tensor, w_ext: torch.Tensor) -> "Prompts":
        w, w_ext = _dev_f32(w, "zeroshot_weights"), _dev_f32(w_ext, "zeroshot_weights_ext")
        if w.dim() != 2 or w_ext.dim() != 2 or w.size(0) != D or w_ext.size(0) != D:
            raise MocError(_lib.E_SHAPE, "prompt matrices must be [512,C] and [512,C_ext]")
        c, ce = w.size(1), w_ext.size(1)
        lib = _lib.load()
        packed = torch.empty(lib.moc_packed_cols(c, ce), D, device=w.device, dtype=torch.float32)
        check(lib.moc_pack_prompts(w.data_ptr(), c, w_ext.data_ptr(), ce, packed.data_ptr(), _stream()))
        pr = Prompts(packed, c, ce)
        if ce > REGW_MAX_COLS or SCORE_IMPL == "tc":
            nb = lib.moc_prompts_tc_bytes(c, ce)
            pr.tc = torch.empty(nb, dtype=torch.uint8, device=w.device)
            check(lib.moc_prepare_prompts_tc(packed.data_ptr(), c, ce, pr.tc.data_ptr(), nb, _stream()))
            off = lib.moc_prompts_tc_flag_offset(c, ce)
            pr.tc_flag = pr.tc[off:off + 4].view(torch.int32)
        return pr

    def check_finite(self) -> None:
        """Raise if the tensor-core scoring kernel produced a non-finite score since the image was built
        (non-finite features, or |x| >= 65504).  Synchronises; call it where the results are read anyway."""
        if self.tc_flag is not None and int(self.tc_flag.item()) != 0:
            raise MocError(_lib.E_ARG, "scoring produced non-finite values (non-finite features or |x| >= 65504)")


def collapse_prompt_bank(bank: torch.Tensor, prompts_per_class: Sequence[int]) -> torch.Tensor:
    """[n_prompts,512] prompt embeddings (class after class) -> zero-shot classifier matrix [512,C] with unit
    columns, as utils/zeroshot_utils.py:29-50 builds it (normalise rows, mean per class, normalise)."""
    bank = _dev_f32(bank, "bank")
    counts = [int(c) for c in prompts_per_class]
    if bank.dim() != 2 or bank.size(1) != D or sum(counts) != bank.size(0) or min(counts, default=0) < 1:
        raise MocError(_lib.E_SHAPE, "bank must be [sum(prompts_per_class),512] with at least one prompt per class")
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    offs_d = torch.tensor(offs, dtype=torch.int32, device=bank.device)
    out = torch.empty(D, len(counts), dtype=torch.float32, device=bank.device)
    _count(1)
    check(_lib.load().moc_collapse_prompt_bank(bank.data_ptr(), offs_d.data_ptr(), len(counts), out.data_ptr(), _stream()))
    return out


@dataclass
class BankPrompts:
    """An UN-COLLAPSED prompt bank on the scoring path (BASELINE.json configs[2]): every prompt
    of every class stays a column of the tensor-core contraction, the per-class mean and the 1/||mean|| rescale of
    utils/zeroshot_utils.py:38-44 happen in the kernel's epilogue.  ``collapsed`` is the ordinary packing of the matrix
    the reference would score against (``collapse_prompt_bank`` + the background columns): the range-free fallback
    and the cross-check."""
    image: torch.Tensor          # uint8: split / swizzled K-block tiles, tail, packed fp32 rows
    n_classes: int
    n_prompts: int
    n_bg: int
    collapsed: "Prompts"
    tc_flag: Optional[torch.Tensor] = None

    @property
    def n_ext(self) -> int:
        return self.n_classes + self.n_bg

    @property
    def n_cols(self) -> int:
        return self.n_prompts + self.n_bg

    @property
    def tc(self):                # "has a tensor-core scoring path with an |x| limit" for the engine's domain logic
        return self.image

    @staticmethod
    def pack(bank: torch.Tensor, prompts_per_class: Sequence[int], w_ext: torch.Tensor) -> "BankPrompts":
        """bank [n_prompts,512] (class after class), prompts per class, and zeroshot_weights_ext [512, C+n_bg] whose
        last n_bg columns are the background prompts (its first C columns are not scored, as in slide_process)."""
        bank = _dev_f32(bank, "bank")
        w_ext = _dev_f32(w_ext, "zeroshot_weights_ext")
        counts = [int(c) for c in prompts_per_class]
        c = len(counts)
        if bank.dim() != 2 or bank.size(1) != D or sum(counts) != bank.size(0) or min(counts, default=0) < 1:
            raise MocError(_lib.E_SHAPE, "bank must be [sum(prompts_per_class),512] with at least one prompt per class")
        if w_ext.dim() != 2 or w_ext.size(0) != D or w_ext.size(1) <= c:
            raise MocError(_lib.E_SHAPE, "zeroshot_weights_ext must be [512, C + n_bg] with n_bg >= 1")
        n_bg = w_ext.size(1) - c
        lib = _lib.load()
        nb = lib.moc_prompt_bank_tc_bytes(bank.size(0), n_bg)
        if nb == 0:
            raise MocError(_lib.E_SHAPE, "prompt bank of %d + %d columns exceeds this build's limit" % (bank.size(0), n_bg))
        offs = [0]
        for n in counts:
            offs.append(offs[-1] + n)
        offs_d = torch.tensor(offs, dtype=torch.int32, device=bank.device)
        bg = w_ext[:, c:].t().contiguous()
        image = torch.empty(nb, dtype=torch.uint8, device=bank.device)
        _count(4)
        check(lib.moc_prepare_prompt_bank_tc(bank.data_ptr(), offs_d.data_ptr(), c, bank.size(0), bg.data_ptr(), n_bg,
                                             image.data_ptr(), nb, _stream()))
        off = lib.moc_prompt_bank_tc_flag_offset(bank.size(0), n_bg)
        w = collapse_prompt_bank(bank, counts)
        return BankPrompts(image, c, bank.size(0), n_bg, Prompts.pack(w, torch.cat([w, w_ext[:, c:]], dim=1).contiguous()),
                           image[off:off + 4].view(torch.int32))


def num_key_planes(n_classes: int) -> int:
    """Planes the scoring kernels write for C classes: 2C+3, or C+4 in the compact layout of wide class sets
    (include/moc_b200.h: a log-sum-exp plane instead of the C softmax planes, which every consumer forms on the fly)."""
    return int(_lib.load().moc_num_key_planes(int(n_classes)))


def key_plane(n_classes: int, which: int) -> int:
    """Index of a named plane (``_lib.PLANE_*``) in the layout the scoring kernels use for C classes; -1 if not stored."""
    return int(_lib.load().moc_key_plane(int(n_classes), int(which)))


def alloc_keys(n_classes: int, rows: int, device) -> torch.Tensor:
    """Key planes [planes, rows] whose plane stride is a multiple of 4 floats: every plane of a slide then sits in the
    same 16-byte phase, which the selection kernel's two-plane (compact softmax) columns need for vector loads."""
    pad = (int(rows) + 3) & ~3
    return torch.empty(num_key_planes(n_classes), pad, dtype=torch.float32, device=device)[:, :rows]


def expand_keys(keys: torch.Tensor, n_classes: int) -> torch.Tensor:
    """The full 2C+3-plane layout [L | softmax | diff | bg sum | bg max] from whatever layout ``keys`` uses for C classes
    (a copy when it is the full one already): what the stand-alone helpers, the zero-shot softmax pooling of wide class
    sets and tests index directly."""
    rows = keys.size(1)
    full = torch.empty(2 * n_classes + 3, max(rows, 1), dtype=torch.float32, device=keys.device)[:, :rows]
    _count(1)
    check(_lib.load().moc_expand_keys(keys.data_ptr(), keys.stride(0), int(n_classes), rows, full.data_ptr(), full.stride(0),
                                      _stream()))
    return full


def score_keys(feat: torch.Tensor, prompts, normalize: bool = False,
               out: Optional[torch.Tensor] = None, max_ctas: int = 0, wide: bool = False,
               check_domain: bool = False) -> torch.Tensor:
    """keys [2C+3, R] (plane-major) for R rows of feat [R,512].  ``max_ctas`` (streaming kernel only) leaves SMs free
    for kernels running on another stream; 0 = the kernel's own best (132 of 148).

    Wide prompt sets are scored on the tensor cores with FP16x3 operands, which overflow for |x| >= 65504 (the fp32
    reference does not).  ``wide=True`` scores on the fp32 CUDA-core kernel instead (no range limit);
    ``check_domain=True`` reads the kernel's overflow flag after the launch (one host sync) and repeats the launch
    that way when it is set - for callers that synchronise right afterwards anyway (slide_process)."""
    feat = _dev_f32(feat, "feat")
    if feat.dim() != 2 or feat.size(1) != D:
        raise MocError(_lib.E_SHAPE, "feat must be [rows,512], got %s" % (tuple(feat.shape),))
    r = feat.size(0)
    if out is None:
        out = alloc_keys(prompts.n_classes, r, feat.device)
    _count(1)
    if isinstance(prompts, BankPrompts):
        if wide:      # range-free: the collapsed matrix on the fp32 kernels (scoring is linear in the prompts)
            return score_keys(feat, prompts.collapsed, normalize, out=out, max_ctas=max_ctas, wide=True)
        if check_domain:
            prompts.tc_flag.zero_()
        check(_lib.load().moc_score_keys_bank_tc(feat.data_ptr(), r, prompts.image.data_ptr(), prompts.n_classes,
                                                 prompts.n_prompts, prompts.n_bg, int(bool(normalize)), out.data_ptr(),
                                                 out.stride(0), _stream()))
        if check_domain and int(prompts.tc_flag.item()) != 0:
            return score_keys(feat, prompts, normalize, out=out, max_ctas=max_ctas, wide=True)
        return out
    if prompts.tc is not None and SCORE_IMPL != "simt" and not wide:
        if check_domain:
            prompts.tc_flag.zero_()
        check(_lib.load().moc_score_keys_tc(feat.data_ptr(), r, prompts.tc.data_ptr(), prompts.n_classes,
                                            prompts.n_ext, int(bool(normalize)), out.data_ptr(), out.stride(0),
                                            _stream()))
        if check_domain and int(prompts.tc_flag.item()) != 0:
            return score_keys(feat, prompts, normalize, out=out, max_ctas=max_ctas, wide=True)
    else:
        check(_lib.load().moc_score_keys_ex(feat.data_ptr(), r, prompts.packed.data_ptr(), prompts.n_classes,
                                            prompts.n_ext, int(bool(normalize)), out.data_ptr(), out.stride(0),
                                            int(max_ctas), _stream()))
    return out


# ---------------------------------------------------------------------------------------------
@dataclass
class Selection:
    """Union of the top-J selections of a batch of slides (device-resident, ragged)."""
    sel_base: torch.Tensor     # int64 [n_slides+1], start of every slide's region
    sel_base_h: List[int]
    sel_rows: torch.Tensor     # int32 [capacity], absolute feat rows, ascending per slide, -1 padded
    sel_local: torch.Tensor    # int32 [capacity], index inside the (masked) bag = the reference's selected_index
    sel_count: torch.Tensor    # int32 [n_slides]
    n_slides: int

    @property
    def capacity(self) -> int:
        return self.sel_base_h[-1]


def selection_layout(offsets_h: Sequence[int], n_classes: int, topj: int):
    bound = topj * (2 * n_classes + 2)
    base = [0]
    for i in range(len(offsets_h) - 1):
        base.append(base[-1] + min(offsets_h[i + 1] - offsets_h[i], bound))
    return base


def select_union(keys: torch.Tensor, offsets: torch.Tensor, offsets_h: Sequence[int], n_classes: int, topj: int,
                 discard_mask: int = 0, row_mask: Optional[torch.Tensor] = None,
                 sel_base: Optional[torch.Tensor] = None, sel_base_h: Optional[List[int]] = None) -> Selection:
    n_slides = len(offsets_h) - 1
    total_rows = int(offsets_h[-1])
    dev = keys.device
    if sel_base_h is None:
        sel_base_h = selection_layout(offsets_h, n_classes, topj)
        sel_base = torch.tensor(sel_base_h, dtype=torch.int64, device=dev)
    cap = max(sel_base_h[-1], 1)
    sel_rows = torch.empty(cap, dtype=torch.int32, device=dev)
    sel_local = torch.empty(cap, dtype=torch.int32, device=dev)
    sel_count = torch.empty(max(n_slides, 1), dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws_bytes = lib.moc_select_workspace_bytes(total_rows, n_slides)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if row_mask is not None:
        if row_mask.dtype == torch.bool:
            row_mask = row_mask.view(torch.uint8)
        row_mask = row_mask.contiguous()
        if not row_mask.is_cuda or row_mask.numel() != total_rows or row_mask.dtype != torch.uint8:
            raise MocError(_lib.E_ARG, "row_mask must be a CUDA bool/uint8 tensor with one entry per row")
    _count(2)
    check(lib.moc_select_union(keys.data_ptr(), keys.stride(0), offsets.data_ptr(), n_slides, total_rows, n_classes,
                               int(topj), int(discard_mask), _ptr(row_mask), sel_base.data_ptr(), sel_rows.data_ptr(),
                               sel_local.data_ptr(), sel_count.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    return Selection(sel_base, list(sel_base_h), sel_rows, sel_local, sel_count, n_slides)


def topj_sorted(values: torch.Tensor, j: int, largest: bool = True, want_values: bool = False):
    """Column-wise Tensor.topk(j, dim=0, largest, sorted=True) of a [N,C] (or [N]) CUDA tensor -> int64 [j,C]."""
    v = values
    if not v.is_cuda:
        raise MocError(_lib.E_ARG, "values must be a CUDA tensor (moc_b200 has no CPU path)")
    if v.dtype != torch.float32:
        v = v.float()
    squeeze = v.dim() == 1
    if squeeze:
        v = v.unsqueeze(1)
    n, c = v.shape
    j = min(int(j), n)
    idx = torch.empty(j, c, dtype=torch.int64, device=v.device)
    vals = torch.empty(j, c, dtype=torch.float32, device=v.device) if want_values else None
    _count(1)
    check(_lib.load().moc_topj_sorted(v.data_ptr(), n, v.stride(0), c, v.stride(1), j, int(bool(largest)),
                                      idx.data_ptr(), c, _ptr(vals), _stream()))
    if squeeze:
        idx = idx[:, 0]
        vals = vals[:, 0] if vals is not None else None
    return (idx, vals) if want_values else idx


def pool_topk(keys: torch.Tensor, offsets: torch.Tensor, n_slides: int, n_classes: int, topk: int,
              sel_plane0: int, sel_step: int, val_plane0: int, val_step: int, smallest: bool = False) -> torch.Tensor:
    out = torch.empty(n_slides, n_classes, dtype=torch.float32, device=keys.device)
    _count(1)
    check(_lib.load().moc_pool_topk(keys.data_ptr(), keys.stride(0), offsets.data_ptr(), n_slides, n_classes,
                                    int(topk), sel_plane0, sel_step, int(smallest), val_plane0, val_step,
                                    out.data_ptr(), _stream()))
    return out


# ---------------------------------------------------------------------------------------------
@dataclass
class HeadParams:
    """senet parameters as four contiguous fp32 CUDA tensors (views into one flat buffer when owned by
    :class:`moc_b200.model.senet`)."""
    w1: torch.Tensor
    b1: torch.Tensor
    w2: torch.Tensor
    b2: torch.Tensor


@dataclass
class HeadOut:
    final: torch.Tensor        # [capacity, C]  gated sum of the four planes per selected row
    bag_logits: torch.Tensor   # [n_slides, C]
    pool_pos: torch.Tensor     # int32 [n_slides, C, topk]
    gate: Optional[torch.Tensor]  # [capacity, 4]
    domain_flag: Optional[torch.Tensor] = None   # int32 [1]: non-zero once the FP16x3 gate kernel left its |x| < 4094 range


class HeadWorkspace:
    """The gate kernels' scratch (their split / swizzled W1 image) plus the int32 domain flag the FP16x3 kernel raises
    (include/moc_b200.h: moc_head_domain_flag_offset).  One per (device, stream); the flag is only ever cleared here,
    by the caller."""

    def __init__(self, device):
        lib = _lib.load()
        self.nbytes = lib.moc_head_forward_workspace_bytes()
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        off = lib.moc_head_domain_flag_offset()
        self.flag = self.buf[off:off + 4].view(torch.int32)

    def clear_flag(self) -> None:
        self.flag.zero_()

    def overflowed(self) -> bool:
        """Synchronises: call it where the results are read anyway."""
        return int(self.flag.item()) != 0


_HEAD_WS = {}


def head_workspace(device) -> HeadWorkspace:
    device = torch.device(device)
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _HEAD_WS.get(key)
    if ws is None:
        if len(_HEAD_WS) > 64:
            _HEAD_WS.clear()
        ws = _HEAD_WS[key] = HeadWorkspace(device)
    return ws


def head_forward(feat: torch.Tensor, keys: torch.Tensor, n_classes: int, sel: Selection, params: HeadParams,
                 active_mask: int, topk: int, want_gate: bool = False, wide: bool = False,
                 ws: Optional["HeadWorkspace"] = None) -> HeadOut:
    """Gate + combination + pooling of every selected row.  ``wide=True`` forces the range-free 3xTF32 gate kernel;
    otherwise ``HeadOut.domain_flag`` tells (after a sync) whether the FP16x3 kernel met a feature outside its range."""
    dev = feat.device
    cap = max(sel.capacity, 1)
    final = torch.empty(cap, n_classes, dtype=torch.float32, device=dev)
    gate = torch.empty(cap, GATES, dtype=torch.float32, device=dev) if want_gate else None
    bag = torch.empty(sel.n_slides, n_classes, dtype=torch.float32, device=dev)
    pos = torch.empty(sel.n_slides, n_classes, topk, dtype=torch.int32, device=dev)
    lib = _lib.load()
    if ws is None:
        ws = head_workspace(dev)
    mask = int(active_mask) | (_lib.HEAD_WIDE_DOMAIN if wide else 0)
    _count(3)
    check(lib.moc_head_forward(feat.data_ptr(), keys.data_ptr(), keys.stride(0), n_classes,
                                       sel.sel_base.data_ptr(), sel.sel_rows.data_ptr(), sel.sel_count.data_ptr(),
                                       sel.n_slides, sel.capacity, params.w1.data_ptr(), params.b1.data_ptr(),
                                       params.w2.data_ptr(), params.b2.data_ptr(), mask, int(topk),
                                       _ptr(gate), final.data_ptr(), bag.data_ptr(), pos.data_ptr(), ws.buf.data_ptr(),
                                       ws.nbytes, _stream()))
    return HeadOut(final, bag, pos, gate, ws.flag)


def cross_entropy(bag_logits: torch.Tensor, labels: torch.Tensor, grad_scale: float = 1.0, want_grad: bool = False,
                  want_pred: bool = False):
    """Per-slide CE (no temperature).  Returns (loss [n], dlogits [n,C] | None, pred int32 [n] | None)."""
    n, c = bag_logits.shape
    dev = bag_logits.device
    labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    loss = torch.empty(n, dtype=torch.float32, device=dev)
    dl = torch.empty(n, c, dtype=torch.float32, device=dev) if want_grad else None
    pred = torch.empty(n, dtype=torch.int32, device=dev) if want_pred else None
    _count(1)
    check(_lib.load().moc_cross_entropy(bag_logits.data_ptr(), labels.data_ptr(), n, c, float(grad_scale),
                                        loss.data_ptr(), _ptr(dl), _ptr(pred), _stream()))
    return loss, dl, pred


def head_backward(feat: torch.Tensor, keys: torch.Tensor, n_classes: int, sel: Selection, params: HeadParams,
                  active_mask: int, topk: int, pool_pos: torch.Tensor, dlogits: torch.Tensor,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Flat gradient [w1 | b1 | w2 | b2] (33 092 floats) of sum_i <dlogits_i, bag_logits_i>."""
    dev = feat.device
    lib = _lib.load()
    if out is None:
        out = torch.empty(NUM_PARAMS, dtype=torch.float32, device=dev)
    ws_bytes = lib.moc_head_backward_workspace_bytes(sel.n_slides, n_classes, topk)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    dlogits = _dev_f32(dlogits, "dlogits")
    _count(2)
    check(lib.moc_head_backward(feat.data_ptr(), keys.data_ptr(), keys.stride(0), n_classes, sel.sel_base.data_ptr(),
                                sel.sel_rows.data_ptr(), sel.sel_count.data_ptr(), sel.n_slides,
                                params.w1.data_ptr(), params.b1.data_ptr(), params.w2.data_ptr(), params.b2.data_ptr(),
                                int(active_mask), int(topk), pool_pos.data_ptr(), dlogits.data_ptr(), out.data_ptr(),
                                ws.data_ptr(), ws_bytes, _stream()))
    return out


def adam_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float = 1e-3, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
              weight_decay: float = 1e-4) -> None:
    """In-place torch.optim.Adam update of a flat fp32 parameter buffer; ``step`` counts from 1."""
    _count(1)
    check(_lib.load().moc_adam_step(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                    params.numel(), int(step), lr, beta1, beta2, eps, weight_decay, _stream()))


def adam_prepare_dev(step_dev: torch.Tensor, scalars: torch.Tensor, lr: float, beta1: float, beta2: float) -> None:
    """Advance the device step counter (int64 [1]) and derive the bias-correction scalars (float32 [2]) from it."""
    _count(1)
    check(_lib.load().moc_adam_prepare_dev(step_dev.data_ptr(), scalars.data_ptr(), lr, beta1, beta2, _stream()))


def adam_apply_dev(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor,
                   scalars: torch.Tensor, beta1: float, beta2: float, eps: float, weight_decay: float) -> None:
    _count(1)
    check(_lib.load().moc_adam_apply_dev(params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                                         params.numel(), scalars.data_ptr(), beta1, beta2, eps, weight_decay, _stream()))


def accumulate_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst += src for two contiguous fp32 CUDA buffers of the same length (micro-batch gradient sum)."""
    assert dst.is_cuda and src.is_cuda and dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
    _count(1)
    check(_lib.load().moc_accumulate(dst.data_ptr(), src.data_ptr(), dst.numel(), _stream()))
    return dst


def gather_selected(feat: torch.Tensor, keys: torch.Tensor, n_classes: int, sel_rows: torch.Tensor, n_sel: int,
                    want_feat: bool = True, want_planes: bool = True):
    """Dense selected_feat [S,512] and the four [S,C] score planes of main_moc.py:355-366."""
    dev = feat.device
    sf = torch.empty(n_sel, D, dtype=torch.float32, device=dev) if want_feat else None
    planes = [torch.empty(n_sel, n_classes, dtype=torch.float32, device=dev) for _ in range(4)] if want_planes else [None] * 4
    _count(1)
    check(_lib.load().moc_gather_selected(feat.data_ptr(), _ptr(keys), keys.stride(0) if keys is not None else 0,
                                          n_classes, sel_rows.data_ptr(), n_sel, _ptr(sf), _ptr(planes[0]),
                                          _ptr(planes[1]), _ptr(planes[2]), _ptr(planes[3]), _stream()))
    return sf, planes


def senet_forward(x: torch.Tensor, params: HeadParams, wide: Optional[bool] = None) -> torch.Tensor:
    """sigmoid(W2 relu(W1 x + b1) + b2) for a dense [rows,512] input.  ``wide=None`` (default): run the FP16x3 kernel,
    read its domain flag (one host sync - this is the module-level API, whose callers synchronise per slide anyway) and
    repeat on the range-free 3xTF32 kernel if a feature was outside |x| < 4094; True / False pick a kernel outright."""
    x = _dev_f32(x, "x")
    if x.dim() != 2 or x.size(1) != D:
        raise MocError(_lib.E_SHAPE, "senet input must be [rows,512], got %s" % (tuple(x.shape),))
    gate = torch.empty(x.size(0), GATES, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    ws = head_workspace(x.device)

    def run(flags):
        _count(2)
        check(lib.moc_senet_forward(x.data_ptr(), x.size(0), params.w1.data_ptr(), params.b1.data_ptr(),
                                    params.w2.data_ptr(), params.b2.data_ptr(), gate.data_ptr(), flags, ws.buf.data_ptr(),
                                    ws.nbytes, _stream()))

    if wide:
        run(_lib.HEAD_WIDE_DOMAIN)
        return gate
    if wide is None:
        ws.clear_flag()
    run(0)
    if wide is None and x.size(0) > 0 and ws.overflowed():
        run(_lib.HEAD_WIDE_DOMAIN)
    return gate


def senet_backward(x: torch.Tensor, dgate: torch.Tensor, params: HeadParams) -> torch.Tensor:
    x, dgate = _dev_f32(x, "x"), _dev_f32(dgate, "dgate")
    lib = _lib.load()
    n = x.size(0)
    out = torch.empty(NUM_PARAMS, dtype=torch.float32, device=x.device)
    if n == 0:
        return out.zero_()
    ws_bytes = lib.moc_senet_backward_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    _count(2)
    check(lib.moc_senet_backward(x.data_ptr(), n, dgate.data_ptr(), params.w1.data_ptr(), params.b1.data_ptr(),
                                 params.w2.data_ptr(), params.b2.data_ptr(), out.data_ptr(), ws.data_ptr(), ws_bytes,
                                 _stream()))
    return out


def split_grads(flat: torch.Tensor):
    """Views of the flat gradient in state_dict order: model.0.weight, model.0.bias, model.2.weight, model.2.bias."""
    o1 = HIDDEN * D
    o2 = o1 + HIDDEN
    o3 = o2 + GATES * HIDDEN
    return (flat[:o1].view(HIDDEN, D), flat[o1:o2], flat[o2:o3].view(GATES, HIDDEN), flat[o3:o3 + GATES])


def row_keys(logits: torch.Tensor, n_fg: int) -> torch.Tensor:
    """Key planes [2*n_fg+3, N] (the full layout, whatever the class count) from logits [N, Ct] (first n_fg columns are
    classes, the rest background): what the stand-alone selector / pooling functions slice."""
    x = _dev_f32(logits, "logits")
    n, ct = x.shape
    keys = torch.empty(2 * n_fg + 3, n, dtype=torch.float32, device=x.device)    # always the full layout
    _count(1)
    check(_lib.load().moc_row_keys(x.data_ptr(), n, x.stride(0), n_fg, ct, keys.data_ptr(), keys.stride(0), _stream()))
    return keys


def take_rows(src: torch.Tensor, idx: torch.Tensor, n_cols: int) -> torch.Tensor:
    """src[idx, :n_cols] for an int64 index vector."""
    src = _dev_f32(src, "src")
    idx = idx.to(device=src.device, dtype=torch.int64).contiguous()
    out = torch.empty(idx.numel(), n_cols, dtype=torch.float32, device=src.device)
    _count(1)
    check(_lib.load().moc_take_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), idx.numel(), n_cols,
                                    out.data_ptr(), _stream()))
    return out


def col_prefix_mean(vals: torch.Tensor, j: int) -> torch.Tensor:
    """vals[:j].mean(dim=0, keepdim=True) for a [J,C] tensor, summed in row order."""
    vals = _dev_f32(vals, "vals")
    out = torch.empty(1, vals.size(1), dtype=torch.float32, device=vals.device)
    _count(1)
    check(_lib.load().moc_col_prefix_mean(vals.data_ptr(), vals.stride(0), vals.size(1), int(j), out.data_ptr(), _stream()))
    return out


_ABLATION_MODES = {"avg": 0, "sum": 1, "max": 2}


def ablation_pool(keys: torch.Tensor, n_classes: int, sel: Selection, how: str, topk: int) -> torch.Tensor:
    """Bag logits [n_slides,C] of the un-gated avg / sum / max combination (main_moc.py:538-555)."""
    if how not in _ABLATION_MODES:
        raise MocError(_lib.E_ARG, "ablation_study must be one of avg, sum, max")
    dev = keys.device
    final = torch.empty(max(sel.capacity, 1), n_classes, dtype=torch.float32, device=dev)
    bag = torch.empty(sel.n_slides, n_classes, dtype=torch.float32, device=dev)
    _count(2)
    check(_lib.load().moc_ablation_forward(keys.data_ptr(), keys.stride(0), n_classes, sel.sel_base.data_ptr(),
                                           sel.sel_rows.data_ptr(), sel.sel_count.data_ptr(), sel.n_slides,
                                           sel.capacity, _ABLATION_MODES[how], int(topk), final.data_ptr(),
                                           bag.data_ptr(), _stream()))
    return bag


# ---------------------------------------------------------------------------------------------
# dense layers and row-wise pieces of the secondary MIL heads (moc_b200.mil_heads)
ACT = {None: 0, "none": 0, "relu": 1, "tanh": 2, "sigmoid": 3}


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
           split: Optional[int] = None, act_tail: Optional[str] = None) -> torch.Tensor:
    """act(x @ weight.T + bias) for a whole bag on the tensor cores (tcgen05 3xTF32).  Columns >= split use act_tail."""
    x = _dev_f32(x, "x")
    w = _dev_f32(weight, "weight")
    if x.dim() != 2 or w.dim() != 2 or x.size(1) != w.size(1):
        raise MocError(_lib.E_SHAPE, "linear: x [N,K] and weight [M,K] expected, got %s and %s" % (tuple(x.shape), tuple(w.shape)))
    n, k = x.shape
    m = w.size(0)
    b = _dev_f32(bias, "bias") if bias is not None else None
    ldy = (m + 3) & ~3
    y = torch.empty(n, ldy, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    ws_bytes = lib.moc_linear_workspace_bytes(m, k)
    if ws_bytes == 0:
        raise MocError(_lib.E_SHAPE, "linear: in_features must be a multiple of 32, got %d" % k)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    sp = m if split is None else int(split)
    _count(2)
    check(lib.moc_linear_forward(x.data_ptr(), x.stride(0), n, k, w.data_ptr(), _ptr(b), m, ACT[act], sp,
                                 ACT[act_tail if split is not None else act], y.data_ptr(), ldy, ws.data_ptr(), ws_bytes,
                                 _stream()))
    return y[:, :m]


def adapter_scores(x: torch.Tensor, adapted: Optional[torch.Tensor], clip_ratio: float, classifier: torch.Tensor
                   ) -> torch.Tensor:
    """Planes [C,N] of normalise(adapted * ratio + x * (1 - ratio)) @ classifier ([512,C]); adapted=None: normalise(x)."""
    x = _dev_f32(x, "x")
    cl = _dev_f32(classifier, "classifier")
    if x.dim() != 2 or x.size(1) != D or cl.dim() != 2 or cl.size(0) != D:
        raise MocError(_lib.E_SHAPE, "adapter_scores: x [N,512] and classifier [512,C] expected")
    a = None
    if adapted is not None:
        a = _dev_f32(adapted, "adapted")
        if a.shape != x.shape:
            raise MocError(_lib.E_SHAPE, "adapter_scores: adapted must have the shape of x")
    n, c = x.size(0), cl.size(1)
    out = torch.empty(c, max(n, 1), dtype=torch.float32, device=x.device)
    _count(1)
    check(_lib.load().moc_adapter_scores(x.data_ptr(), _ptr(a), float(clip_ratio), cl.data_ptr(), c, n, out.data_ptr(),
                                         out.stride(0), _stream()))
    return out[:, :n]


def gated_attention_scores(ab: torch.Tensor, hidden: int, wc: torch.Tensor, bc: float) -> torch.Tensor:
    """a_raw [N] from the stacked [tanh | sigmoid] branch activations ab [N, 2*hidden]."""
    n = ab.size(0)
    out = torch.empty(max(n, 1), dtype=torch.float32, device=ab.device)
    wc = _dev_f32(wc, "wc").reshape(-1)
    _count(1)
    check(_lib.load().moc_gated_attention_scores(ab.data_ptr(), ab.stride(0), int(hidden), wc.data_ptr(), float(bc), n,
                                                 out.data_ptr(), _stream()))
    return out[:n]


def attention_pool(a_raw: torch.Tensor, h: torch.Tensor, w_cls: torch.Tensor, b_cls: Optional[torch.Tensor]):
    """(pooled [1,L], logits [1,C], probs [1,C], y_hat int32 [1]) of softmax(a_raw) @ h through the bag classifier."""
    n, width = h.shape
    w_cls = _dev_f32(w_cls, "w_cls")
    c = w_cls.size(0)
    dev = h.device
    pooled = torch.empty(1, width, dtype=torch.float32, device=dev)
    logits = torch.empty(1, c, dtype=torch.float32, device=dev)
    probs = torch.empty(1, c, dtype=torch.float32, device=dev)
    yhat = torch.empty(1, dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws_bytes = lib.moc_attention_pool_workspace_bytes(n, width)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    b = _dev_f32(b_cls, "b_cls") if b_cls is not None else None
    _count(2)
    check(lib.moc_attention_pool(a_raw.data_ptr(), h.data_ptr(), h.stride(0), width, n, w_cls.data_ptr(), _ptr(b), c,
                                 pooled.data_ptr(), logits.data_ptr(), probs.data_ptr(), yhat.data_ptr(), ws.data_ptr(),
                                 ws_bytes, _stream()))
    return pooled, logits, probs, yhat


def row_softmax(logits: torch.Tensor) -> torch.Tensor:
    n, c = logits.shape
    out = torch.empty(n, c, dtype=torch.float32, device=logits.device)
    _count(1)
    check(_lib.load().moc_row_softmax(logits.data_ptr(), logits.stride(0), c, n, out.data_ptr(), c, _stream()))
    return out


def linear_wgrad(g: torch.Tensor, x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False
                 ) -> torch.Tensor:
    """dW [M,K] = g^T @ x for g = d(loss)/dy [N,M] and the layer input x [N,K] (tcgen05 3xTF32, deterministic)."""
    g = _dev_f32(g, "g")
    x = _dev_f32(x, "x")
    if g.dim() != 2 or x.dim() != 2 or g.size(0) != x.size(0):
        raise MocError(_lib.E_SHAPE, "linear_wgrad: g [N,M] and x [N,K] expected, got %s and %s" % (tuple(g.shape), tuple(x.shape)))
    n, m = g.shape
    k = x.size(1)
    if out is None:
        out = torch.empty(m, k, dtype=torch.float32, device=g.device)
        accumulate = False
    lib = _lib.load()
    ws_bytes = lib.moc_linear_wgrad_workspace_bytes(n, m, k)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=g.device)
    _count(2)
    check(lib.moc_linear_wgrad(g.data_ptr(), g.stride(0), m, x.data_ptr(), x.stride(0), k, n, out.data_ptr(), out.stride(0),
                               1 if accumulate else 0, ws.data_ptr(), ws_bytes, _stream()))
    return out


def abmil_backward(x: torch.Tensor, h1: torch.Tensor, ab: torch.Tensor, hidden: int, a_raw: torch.Tensor,
                   pooled: torch.Tensor, w_ab: torch.Tensor, wc: torch.Tensor, w_cls: torch.Tensor, dlogits: torch.Tensor):
    """Parameter gradients of ABMIL for one bag: (d_wfc, d_bfc, d_wab, d_bab, d_wc, d_bc, d_wcls, d_bcls)."""
    x = _dev_f32(x, "x")
    n, k_in = x.shape
    width = h1.size(1)
    c = w_cls.size(0)
    dev = x.device
    w_ab, wc, w_cls = _dev_f32(w_ab, "w_ab"), _dev_f32(wc, "wc").reshape(-1), _dev_f32(w_cls, "w_cls")
    dl = _dev_f32(dlogits, "dlogits").reshape(-1)
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    d_wfc, d_bfc, d_wab, d_bab = f(width, k_in), f(width), f(2 * hidden, width), f(2 * hidden)
    d_wc, d_bc, d_wcls, d_bcls = f(hidden), f(1), f(c, width), f(c)
    lib = _lib.load()
    ws_bytes = lib.moc_abmil_backward_workspace_bytes(n, k_in, width, hidden)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _count(14)
    check(lib.moc_abmil_backward(x.data_ptr(), x.stride(0), k_in, n, h1.data_ptr(), h1.stride(0), width, ab.data_ptr(),
                                 ab.stride(0), int(hidden), a_raw.data_ptr(), pooled.data_ptr(), w_ab.data_ptr(),
                                 wc.data_ptr(), w_cls.data_ptr(), c, dl.data_ptr(), d_wfc.data_ptr(), d_bfc.data_ptr(),
                                 d_wab.data_ptr(), d_bab.data_ptr(), d_wc.data_ptr(), d_bc.data_ptr(), d_wcls.data_ptr(),
                                 d_bcls.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    return d_wfc, d_bfc, d_wab, d_bab, d_wc, d_bc, d_wcls, d_bcls


def transpose(mat: torch.Tensor) -> torch.Tensor:
    """Contiguous transpose of a [R,C] fp32 CUDA matrix (our kernel, not a strided view)."""
    mat = _dev_f32(mat, "mat")
    r, c = mat.shape
    out = torch.empty(c, r, dtype=torch.float32, device=mat.device)
    _count(1)
    check(_lib.load().moc_transpose(mat.data_ptr(), r, c, out.data_ptr(), _stream()))
    return out


def adapter_backward_rows(x_rows: torch.Tensor, a2_rows: torch.Tensor, clip_ratio: float, classifier: torch.Tensor,
                          cls_of_row: torch.Tensor, g_of_row: torch.Tensor) -> torch.Tensor:
    """Gradient at the adapter output [R,512] of Conch_CLIP_Ada for R pooled (row, class) pairs."""
    x_rows, a2_rows = _dev_f32(x_rows, "x_rows"), _dev_f32(a2_rows, "a2_rows")
    cl = _dev_f32(classifier, "classifier")
    r = x_rows.size(0)
    out = torch.empty(r, D, dtype=torch.float32, device=x_rows.device)
    cls_of_row = cls_of_row.to(device=x_rows.device, dtype=torch.int32).contiguous()
    g_of_row = _dev_f32(g_of_row, "g_of_row")
    _count(1)
    check(_lib.load().moc_adapter_backward_rows(x_rows.data_ptr(), a2_rows.data_ptr(), float(clip_ratio), cl.data_ptr(),
                                                cl.size(1), cls_of_row.data_ptr(), g_of_row.data_ptr(), r, out.data_ptr(),
                                                _stream()))
    return out


def mask_positive_(g: torch.Tensor, ref: torch.Tensor) -> torch.Tensor:
    """In place: g[i] = 0 where ref[i] <= 0 (ReLU backward)."""
    assert g.is_contiguous() and ref.is_contiguous() and g.shape == ref.shape
    _count(1)
    check(_lib.load().moc_mask_positive(g.data_ptr(), ref.data_ptr(), g.numel(), _stream()))
    return g


def mil_fc_backward(x_row: torch.Tensor, hid_row: torch.Tensor, w_last: torch.Tensor, dtop: torch.Tensor):
    """(d_w0 [H1,K0], d_b0 [H1], d_wl [C,H1], d_bl [C]) of MIL_fc from the gradient at the selected instance's logits."""
    x_row, hid_row = _dev_f32(x_row, "x_row").reshape(-1), _dev_f32(hid_row, "hid_row").reshape(-1)
    w_last, dtop = _dev_f32(w_last, "w_last"), _dev_f32(dtop, "dtop").reshape(-1)
    k0, h1, c = x_row.numel(), hid_row.numel(), w_last.size(0)
    dev = x_row.device
    d_w0 = torch.empty(h1, k0, dtype=torch.float32, device=dev)
    d_b0 = torch.empty(h1, dtype=torch.float32, device=dev)
    d_wl = torch.empty(c, h1, dtype=torch.float32, device=dev)
    d_bl = torch.empty(c, dtype=torch.float32, device=dev)
    _count(1)
    check(_lib.load().moc_mil_fc_backward(x_row.data_ptr(), k0, hid_row.data_ptr(), h1, w_last.data_ptr(), c, dtop.data_ptr(),
                                          d_w0.data_ptr(), d_b0.data_ptr(), d_wl.data_ptr(), d_bl.data_ptr(), _stream()))
    return d_w0, d_b0, d_wl, d_bl
