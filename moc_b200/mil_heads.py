"""The secondary MIL heads north_star names next to MOC proper, with the reference's module surface.

None of them is instantiated by a shipped driver of the reference, so they are built
with the reference's attribute names: parameters live in ordinary ``nn.Linear`` containers (their ``state_dict``s
interchange), ``forward`` runs every dense layer on the tensor cores (``ops.linear``: tcgen05 3xTF32) and
everything else in our row kernels.  ABMIL also trains: with gradients enabled its ``forward`` returns logits that
carry a graph whose backward is ours (``ops.abmil_backward``: tcgen05 weight gradients, deterministic reductions),
so the reference's ``loss.backward(); optimizer.step()`` loop (utils/core_utils.py:398-416) runs unchanged.  The
other two heads train the same way; their gradient lives on a few rows only (the top-j rows of each class for
``Conch_CLIP_Ada``, the one max-probability instance for ``MIL_fc``), which the backward gathers first.

* ``Conch_CLIP_Ada``  models/model_adapters.py:148-215 - adapter MLP, residual blend, normalise, score, top-j mean
* ``CLAM_SB``         models/model_clam.py:77-219 with ``instance_loss_fn=None`` (= ABMIL): gated attention pooling
* ``MIL_fc``          models/model_mil.py:11-51 - max-probability instance
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import MocError, E_ARG, E_SHAPE


def _pool_topj_planes(planes: torch.Tensor, topj: int) -> torch.Tensor:
    """[C,N] planes -> [1,C]: mean of the min(topj, N) largest per class (Conch_CLIP_Ada.topj_pooling)."""
    c, n = planes.shape
    offs = torch.tensor([0, n], dtype=torch.int64, device=planes.device)
    return ops.pool_topk(planes, offs, 1, c, int(topj), 0, 1, 0, 1, False)



def _abmil_forward(h, w_fc, b_fc, w_a, b_a, w_b, b_b, w_c, b_c, w_cls, b_cls):
    """CLAM_SB.forward_single (models/model_clam.py:175-212) on our kernels; also returns what the backward needs."""
    d = w_a.size(0)
    hh = ops.linear(h, w_fc, b_fc, "relu")
    w_ab = torch.cat([w_a, w_b], dim=0)
    b_ab = torch.cat([b_a, b_b], dim=0)
    ab = ops.linear(hh, w_ab, b_ab, "tanh", split=d, act_tail="sigmoid")
    a_raw = ops.gated_attention_scores(ab, d, w_c, float(b_c))
    pooled, logits, probs, yhat = ops.attention_pool(a_raw, hh, w_cls, b_cls)
    return logits, probs, yhat.to(torch.int64).view(1, 1), a_raw.unsqueeze(0), pooled, hh, ab, w_ab


class _AbmilFunction(torch.autograd.Function):
    """logits = ABMIL(h; parameters) with the hand-written backward (ops.abmil_backward): gradients for the ten
    parameter tensors, none for the bag (the reference never differentiates with respect to the features)."""

    @staticmethod
    def forward(ctx, h, w_fc, b_fc, w_a, b_a, w_b, b_b, w_c, b_c, w_cls, b_cls):
        logits, probs, yhat, a_row, pooled, hh, ab, w_ab = _abmil_forward(h, w_fc, b_fc, w_a, b_a, w_b, b_b, w_c, b_c,
                                                                          w_cls, b_cls)
        ctx.save_for_backward(h, hh, ab, a_row, pooled, w_ab, w_c, w_cls)
        ctx.hidden = w_a.size(0)
        ctx.mark_non_differentiable(probs, yhat, a_row, pooled)
        return logits, probs, yhat, a_row, pooled

    @staticmethod
    def backward(ctx, dlogits, *unused):
        h, hh, ab, a_row, pooled, w_ab, w_c, w_cls = ctx.saved_tensors
        d = ctx.hidden
        d_wfc, d_bfc, d_wab, d_bab, d_wc, d_bc, d_wcls, d_bcls = ops.abmil_backward(
            h, hh, ab, d, a_row.reshape(-1), pooled, w_ab, w_c, w_cls, dlogits.contiguous())
        return (None, d_wfc, d_bfc, d_wab[:d], d_bab[:d], d_wab[d:], d_bab[d:], d_wc.view(1, d), d_bc, d_wcls, d_bcls)



class _ClipAdaFunction(torch.autograd.Function):
    """pooled logits = Conch_CLIP_Ada.forward(feat; adapter weights) with the hand-written backward: only the top-j rows
    of each class carry gradient, so the backward gathers those <= j*C rows and runs the adapter's two layers backwards
    on them (row kernel + tensor-core weight gradients)."""

    @staticmethod
    def forward(ctx, feat, w_a, w_b, classifier, clip_ratio, topj):
        a1 = ops.linear(feat, w_a, None, "relu")
        a2 = ops.linear(a1, w_b, None, "relu")
        planes = ops.adapter_scores(feat, a2, clip_ratio, classifier)          # [C, N]
        pooled = _pool_topj_planes(planes, topj)
        ctx.save_for_backward(feat, a1, a2, planes, w_b, classifier)
        ctx.clip_ratio, ctx.topj = float(clip_ratio), int(topj)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        feat, a1, a2, planes, w_b, classifier = ctx.saved_tensors
        c, n = planes.shape
        j = min(ctx.topj, n)
        idx = ops.topj_sorted(planes.t(), j, largest=True)                       # int64 [j, C]: the pooled rows
        rows = idx.t().reshape(-1)                                               # class after class
        cls_of_row = torch.arange(c, device=feat.device, dtype=torch.int32).repeat_interleave(j)
        g_of_row = (dpooled.reshape(-1).float() / j).repeat_interleave(j)
        x_r, a1_r, a2_r = ops.take_rows(feat, rows, feat.size(1)), ops.take_rows(a1, rows, a1.size(1)), ops.take_rows(a2, rows, a2.size(1))
        da2 = ops.adapter_backward_rows(x_r, a2_r, ctx.clip_ratio, classifier, cls_of_row, g_of_row)
        d_wb = ops.linear_wgrad(da2, a1_r)                                       # [512, 128]
        da1 = ops.mask_positive_(ops.linear(da2, ops.transpose(w_b), None, None).contiguous(), a1_r)
        d_wa = ops.linear_wgrad(da1, x_r)                                        # [128, 512]
        return None, d_wa, d_wb, None, None, None


class _MilFcFunction(torch.autograd.Function):
    """top_instance logits = MIL_fc(h; parameters): the gradient reaches the parameters through the one selected
    instance only (models/model_mil.py:40-42)."""

    @staticmethod
    def forward(ctx, h, w0, b0, wl, bl):
        hid = ops.linear(h, w0, b0, "relu")
        logits = ops.linear(hid, wl, bl, None).contiguous()
        y_probs = ops.row_softmax(logits)
        top_idx = ops.topj_sorted(y_probs[:, 1], 1, largest=True).view(1,)
        top_instance = ops.take_rows(logits, top_idx, logits.size(1))
        hid_row = ops.take_rows(hid, top_idx, hid.size(1))
        x_row = ops.take_rows(h, top_idx, h.size(1))
        ctx.save_for_backward(x_row, hid_row, wl)
        ctx.mark_non_differentiable(y_probs, top_idx, hid_row)
        return top_instance, y_probs, top_idx, hid_row

    @staticmethod
    def backward(ctx, dtop, *unused):
        x_row, hid_row, wl = ctx.saved_tensors
        d_w0, d_b0, d_wl, d_bl = ops.mil_fc_backward(x_row, hid_row, wl, dtop.contiguous())
        return None, d_w0, d_b0, d_wl, d_bl


class Conch_CLIP_Ada(nn.Module):
    def __init__(self, c_in=512, reduction=4, num_classes=2, classifier_tensor=None, clip_ratio=0.1, topj=10):
        super().__init__()
        if c_in != ops.D:
            raise ValueError("this build of moc_b200 implements c_in=512 (CONCH embeddings) only")
        self.adapter = nn.Sequential(
            nn.Linear(c_in, c_in // reduction, bias=False),
            nn.ReLU(inplace=True),
            nn.Linear(c_in // reduction, c_in, bias=False),
            nn.ReLU(inplace=True),
        )
        self.topj = topj
        nn.init.kaiming_normal_(self.adapter[0].weight, a=np.sqrt(5))
        nn.init.kaiming_normal_(self.adapter[2].weight, a=np.sqrt(5))
        self.classifier = classifier_tensor
        self.num_classes = num_classes
        self.clip_ratio = clip_ratio

    def topj_pooling(self, logits, topj=10):
        """logits [N,C] -> [1,C] (models/model_adapters.py:173-183)."""
        with torch.no_grad():
            return _pool_topj_planes(logits.t().contiguous().float(), topj)

    def forward(self, feat):
        w_a, w_b = self.adapter[0].weight, self.adapter[2].weight
        if torch.is_grad_enabled() and (w_a.requires_grad or w_b.requires_grad):
            # adapter training: the pooled logits carry a graph whose backward is ours
            return _ClipAdaFunction.apply(feat, w_a, w_b, self.classifier, self.clip_ratio, self.topj)
        with torch.no_grad():
            a1 = ops.linear(feat, w_a, None, "relu")
            a2 = ops.linear(a1, w_b, None, "relu")
            planes = ops.adapter_scores(feat, a2, self.clip_ratio, self.classifier)
            return _pool_topj_planes(planes, self.topj)

    @torch.no_grad()
    def forward_disable_ada(self, feat):
        planes = ops.adapter_scores(feat, None, 0.0, self.classifier)
        return _pool_topj_planes(planes, self.topj)


class Attn_Net_Gated(nn.Module):
    """Parameter container with the reference's names (models/model_clam.py:41-64)."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super().__init__()
        a, b = [nn.Linear(L, D), nn.Tanh()], [nn.Linear(L, D), nn.Sigmoid()]
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(D, n_classes)

    @torch.no_grad()
    def forward(self, x):
        d = self.attention_a[0].out_features
        if self.attention_c.out_features != 1:
            raise MocError(E_SHAPE, "Attn_Net_Gated: only the single-branch attention (n_classes=1) is implemented")
        w = torch.cat([self.attention_a[0].weight, self.attention_b[0].weight], dim=0)
        b = torch.cat([self.attention_a[0].bias, self.attention_b[0].bias], dim=0)
        ab = ops.linear(x, w, b, "tanh", split=d, act_tail="sigmoid")
        a_raw = ops.gated_attention_scores(ab, d, self.attention_c.weight, float(self.attention_c.bias))
        return a_raw.unsqueeze(1), x


class CLAM_SB(nn.Module):
    def __init__(self, gate=True, size_arg="small", dropout=False, k_sample=8, n_classes=2,
                 instance_loss_fn=None, subtyping=False, conch_init=False, conch_freeze=False):
        super().__init__()
        self.size_dict = {"small": [1024, 512, 256], "big": [1024, 512, 384], "benchmark": [384, 512, 256],
                          "conch": [512, 512, 384], "gigapath": [1536, 512, 256], "virchow": [2560, 512, 256]}
        if not gate:
            raise ValueError("moc_b200.CLAM_SB implements the gated attention network (gate=True) only")
        if conch_init:
            raise ValueError("conch_init loads a checkpoint from the authors' home directory; load a state_dict instead")
        size = self.size_dict[size_arg]
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout:
            fc.append(nn.Dropout(0.25))
        fc.append(Attn_Net_Gated(L=size[1], D=size[2], dropout=dropout, n_classes=1))
        self.attention_net = nn.Sequential(*fc)
        self.classifiers = nn.Linear(size[1], n_classes)
        self.instance_classifiers = nn.ModuleList([nn.Linear(size[1], 2) for _ in range(n_classes)])
        self.k_sample = k_sample
        self.instance_loss_fn = instance_loss_fn
        self.n_classes = n_classes
        self.subtyping = subtyping
        for m in self.modules():  # utils/utils.py:399-403
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                m.bias.data.zero_()

    def relocate(self):
        self.to(torch.device("cuda"))

    def _params(self):
        att = self.attention_net[-1]
        fc = self.attention_net[0]
        return (fc.weight, fc.bias, att.attention_a[0].weight, att.attention_a[0].bias, att.attention_b[0].weight,
                att.attention_b[0].bias, att.attention_c.weight, att.attention_c.bias, self.classifiers.weight,
                self.classifiers.bias)

    def forward_single(self, h, label=None, instance_eval=False, return_features=False, attention_only=False):
        if instance_eval:
            raise MocError(E_ARG, "instance-level clustering (instance_eval=True) is outside the ABMIL path built here")
        if self.training and any(isinstance(m, nn.Dropout) for m in self.modules()):
            raise MocError(E_ARG, "dropout in training mode is not implemented: build the model with dropout=False or call .eval()")
        if self.attention_net[-1].attention_c.out_features != 1:
            raise MocError(E_SHAPE, "Attn_Net_Gated: only the single-branch attention (n_classes=1) is implemented")
        params = self._params()
        if attention_only:
            with torch.no_grad():
                return _abmil_forward(h, *params)[3]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # training (utils/core_utils.py:391-416): logits carry the graph; loss.backward() runs our backward kernels
            logits, probs, yhat, a_row, pooled = _AbmilFunction.apply(h, *params)
        else:
            with torch.no_grad():
                logits, probs, yhat, a_row, pooled = _abmil_forward(h, *params)[:5]
        results = {}
        if return_features:
            results["features"] = pooled
        return logits, probs, yhat, a_row, results

    def forward(self, h, label=None, instance_eval=False, return_features=False, attention_only=False):
        if h.dim() == 3:
            outs = [self.forward_single(h[i], None, instance_eval, return_features, attention_only)
                    for i in range(h.shape[0])]
            if attention_only:
                return torch.stack(outs, dim=0)
            return tuple(torch.stack([o[k] for o in outs], dim=0) for k in range(4)) + ([o[4] for o in outs],)
        return self.forward_single(h, label, instance_eval, return_features, attention_only)


class MIL_fc(nn.Module):
    def __init__(self, gate=True, size_arg="benchmark", dropout=False, n_classes=2, top_k=1):
        super().__init__()
        assert n_classes == 2
        self.size_dict = {"small": [1024, 512], "benchmark": [384, 512]}
        size = self.size_dict[size_arg]
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout:
            fc.append(nn.Dropout(0.25))
        fc.append(nn.Linear(size[1], n_classes))
        self.classifier = nn.Sequential(*fc)
        self.top_k = top_k

    def relocate(self):
        self.classifier.to(torch.device("cuda"))

    def forward(self, h, return_features=False):
        if self.top_k != 1:
            raise MocError(E_ARG, "MIL_fc: the reference's .view(1,) admits top_k=1 only (models/model_mil.py:40)")
        if self.training and any(isinstance(m, nn.Dropout) for m in self.modules()):
            raise MocError(E_ARG, "dropout in training mode is not implemented: build the model with dropout=False or call .eval()")
        l0, l1 = self.classifier[0], self.classifier[-1]
        params = (l0.weight, l0.bias, l1.weight, l1.bias)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            top_instance, y_probs, top_idx, hid_row = _MilFcFunction.apply(h, *params)
        else:
            with torch.no_grad():
                top_instance, y_probs, top_idx, hid_row = _MilFcFunction.forward(_NoCtx(), h, *params)
        with torch.no_grad():
            y_prob = ops.row_softmax(top_instance.detach())
            y_hat = ops.topj_sorted(top_instance.detach().t().contiguous(), 1, largest=True).view(1, 1)
        results = {}
        if return_features:
            results["features"] = hid_row
        return top_instance, y_prob, y_hat, y_probs, results


class _NoCtx:
    """Stand-in for the autograd context when a Function's forward is used without a graph."""

    def save_for_backward(self, *a):
        pass

    def mark_non_differentiable(self, *a):
        pass
