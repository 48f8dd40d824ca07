"""The meta-learner ``senet`` with the reference's module surface (main_moc.py:299-312).

``state_dict`` keys are ``model.{0,2}.{weight,bias}`` so checkpoints interchange with the reference's
``best_model_shot_*_fold_*.pt`` (main_moc.py:628).  ``forward`` runs the CUDA gate kernel and is differentiable
with respect to the parameters (the only trainable tensors on the MOC path); inputs are constants as in the
reference, where ``selected_feat`` never requires grad.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _GateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        p = ops.HeadParams(w1.detach().contiguous(), b1.detach().contiguous(), w2.detach().contiguous(),
                           b2.detach().contiguous())
        ctx.save_for_backward(x, p.w1, p.b1, p.w2, p.b2)
        return ops.senet_forward(x, p)

    @staticmethod
    def backward(ctx, dgate):
        x, w1, b1, w2, b2 = ctx.saved_tensors
        flat = ops.senet_backward(x, dgate.contiguous(), ops.HeadParams(w1, b1, w2, b2))
        g1, gb1, g2, gb2 = ops.split_grads(flat)
        return None, g1, gb1, g2, gb2


class senet(nn.Module):
    """512 -> 64 -> ReLU -> 4 -> Sigmoid; one gate per classifier of the bank."""

    def __init__(self, in_dim: int = 512, out_dim: int = 4):
        super().__init__()
        if in_dim != ops.D or out_dim != ops.GATES:
            raise ValueError("this build of moc_b200 implements senet(512, 4) only (main_moc.py:315)")
        self.hidden_dim = ops.HIDDEN
        # same construction order as the reference, so a seeded init draws identical weights
        self.model = nn.Sequential(
            nn.Linear(in_dim, self.hidden_dim),
            nn.ReLU(),
            nn.Linear(self.hidden_dim, out_dim),
            nn.Sigmoid(),
        )

    def head_params(self) -> ops.HeadParams:
        l1, l2 = self.model[0], self.model[2]
        return ops.HeadParams(l1.weight.detach().contiguous(), l1.bias.detach().contiguous(),
                              l2.weight.detach().contiguous(), l2.bias.detach().contiguous())

    def parameters_in_order(self):
        l1, l2 = self.model[0], self.model[2]
        return [l1.weight, l1.bias, l2.weight, l2.bias]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        l1, l2 = self.model[0], self.model[2]
        if not l1.weight.is_cuda:
            raise ops.MocError(-1, "senet lives on a CUDA device in moc_b200 (call .to('cuda')); there is no CPU path")
        x = x.to(l1.weight.device)
        return _GateFn.apply(x, l1.weight, l1.bias, l2.weight, l2.bias)
