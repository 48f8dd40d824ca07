"""CPU restatement of the secondary MIL heads  --  TEST INFRASTRUCTURE, not a product path.

Plain torch fp32 on the CPU, one function per reference forward, each citing the lines it follows.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import this.  Pinned by tests/golden/heads_*.npz, which
oracle/make_golden_heads.py produced by running the reference's own modules (models/model_adapters.py,
models/model_clam.py, models/model_mil.py) on seeded inputs.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def topj_mean(logits: torch.Tensor, topj: int) -> torch.Tensor:
    """Conch_CLIP_Ada.topj_pooling (models/model_adapters.py:173-183)."""
    maxj = min(topj, logits.size(0))
    values, _ = logits.topk(maxj, 0, True, True)
    return values[:min(topj, maxj)].mean(dim=0, keepdim=True)


def clip_ada_forward(sd: Dict[str, torch.Tensor], classifier: torch.Tensor, feat: torch.Tensor, clip_ratio: float,
                     topj: int) -> torch.Tensor:
    """Conch_CLIP_Ada.forward (models/model_adapters.py:185-193): adapter = Linear(512,128,no bias) ReLU
    Linear(128,512,no bias) ReLU (:152-157); blend (:187); normalise (:188); score (:191); top-j mean (:192)."""
    a = F.relu(F.linear(F.relu(F.linear(feat, sd["adapter.0.weight"])), sd["adapter.2.weight"]))
    f = a * clip_ratio + feat * (1 - clip_ratio)
    f = f / f.norm(dim=-1, keepdim=True)
    return topj_mean(f @ classifier, topj)


def clip_ada_forward_disable_ada(classifier: torch.Tensor, feat: torch.Tensor, topj: int) -> torch.Tensor:
    """Conch_CLIP_Ada.forward_disable_ada (models/model_adapters.py:210-215)."""
    f = feat / feat.norm(dim=-1, keepdim=True)
    return topj_mean(f @ classifier, topj)


def abmil_forward(sd: Dict[str, torch.Tensor], h: torch.Tensor, att: str = "attention_net.2."):
    """CLAM_SB.forward_single without instance evaluation (models/model_clam.py:175-219) on the gated attention
    network (:41-64): returns (logits, Y_prob, Y_hat, A_raw, pooled)."""
    hh = F.relu(F.linear(h, sd["attention_net.0.weight"], sd["attention_net.0.bias"]))            # :83, :177
    a = torch.tanh(F.linear(hh, sd[att + "attention_a.0.weight"], sd[att + "attention_a.0.bias"]))  # :44-46, :59
    b = torch.sigmoid(F.linear(hh, sd[att + "attention_b.0.weight"], sd[att + "attention_b.0.bias"]))  # :48-49, :60
    A = F.linear(a * b, sd[att + "attention_c.weight"], sd[att + "attention_c.bias"])            # :61-62
    A_raw = A.t()                                                                                 # :178, :181
    M = torch.mm(F.softmax(A_raw, dim=1), hh)                                                     # :182, :209
    logits = F.linear(M, sd["classifiers.weight"], sd["classifiers.bias"])                        # :210
    return logits, F.softmax(logits, dim=1), torch.topk(logits, 1, dim=1)[1], A_raw, M           # :211-212


def mil_fc_forward(sd: Dict[str, torch.Tensor], h: torch.Tensor):
    """MIL_fc.forward (models/model_mil.py:30-51) with top_k = 1: returns (top_instance, Y_prob, Y_hat, y_probs)."""
    keys = sorted(k for k in sd if k.endswith("weight"))
    k0, k1 = keys[0], keys[-1]
    hid = F.relu(F.linear(h, sd[k0], sd[k0.replace("weight", "bias")]))
    logits = F.linear(hid, sd[k1], sd[k1.replace("weight", "bias")])     # :35
    y_probs = F.softmax(logits, dim=1)                                    # :38
    idx = torch.topk(y_probs[:, 1], 1, dim=0)[1].view(1,)                 # :40
    top = torch.index_select(logits, dim=0, index=idx)                    # :42
    return top, F.softmax(top, dim=1), torch.topk(top, 1, dim=1)[1], y_probs   # :44-45


def abmil_loss_and_grads(sd: Dict[str, torch.Tensor], h: torch.Tensor, label: int):
    """One ABMIL training step's loss and parameter gradients, as utils/core_utils.py:391-414 obtains them:
    logits = model(data); loss = CrossEntropyLoss()(logits, label); loss.backward()  -  torch autograd over the
    restated forward above.  Returns (loss, {state_dict key: gradient})."""
    names = [k for k in sd if not k.startswith("instance_classifiers")]
    leaf = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    logits = abmil_forward(leaf, h)[0]
    loss = F.cross_entropy(logits, torch.tensor([int(label)]))
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), dict(zip(names, grads))


def clip_ada_loss_and_grads(sd: Dict[str, torch.Tensor], classifier: torch.Tensor, feat: torch.Tensor, clip_ratio: float,
                            topj: int, label: int, logit_scale: float = 56.3477):
    """Adapter training step of Conch_CLIP_Ada: loss = CE(logit_scale * forward(feat), label) (the logit scale of
    models/model_adapters.py:190), gradients of the two adapter weights by torch autograd over the restated forward."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss = F.cross_entropy(clip_ada_forward(leaf, classifier, feat, clip_ratio, topj) * logit_scale, torch.tensor([int(label)]))
    names = list(leaf)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), dict(zip(names, grads))


def mil_fc_loss_and_grads(sd: Dict[str, torch.Tensor], h: torch.Tensor, label: int):
    """MIL_fc training step as utils/core_utils.py:391-414 runs it: loss = CE(top_instance, label); autograd."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    loss = F.cross_entropy(mil_fc_forward(leaf, h)[0], torch.tensor([int(label)]))
    names = list(leaf)
    grads = torch.autograd.grad(loss, [leaf[k] for k in names])
    return loss.detach(), dict(zip(names, grads))
