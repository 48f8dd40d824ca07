"""Golden vectors for the secondary MIL heads from the reference's own modules  --  TEST INFRASTRUCTURE.

    python oracle/make_golden_heads.py         (build container only: needs /root/reference)

Imports models/model_adapters.py, models/model_clam.py and models/model_mil.py UNMODIFIED from the read-only
checkout (with empty stand-ins for the absent third-party imports they make at module level: openslide, the CONCH
model factory, nystrom_attention), instantiates Conch_CLIP_Ada, CLAM_SB (instance_loss_fn=None = ABMIL) and MIL_fc
with seeded initialisation, runs their forward on seeded bags and stores inputs, parameters and outputs in
tests/golden/heads_*.npz.  ``--backward`` writes only the ABMIL training-step gradients (heads_abmil_bwd_c*.npz).
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


def import_reference_models():
    ref = ref_loader.REFERENCE_ROOT
    for name in ("openslide", "nystrom_attention"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["nystrom_attention"].NystromAttention = object
    sys.path.insert(0, ref)
    stub = types.ModuleType("models.model_conch")
    stub.conch_coca = stub.conch_lora = lambda *a, **k: None
    import importlib
    models = importlib.import_module("models")
    sys.modules["models.model_conch"] = stub
    models.model_conch = stub
    clam = importlib.import_module("models.model_clam")
    mil = importlib.import_module("models.model_mil")
    ada = importlib.import_module("models.model_adapters")
    return clam, mil, ada


def sd_np(sd, prefix):
    """Parameters were rounded to fp16-representable values before the reference ran (fp16_exact), so storing
    them as fp16 loses nothing and halves the fixtures."""
    return {prefix + k: v.detach().half().numpy().copy() for k, v in sd.items()}


def fp16_exact(module):
    with torch.no_grad():
        for p_ in module.parameters():
            p_.copy_(p_.half().float())


def main():
    assert ref_loader.has_checkout()
    torch.set_num_threads(1)
    clam, mil, ada = import_reference_models()
    out_all = {}

    # ---- Conch_CLIP_Ada (C = 2 and 3), bags of 37 / 700 patches (N < topj and N > topj), un-normalised inputs
    for c, sizes, topj, seed in ((2, [37, 700], 50, 31), (3, [300], 10, 32)):
        w, _ = synthetic.prompt_matrices(c)
        torch.manual_seed(seed)
        m = ada.Conch_CLIP_Ada(c_in=512, reduction=4, num_classes=c, classifier_tensor=w, clip_ratio=0.1, topj=topj).eval()
        fp16_exact(m)
        out = {"C": c, "topj": topj, "clip_ratio": 0.1, "classifier": w.numpy(), "n_bags": len(sizes)}
        out.update(sd_np(m.state_dict(), "sd_"))
        bags, _ = synthetic.make_cohort(len(sizes), sizes, c, cohort_seed=seed)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for i, x in enumerate(bags):
                x = (x * (0.5 + 2.0 * torch.rand(x.size(0), 1, generator=g))).half().float()
                out["feat_%d" % i] = x.half().numpy()
                out["forward_%d" % i] = m.forward(x).numpy()
                out["forward_disable_ada_%d" % i] = m.forward_disable_ada(x).numpy()
        np.savez_compressed(os.path.join(OUT, "heads_clip_ada_c%d.npz" % c), **out)
        print("wrote heads_clip_ada_c%d" % c)

    # ---- ABMIL = CLAM_SB(size_arg="conch", instance_loss_fn=None)
    for c, sizes, seed in ((2, [65, 900], 41), (3, [513], 42)):
        torch.manual_seed(seed)
        m = clam.CLAM_SB(gate=True, size_arg="conch", dropout=False, n_classes=c, instance_loss_fn=None).eval()
        # initialize_weights zeroes every bias; give them values so the bias paths are exercised
        with torch.no_grad():
            for p_ in m.parameters():
                if p_.dim() == 1:
                    p_.copy_(0.05 * torch.randn(p_.shape))
        fp16_exact(m)
        out = {"C": c, "n_bags": len(sizes)}
        out.update(sd_np(m.state_dict(), "sd_"))
        bags, _ = synthetic.make_cohort(len(sizes), sizes, c, cohort_seed=seed)
        with torch.no_grad():
            for i, x in enumerate(bags):
                x = (x * 4.0).half().float()
                out["feat_%d" % i] = x.half().numpy()
                logits, y_prob, y_hat, a_raw, res = m(x, return_features=True)
                out["logits_%d" % i], out["y_prob_%d" % i] = logits.numpy(), y_prob.numpy()
                out["y_hat_%d" % i], out["a_raw_%d" % i] = y_hat.numpy(), a_raw.numpy()
                out["features_%d" % i] = res["features"].numpy()
                out["attention_only_%d" % i] = m(x, attention_only=True).numpy()
        np.savez_compressed(os.path.join(OUT, "heads_abmil_c%d.npz" % c), **out)
        print("wrote heads_abmil_c%d" % c)

    # ---- MIL_fc (384-d "benchmark" inputs, 2 classes)
    torch.manual_seed(51)
    m = mil.MIL_fc(size_arg="benchmark", dropout=False, n_classes=2, top_k=1).eval()
    fp16_exact(m)
    out = {"n_bags": 2}
    out.update(sd_np(m.state_dict(), "sd_"))
    g = torch.Generator().manual_seed(52)
    with torch.no_grad():
        for i, n in enumerate((40, 777)):
            x = torch.randn(n, 384, generator=g).half().float()
            out["feat_%d" % i] = x.half().numpy()
            top, y_prob, y_hat, y_probs, _ = m(x)
            out["top_instance_%d" % i], out["y_prob_%d" % i] = top.numpy(), y_prob.numpy()
            out["y_hat_%d" % i], out["y_probs_%d" % i] = y_hat.numpy(), y_probs.numpy()
    np.savez_compressed(os.path.join(OUT, "heads_mil_fc.npz"), **out)
    print("wrote heads_mil_fc")


def main_backward():
    """ABMIL training step through the reference's own module and torch autograd (utils/core_utils.py:391-416:
    logits = model(data); loss = CE(logits, label); loss.backward()): parameter gradients of the last bag of
    each heads_abmil_c*.npz configuration, rebuilt from the same seeds (checked against the stored parameters).
    Stored in heads_abmil_bwd_c*.npz: loss, every bias / small gradient in full, the three large matrices in full
    for C=2 and as every 5th element for C=3."""
    assert ref_loader.has_checkout()
    torch.set_num_threads(1)
    clam, _, _ = import_reference_models()
    for c, sizes, seed in ((2, [65, 900], 41), (3, [513], 42)):
        torch.manual_seed(seed)
        m = clam.CLAM_SB(gate=True, size_arg="conch", dropout=False, n_classes=c, instance_loss_fn=None).train()
        with torch.no_grad():
            for p_ in m.parameters():
                if p_.dim() == 1:
                    p_.copy_(0.05 * torch.randn(p_.shape))
        fp16_exact(m)
        stored = np.load(os.path.join(OUT, "heads_abmil_c%d.npz" % c))
        for k, v in m.state_dict().items():
            assert np.array_equal(v.half().numpy(), stored["sd_" + k]), k
        i = len(sizes) - 1
        x = torch.from_numpy(stored["feat_%d" % i]).float()
        label = torch.tensor([(i + 1) % c])
        logits, _, _, _, _ = m(x)
        assert np.array_equal(logits.detach().numpy(), stored["logits_%d" % i])
        loss = torch.nn.CrossEntropyLoss()(logits, label)
        loss.backward()
        out = {"bag": i, "label": int(label), "loss": float(loss)}
        for k, p_ in m.named_parameters():
            if k.startswith("instance_classifiers"):
                assert p_.grad is None
                continue
            g = p_.grad.numpy()
            if c == 3 and g.size > 4096:
                out["grad5_" + k] = g.reshape(-1)[::5].copy()
                out["norm_" + k] = np.float64(np.linalg.norm(g.astype(np.float64)))
            else:
                out["grad_" + k] = g
        np.savez_compressed(os.path.join(OUT, "heads_abmil_bwd_c%d.npz" % c), **out)
        print("wrote heads_abmil_bwd_c%d  loss %.6f" % (c, float(loss)))


def main_backward_sparse():
    """Training-step gradients of the two heads whose gradient lives on a few rows, through the reference's own modules
    and torch autograd: Conch_CLIP_Ada (both bags of heads_clip_ada_c2/c3: fewer rows than topj, and more) and MIL_fc.
    Parameters and inputs are rebuilt from the seeds of main() and checked against the stored forward goldens."""
    assert ref_loader.has_checkout()
    torch.set_num_threads(1)
    _, mil, ada = import_reference_models()
    for c, sizes, topj, seed in ((2, [37, 700], 50, 31), (3, [300], 10, 32)):
        w, _ = synthetic.prompt_matrices(c)
        torch.manual_seed(seed)
        m = ada.Conch_CLIP_Ada(c_in=512, reduction=4, num_classes=c, classifier_tensor=w, clip_ratio=0.1, topj=topj).train()
        fp16_exact(m)
        stored = np.load(os.path.join(OUT, "heads_clip_ada_c%d.npz" % c))
        out = {"n_bags": len(sizes)}
        for i in range(len(sizes)):
            x = torch.from_numpy(stored["feat_%d" % i]).float()
            m.zero_grad()
            logits = m.forward(x)
            assert np.array_equal(logits.detach().numpy(), stored["forward_%d" % i])
            label = torch.tensor([(i + 1) % c])
            loss = torch.nn.functional.cross_entropy(logits * 56.3477, label)   # CONCH's logit scale (model_adapters.py:190)
            loss.backward()
            out["label_%d" % i], out["loss_%d" % i] = int(label), float(loss.detach())
            for k, p_ in m.named_parameters():
                out["grad_%d_%s" % (i, k)] = p_.grad.numpy().copy()
        np.savez_compressed(os.path.join(OUT, "heads_clip_ada_bwd_c%d.npz" % c), **out)
        print("wrote heads_clip_ada_bwd_c%d" % c)

    torch.manual_seed(51)
    m = mil.MIL_fc(size_arg="benchmark", dropout=False, n_classes=2, top_k=1).train()
    fp16_exact(m)
    stored = np.load(os.path.join(OUT, "heads_mil_fc.npz"))
    out = {}
    i = 1
    x = torch.from_numpy(stored["feat_%d" % i]).float()
    top, _, _, _, _ = m(x)
    assert np.array_equal(top.detach().numpy(), stored["top_instance_%d" % i])
    loss = torch.nn.functional.cross_entropy(top, torch.tensor([1]))
    loss.backward()
    out["bag"], out["label"], out["loss"] = i, 1, float(loss.detach())
    for k, p_ in m.named_parameters():
        out["grad_" + k] = p_.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "heads_mil_fc_bwd.npz"), **out)
    print("wrote heads_mil_fc_bwd")


if __name__ == "__main__":
    if "--backward-sparse" in sys.argv:
        main_backward_sparse()
    elif "--backward" in sys.argv:
        main_backward()
    else:
        main()
        main_backward()
        main_backward_sparse()
