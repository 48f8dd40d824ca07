"""The oracle restatement replayed against fixtures produced by the reference's own code
(oracle/make_golden.py).  Same ATen build underneath => forward quantities must agree bit-for-bit;
the hand-written backward/Adam is compared at 1e-6."""
import numpy as np
import pytest
import torch

from oracle import moc_oracle as O

SLIDE_CASES = ["slide_c2", "slide_c2_j64", "slide_c3", "slide_c30"]


def T(a):
    return torch.from_numpy(np.asarray(a))


def feat_of(g, key):
    return T(g[key]).float()


def params_of(g, prefix):
    return O.SenetParams(T(g[prefix + "model_0_weight"]).clone(), T(g[prefix + "model_0_bias"]).clone(),
                         T(g[prefix + "model_2_weight"]).clone(), T(g[prefix + "model_2_bias"]).clone())


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", SLIDE_CASES)
def test_scores_selectors_poolers(golden, name):
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]), T(g["W_ext"])
    for i in range(int(g["n_slides"])):
        p = "s%d_" % i
        x = feat_of(g, p + "feat")
        lo, le = O.score(x, w, we)
        assert torch.equal(lo, T(g[p + "L"])) and torch.equal(le, T(g[p + "Le"]))
        assert torch.equal(O.index_topj(lo, [j]), T(g[p + "idx_topj"]))
        assert torch.equal(O.index_delta_softmax(lo, [j]), T(g[p + "idx_dsoftmax"]))
        assert torch.equal(O.index_delta_diff(lo, [j]), T(g[p + "idx_ddiff"]))
        assert torch.equal(O.index_bottomk_irrel(le, [j], c), T(g[p + "idx_bottomk"]))
        assert torch.equal(O.topj_pooling(lo, [k])[1][k], T(g[p + "pool_topj"]))
        assert torch.equal(O.delta_softmax_pooling(lo, [k])[1][k], T(g[p + "pool_dsoftmax"]))
        assert torch.equal(O.delta_diff_pooling(lo, [k])[1][k], T(g[p + "pool_ddiff"]))
        assert torch.equal(O.bottomk_irrel_pooling(le, [k], coords_list=c)[1][k], T(g[p + "pool_bottomk"]))


@pytest.mark.parametrize("name", SLIDE_CASES)
def test_slide_process_gate_and_bag_logits(golden, name):
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]), T(g["W_ext"])
    prm = params_of(g, "sd_")
    for i in range(int(g["n_slides"])):
        x = feat_of(g, "s%d_feat" % i)
        for di, disc in enumerate(g["discards"]):
            disc = tuple(d for d in str(disc).split("|") if d)
            q = "s%d_d%d_" % (i, di)
            r = O.slide_process(x, w, we, c, j, discard_classifiers=disc)
            assert r["selected_index"] == g[q + "selected_index"].tolist()
            assert torch.equal(r["logits_top_classifier"], T(g[q + "plane_top"]))
            assert torch.equal(r["logits_delta_softmax_classifier"], T(g[q + "plane_dsoftmax"]))
            assert torch.equal(r["logits_delta_diff_classifier"], T(g[q + "plane_ddiff"]))
            assert torch.equal(r["logits_bottomk_irrel_classifier"], T(g[q + "plane_bottomk"]))
            gate, _ = O.senet_forward(prm, r["selected_feat"])
            torch.testing.assert_close(gate, T(g[q + "gate"]), rtol=0, atol=1e-7)
            f = O.combine(gate, r, O.active_classifiers(disc, "eval"))
            torch.testing.assert_close(f, T(g[q + "final"]), rtol=1e-6, atol=1e-7)
            torch.testing.assert_close(O.bag_logits(f, k), T(g[q + "bag_logits"]), rtol=1e-6, atol=1e-7)


def test_selection_set_identities(golden):
    """Facts the CUDA design leans on: delta-diff columns are identical; the bottom-k index set is the
    bottom-maxj of the background sum."""
    for name in SLIDE_CASES:
        g = golden(name)
        c, j = int(g["C"]), int(g["J"])
        for i in range(int(g["n_slides"])):
            p = "s%d_" % i
            dd = g[p + "idx_ddiff"]
            assert all(set(dd[:, 0]) == set(dd[:, cc]) for cc in range(dd.shape[1]))
            bg = g[p + "Le"][:, c:].sum(axis=1)
            maxj = min(j, bg.shape[0])
            thr = O.rank_threshold(bg, maxj, largest=False)
            got = set(g[p + "idx_bottomk"].flatten().tolist())
            assert len(got) == maxj
            assert all(bg[r] <= thr for r in got)
            assert set(np.nonzero(bg < thr)[0].tolist()) <= got


@pytest.mark.parametrize("name", ["loop_c2", "loop_c3_discard"])
def test_train_eval_loops(golden, name):
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    disc = tuple(d for d in str(g["discard"]).split("|") if d)
    w, we = T(g["W"]), T(g["W_ext"])
    tr = O.BagList([feat_of(g, "train_feat_%d" % i) for i in range(int(g["n_train"]))],
                   g["train_labels"].tolist(), repeat_num=int(g["repeat_num"]))
    va = O.BagList([feat_of(g, "val_feat_%d" % i) for i in range(int(g["n_val"]))], g["val_labels"].tolist())

    def ev(d):
        return np.asarray([d["loss"], d["acc"], d["auc"]])

    np.testing.assert_allclose(ev(O.zs_evaluation(tr, w, we, c, k)), g["zs_train"], rtol=1e-6)
    np.testing.assert_allclose(ev(O.zs_evaluation(va, w, we, c, k)), g["zs_val"], rtol=1e-6)
    np.testing.assert_allclose(ev(O.zs_evaluation(va, w, we, c, k, "delta_softmax")), g["zs_val_dsoftmax"], rtol=1e-6)
    np.testing.assert_allclose(ev(O.zs_evaluation(va, w, we, c, k, "delta_diff")), g["zs_val_ddiff"], rtol=1e-6)
    np.testing.assert_allclose(ev(O.zs_evaluation(va, w, we, c, k, "bottomk_irrel")), g["zs_val_bottomk"], rtol=1e-6)
    for how in ("avg", "sum", "max"):
        rows = [O.ablation_logits(x, w, we, c, j, k, how) for x in va.bags]
        loss = sum(float(O.cross_entropy(r, y)) for r, y in zip(rows, va.labels))
        m = O._metrics(torch.cat(rows, 0), va.labels, loss, len(va), va.real_len())
        np.testing.assert_allclose(ev(m), g["ablation_val_" + how], rtol=1e-6)

    prm = params_of(g, "sd0_")
    st = O.AdamState()
    masks = [T(g["mask_%d" % i]) for i in range(int(g["n_masks"]))]
    per = int(g["repeat_num"])
    for e in range(int(g["epochs"])):
        losses = O.train_epoch(prm, st, tr, w, we, c, j, k, disc, masks[e * per:(e + 1) * per])
        np.testing.assert_allclose(losses, g["train_losses_e%d" % e], rtol=2e-6)
        for key, t in prm.state_dict().items():
            np.testing.assert_allclose(t.numpy(), g["sd_e%d_" % e + key.replace(".", "_")], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ev(O.evaluation(prm, tr, w, we, c, j, k, disc)), g["eval_train_e%d" % e], rtol=2e-6)
        np.testing.assert_allclose(ev(O.evaluation(prm, va, w, we, c, j, k, disc)), g["eval_val_e%d" % e], rtol=2e-6)
    assert st.step == int(g["adam_step"])
    for gi in range(4):
        np.testing.assert_allclose(st.m[gi].numpy(), g["adam_m_%d" % gi], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(st.v[gi].numpy(), g["adam_v_%d" % gi], rtol=1e-4, atol=1e-12)


def test_senet_init_matches_linear_default():
    """SenetParams.init reproduces nn.Linear's default distribution bounds (not the stream)."""
    p = O.SenetParams.init(0)
    assert p.w1.shape == (64, 512) and p.b1.shape == (64,) and p.w2.shape == (4, 64) and p.b2.shape == (4,)
    assert float(p.w1.abs().max()) <= 1 / np.sqrt(512) and float(p.w2.abs().max()) <= 1 / 8
    assert sum(t.numel() for t in p.tensors()) == 33092


@pytest.mark.parametrize("name", ["bank_rcc_ext", "bank_stress"])
def test_prompt_bank_collapse(golden, name):
    """zero_shot_classifier (utils/zeroshot_utils.py:20-51) run on a stand-in text tower vs the oracle's restatement."""
    g = golden(name)
    w = O.collapse_prompt_bank(T(g["bank"]).float(), g["prompts_per_class"].tolist())
    assert torch.equal(w, T(g["W"]))
    assert torch.allclose(w.norm(dim=0), torch.ones(w.size(1)), atol=1e-6)


# ---- secondary MIL heads: oracle restatement vs the reference's own modules (oracle/make_golden_heads.py) ----------
def _sd(g):
    return {k[3:]: T(v).float() for k, v in g.items() if k.startswith("sd_")}


@pytest.mark.parametrize("name", ["heads_clip_ada_c2", "heads_clip_ada_c3"])
def test_heads_clip_ada(golden, name):
    from oracle import moc_oracle_heads as H
    g = golden(name)
    sd, cl = _sd(g), T(g["classifier"])
    for i in range(int(g["n_bags"])):
        x = T(g["feat_%d" % i]).float()
        assert torch.equal(H.clip_ada_forward(sd, cl, x, float(g["clip_ratio"]), int(g["topj"])), T(g["forward_%d" % i]))
        assert torch.equal(H.clip_ada_forward_disable_ada(cl, x, int(g["topj"])), T(g["forward_disable_ada_%d" % i]))


@pytest.mark.parametrize("name", ["heads_abmil_c2", "heads_abmil_c3"])
def test_heads_abmil(golden, name):
    from oracle import moc_oracle_heads as H
    g = golden(name)
    sd = _sd(g)
    for i in range(int(g["n_bags"])):
        logits, y_prob, y_hat, a_raw, pooled = H.abmil_forward(sd, T(g["feat_%d" % i]).float())
        assert torch.equal(logits, T(g["logits_%d" % i])) and torch.equal(y_prob, T(g["y_prob_%d" % i]))
        assert torch.equal(y_hat, T(g["y_hat_%d" % i])) and torch.equal(a_raw, T(g["a_raw_%d" % i]))
        assert torch.equal(pooled, T(g["features_%d" % i])) and torch.equal(a_raw, T(g["attention_only_%d" % i]))


@pytest.mark.parametrize("name", ["heads_abmil_c2", "heads_abmil_c3"])
def test_heads_abmil_backward(golden, name):
    """The oracle's training-step gradients (autograd over the restated forward) against the ones autograd produced
    through the reference's own CLAM_SB module."""
    from oracle import moc_oracle_heads as H
    g, gb = golden(name), golden(name.replace("abmil", "abmil_bwd"))
    sd = _sd(g)
    loss, grads = H.abmil_loss_and_grads(sd, T(g["feat_%d" % int(gb["bag"])]).float(), int(gb["label"]))
    assert abs(float(loss) - float(gb["loss"])) < 1e-6
    for k, gr in grads.items():
        if "grad_" + k in gb:
            ref = T(gb["grad_" + k])
            assert (gr - ref).abs().max() <= 1e-6 * ref.abs().max() + 1e-12, k
        else:
            ref = T(gb["grad5_" + k])
            assert (gr.reshape(-1)[::5] - ref).abs().max() <= 1e-6 * ref.abs().max() + 1e-12, k


@pytest.mark.parametrize("name", ["heads_clip_ada_c2", "heads_clip_ada_c3"])
def test_heads_clip_ada_backward(golden, name):
    from oracle import moc_oracle_heads as H
    g, gb = golden(name), golden(name.replace("clip_ada", "clip_ada_bwd"))
    sd, cl = _sd(g), T(g["classifier"])
    for i in range(int(gb["n_bags"])):
        loss, grads = H.clip_ada_loss_and_grads(sd, cl, T(g["feat_%d" % i]).float(), float(g["clip_ratio"]), int(g["topj"]),
                                                int(gb["label_%d" % i]))
        assert abs(float(loss) - float(gb["loss_%d" % i])) < 1e-6
        for k, gr in grads.items():
            ref = T(gb["grad_%d_%s" % (i, k)])
            assert (gr - ref).abs().max() <= 1e-6 * ref.abs().max() + 1e-12, (i, k)


def test_heads_mil_fc_backward(golden):
    from oracle import moc_oracle_heads as H
    g, gb = golden("heads_mil_fc"), golden("heads_mil_fc_bwd")
    loss, grads = H.mil_fc_loss_and_grads(_sd(g), T(g["feat_%d" % int(gb["bag"])]).float(), int(gb["label"]))
    assert abs(float(loss) - float(gb["loss"])) < 1e-6
    for k, gr in grads.items():
        ref = T(gb["grad_" + k])
        assert (gr - ref).abs().max() <= 1e-6 * ref.abs().max() + 1e-12, k


def test_heads_mil_fc(golden):
    from oracle import moc_oracle_heads as H
    g = golden("heads_mil_fc")
    sd = _sd(g)
    for i in range(int(g["n_bags"])):
        top, y_prob, y_hat, y_probs = H.mil_fc_forward(sd, T(g["feat_%d" % i]).float())
        assert torch.equal(top, T(g["top_instance_%d" % i])) and torch.equal(y_prob, T(g["y_prob_%d" % i]))
        assert torch.equal(y_hat, T(g["y_hat_%d" % i])) and torch.equal(y_probs, T(g["y_probs_%d" % i]))


def test_ext_class_columns_differ_from_w(golden):
    """Prompt files whose W_ext[:, :C] is not W (tests/golden/zs_extfg.npz, the reference's own zs_evaluation /
    evaluation): bottomk_irrel pooling reads (feats @ W_ext)[:, :C], everything else scores classes against W."""
    g = golden("zs_extfg")
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]), T(g["W_ext"])
    assert not torch.equal(we[:, :c], w)
    data = O.BagList([feat_of(g, "feat_%d" % i) for i in range(int(g["n_slides"]))], g["labels"].tolist())
    ev = lambda d: np.asarray([d["loss"], d["acc"], d["auc"]])
    for ours, key in (("topj", "zs_topj"), ("delta_softmax", "zs_dsoftmax"), ("delta_diff", "zs_ddiff"),
                      ("bottomk_irrel", "zs_bottomk")):
        np.testing.assert_allclose(ev(O.zs_evaluation(data, w, we, c, k, ours)), g[key], rtol=1e-6)
    _, lg = O.zs_evaluation(data, w, we, c, k, "bottomk_irrel", return_logits=True)
    np.testing.assert_array_equal(lg.numpy(), g["bottomk_logits"])
    np.testing.assert_allclose(ev(O.evaluation(params_of(g, "sd_"), data, w, we, c, j, k)), g["eval"], rtol=1e-6)


def test_dp_microbatch_variant_of_the_oracle(golden):
    """The checker of the data-parallel training mode: G=1 is the reference's loop; G >= epoch length is one Adam step
    on the sum of all per-slide gradients taken at the initial parameters."""
    g = golden("loop_c2")
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]), T(g["W_ext"])
    rep = int(g["repeat_num"])
    data = O.BagList([feat_of(g, "train_feat_%d" % i) for i in range(int(g["n_train"]))], g["train_labels"].tolist(),
                     repeat_num=rep)
    masks = [T(g["mask_%d" % i]) for i in range(rep)]
    a, b = params_of(g, "sd0_"), params_of(g, "sd0_")
    la = O.train_epoch(a, O.AdamState(), data, w, we, c, j, k, (), masks)
    lb = O.train_epoch(b, O.AdamState(), data, w, we, c, j, k, (), masks, dp_microbatch=1)
    assert la == lb and all(torch.equal(x, y) for x, y in zip(a.tensors(), b.tensors()))
    p0, whole = params_of(g, "sd0_"), params_of(g, "sd0_")
    st = O.AdamState()
    lw = O.train_epoch(whole, st, data, w, we, c, j, k, (), masks, dp_microbatch=100)
    assert st.step == 1
    total, losses = None, []
    for i in range(rep):
        feat, lbl = data[i]
        slide = O.slide_process(feat, w, we, c, j, mask=masks[i])
        loss, _, grads = O.head_forward_backward(p0, slide, lbl, k)
        losses.append(float(loss))
        total = grads if total is None else [x + y for x, y in zip(total, grads)]
    O.adam_step(p0, total, O.AdamState())
    assert lw == losses and all(torch.equal(x, y) for x, y in zip(p0.tensors(), whole.tensors()))


@pytest.mark.parametrize("c,n", [(9, 5000), (30, 20000)])
def test_log_softmax_ranks_a_slide_like_softmax(c, n):
    """The compact key layout of wide class sets (include/moc_b200.h) stores lse = log sum exp(L) instead of the C softmax
    planes; its softmax selection takes the top J of L_c - lse where the reference takes the top J of softmax(L)_c
    (utils/patch_selection_classifier_index.py:28-36).  The two are the same ranking up to fp32 rounding among
    near-equal keys: on the oracle's own scores the two top-J sets may differ only in rows whose softmax value lies
    within the parity tolerance of the rank-J value; and expf(L_c - lse) reproduces the softmax plane within 1e-6."""
    from moc_b200 import synthetic
    from tests.helpers import assert_topj_set
    j = 400
    w, we = synthetic.prompt_matrices(c)
    x = synthetic.make_bag(n, 1, we, c, seed=4242 + c)
    lo, _ = O.score(x, w, we)
    soft = torch.softmax(lo, dim=1)
    m = lo.max(dim=1, keepdim=True).values
    lse = m + torch.log(torch.exp(lo - m).sum(dim=1, keepdim=True))        # what the scoring epilogues store (fp32)
    logsoft = lo - lse
    assert (torch.exp(logsoft) - soft).abs().max().item() < 1e-6
    for col in range(c):
        ref = soft[:, col].topk(j).indices.tolist()
        got = logsoft[:, col].topk(j).indices.tolist()
        assert_topj_set(got, ref, soft[:, col].numpy(), j)

