"""The reference-shaped Python API (slide_process, selectors, poolers, senet, loops) on the GPU against the
golden fixtures written by the reference's own code."""
import types

import numpy as np
import pytest
import torch

from oracle import moc_oracle as O
from tests.helpers import assert_topj_set, assert_union_set, close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _args(c, j, k, discard=()):
    return types.SimpleNamespace(disable_tqdm=True, n_classes=c, topj=j, topk=k, discard_classifiers=list(discard),
                                 pretrain="conch", ablation_study="none", cache_scores=False)


@pytest.mark.parametrize("name", ["slide_c2", "slide_c2_j64", "slide_c3", "slide_c30"])
def test_slide_process_golden(golden, name):
    from moc_b200 import senet, slide_process, topj_pooling
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]).to(DEV), T(g["W_ext"]).to(DEV)
    model = senet(512, 4)
    model.load_state_dict({kk[3:].replace("model_0_", "model.0.").replace("model_2_", "model.2."): T(v)
                           for kk, v in g.items() if kk.startswith("sd_")})
    model.to(DEV).eval()
    for i in range(int(g["n_slides"])):
        x = T(g["s%d_feat" % i]).float()  # host tensor: slide_process moves it, like the reference
        ok = O.selection_keys(x, w.cpu(), we.cpu(), c)
        okeys = np.concatenate([ok["logit"].T, ok["softmax"].T, ok["delta"][None], ok["bg_sum"][None], ok["bg_max"][None]], 0)
        for di, disc in enumerate(g["discards"]):
            disc = [d for d in str(disc).split("|") if d]
            q = "s%d_d%d_" % (i, di)
            r = slide_process(x, w, we, n_classes=c, topj=j, discard_classifiers=disc)
            assert isinstance(r["selected_index"], list)
            ref_idx = g[q + "selected_index"].tolist()
            common = assert_union_set(r["selected_index"], ref_idx, okeys, c, j, disc)
            same = r["selected_index"] == ref_idx
            gp = [r["selected_index"].index(v) for v in common]        # all rows unless a rank-J tie was swapped
            rp = [ref_idx.index(v) for v in common]
            assert torch.equal(r["selected_feat"].cpu(), x[r["selected_index"]])
            close(r["logits_top_classifier"][gp], g[q + "plane_top"][rp])
            close(r["logits_delta_softmax_classifier"][gp], g[q + "plane_dsoftmax"][rp])
            close(r["logits_delta_diff_classifier"][gp], g[q + "plane_ddiff"][rp])
            close(r["logits_bottomk_irrel_classifier"][gp], g[q + "plane_bottomk"][rp])
            with torch.no_grad():
                gate = model(r["selected_feat"])
            close(gate[gp], g[q + "gate"][rp], rtol=1e-4, atol=1e-6)
            # the reference's own eval-time combination written with torch ops on our tensors
            f = gate[:, 0:1] * r["logits_top_classifier"]
            if "delta_softmax" not in disc:
                f = f + gate[:, 1:2] * r["logits_delta_softmax_classifier"]
            if "delta_diff" not in disc:
                f = f + gate[:, 2:3] * r["logits_delta_diff_classifier"]
            f = f + gate[:, 3:4] * r["logits_bottomk_irrel_classifier"]
            pooled = topj_pooling(f, [k])[1][k]
            close(pooled, f.double().topk(min(k, f.size(0)), dim=0).values.mean(dim=0, keepdim=True))
            if same:   # a swapped rank-J tie changes which rows are pooled: the golden holds for the identical selection
                close(pooled, g[q + "bag_logits"])


@pytest.mark.parametrize("name", ["slide_c2", "slide_c3", "slide_c30"])
def test_selectors_and_poolers_golden(golden, name):
    import moc_b200 as M
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    for i in range(int(g["n_slides"])):
        lo, le = T(g["s%d_L" % i]), T(g["s%d_Le" % i])
        lod, led = lo.to(DEV), le.to(DEV)
        n = lo.size(0)
        sm = torch.softmax(lo, dim=1).numpy()
        srt = np.sort(lo.numpy().astype(np.float64), axis=1)
        delta = np.abs(srt[:, -1] - srt[:, -2])
        bg = le[:, c:].sum(dim=1).numpy()
        i1 = M.index_topj_classifier(lod, [j]).cpu().numpy()
        i2 = M.index_delta_softmax_classifier(lod, [j]).cpu().numpy()
        i3 = M.index_delta_diff_classifier(lod, [j]).cpu().numpy()
        i4 = M.index_bottomk_irrel_classifier(led, [j], c).cpu().numpy()
        for arr, key in ((i1, "idx_topj"), (i2, "idx_dsoftmax"), (i3, "idx_ddiff"), (i4, "idx_bottomk")):
            assert arr.shape == g["s%d_%s" % (i, key)].shape and arr.dtype == np.int64
        for cc in range(c):
            assert_topj_set(i1[:, cc], g["s%d_idx_topj" % i][:, cc], lo[:, cc].numpy(), j)
            assert_topj_set(i2[:, cc], g["s%d_idx_dsoftmax" % i][:, cc], sm[:, cc], j)
            assert_topj_set(i3[:, cc], g["s%d_idx_ddiff" % i][:, cc], delta, j)
            assert_topj_set(i4[:, cc], g["s%d_idx_bottomk" % i][:, cc], bg, j, largest=False)
        for fn, key, arg in ((M.topj_pooling, "pool_topj", lod), (M.delta_softmax_classifier_pooling, "pool_dsoftmax", lod),
                             (M.delta_diff_classifier_pooling, "pool_ddiff", lod)):
            preds, pooled = fn(arg, [k])
            close(pooled[k], g["s%d_%s" % (i, key)])
            assert int(preds[k]) == int(np.argmax(g["s%d_%s" % (i, key)]))
        preds, pooled, idx = M.bottomk_irrel_classifier_pooling(led, [k], return_indices=True, coords_list=c)
        close(pooled[k], g["s%d_pool_bottomk" % i])
        assert idx.shape == (min(k, n), c)
        # several pooling sizes in one call, as the reference's dict-of-j API allows
        preds, pooled = M.topj_pooling(lod, [1, 5, 50])
        for jj in (1, 5, 50):
            ref = lo.topk(min(50, n), 0)[0][:min(jj, n)].mean(dim=0, keepdim=True)
            close(pooled[jj], ref)


def test_senet_autograd_matches_torch():
    """senet.forward through the CUDA kernels is differentiable w.r.t. its parameters like the nn.Sequential."""
    from moc_b200 import senet
    torch.manual_seed(0)
    m = senet(512, 4).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(512, 64), torch.nn.ReLU(), torch.nn.Linear(64, 4), torch.nn.Sigmoid())
    ref.load_state_dict({k.replace("model.", ""): v.cpu() for k, v in m.state_dict().items()})
    assert sorted(m.state_dict().keys()) == ["model.0.bias", "model.0.weight", "model.2.bias", "model.2.weight"]
    x = torch.randn(300, 512) / 22.6
    coef = torch.randn(300, 4)
    coef[torch.rand(300) > 0.1] = 0  # sparse upstream gradient, like top-K pooling produces
    out = m(x.to(DEV))
    (out * coef.to(DEV)).sum().backward()
    out_ref = ref(x)
    (out_ref * coef).sum().backward()
    close(out, out_ref, rtol=1e-5, atol=1e-6)
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), ref.named_parameters()):
        scale = float(p2.grad.abs().max())
        assert float((p1.grad.cpu() - p2.grad).abs().max()) <= 1e-4 * scale + 1e-9, n1


@pytest.mark.parametrize("name", ["loop_c2", "loop_c3_discard"])
def test_loops_golden(golden, name):
    """zs_evaluation / ablation_evaluation / train / evaluation against the reference's own loop outputs."""
    import moc_b200 as M
    from moc_b200 import loops
    g = golden(name)
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    disc = [d for d in str(g["discard"]).split("|") if d]
    loops.set_prompts(T(g["W"]).to(DEV), T(g["W_ext"]).to(DEV))
    args = _args(c, j, k, disc)
    tr_store = M.RaggedBagStore.from_bags([T(g["train_feat_%d" % i]).float() for i in range(int(g["n_train"]))],
                                          g["train_labels"].tolist(), DEV)
    va_store = M.RaggedBagStore.from_bags([T(g["val_feat_%d" % i]).float() for i in range(int(g["n_val"]))],
                                          g["val_labels"].tolist(), DEV)
    tr = M.BagLoader(M.BagDataset(tr_store, repeat_num=int(g["repeat_num"])))
    va = M.BagLoader(M.BagDataset(va_store))

    def ev(d):
        return np.asarray([d["loss"], d["acc"], d["auc"]])

    def same(d, key, rtol=2e-5):
        got, ref = ev(d), g[key]
        np.testing.assert_allclose(got[0], ref[0], rtol=rtol, err_msg=key + " loss")
        assert got[1] == ref[1], key + " acc"
        assert got[2] == ref[2], key + " auc"

    same(M.zs_evaluation(tr, DEV, args), "zs_train")
    same(M.zs_evaluation(va, DEV, args), "zs_val")
    same(M.zs_evaluation(va, DEV, args, pooling_func=M.delta_softmax_classifier_pooling), "zs_val_dsoftmax")
    same(M.zs_evaluation(va, DEV, args, pooling_func=M.delta_diff_classifier_pooling), "zs_val_ddiff")
    same(M.zs_evaluation(va, DEV, args, pooling_func=M.bottomk_irrel_classifier_pooling), "zs_val_bottomk")
    for how in ("avg", "sum", "max"):
        args.ablation_study = how
        same(M.ablation_evaluation(va, DEV, args), "ablation_val_" + how)
    args.ablation_study = "none"
    # reference quirk kept: ablation_evaluation/evaluation restore repeat_num = len(dataset), so None becomes n
    assert tr.dataset.repeat_num == int(g["repeat_num"]) and va.dataset.repeat_num == int(g["n_val"])

    model = M.senet(512, 4)
    model.load_state_dict({kk[4:].replace("model_0_", "model.0.").replace("model_2_", "model.2."): T(v)
                           for kk, v in g.items() if kk.startswith("sd0_")})
    model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    masks = [T(g["mask_%d" % i]) for i in range(int(g["n_masks"]))]
    per = int(g["repeat_num"])
    for e in range(int(g["epochs"])):
        losses = M.train(model, tr, opt, DEV, args, masks=masks[e * per:(e + 1) * per])
        np.testing.assert_allclose(losses.cpu().numpy(), g["train_losses_e%d" % e], rtol=5e-5)
        for key, t in model.state_dict().items():
            ref = g["sd_e%d_" % e + key.replace(".", "_")]
            assert np.abs(t.cpu().numpy() - ref).max() < 2e-5, key  # lr = 1e-3: well under one Adam step
        same(M.evaluation(model, tr, DEV, args), "eval_train_e%d" % e, rtol=5e-5)
        same(M.evaluation(model, va, DEV, args), "eval_val_e%d" % e, rtol=5e-5)
    # the torch optimizer's own state was advanced by our kernel and stays a valid torch state
    p0 = model.model[0].weight
    assert int(opt.state[p0]["step"]) == int(g["adam_step"])
    np.testing.assert_allclose(opt.state[p0]["exp_avg"].cpu().numpy(), g["adam_m_0"], rtol=1e-3, atol=1e-8)
    opt.state_dict()


def test_cached_scores_are_bit_identical(golden):
    import moc_b200 as M
    from moc_b200 import loops
    g = golden("loop_c2")
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    loops.set_prompts(T(g["W"]).to(DEV), T(g["W_ext"]).to(DEV))
    store = M.RaggedBagStore.from_bags([T(g["val_feat_%d" % i]).float() for i in range(int(g["n_val"]))],
                                       g["val_labels"].tolist(), DEV)
    torch.manual_seed(1)
    model = M.senet(512, 4).to(DEV)
    a = M.MocEngine(loops.zeroshot_weights, loops.zeroshot_weights_ext, j, k, cache_scores=False)
    b = M.MocEngine(loops.zeroshot_weights, loops.zeroshot_weights_ext, j, k, cache_scores=True)
    la = a.eval_logits(store, model.head_params())
    lb1 = b.eval_logits(store, model.head_params())
    lb2 = b.eval_logits(store, model.head_params())
    assert torch.equal(la, lb1) and torch.equal(la, lb2)
    # waves: force several small waves and compare with the single-wave result
    s = M.MocEngine(loops.zeroshot_weights, loops.zeroshot_weights_ext, j, k, max_wave_rows=300)
    assert torch.equal(la, s.eval_logits(store, model.head_params()))


def test_score_cta_calibration_changes_nothing_but_speed(monkeypatch):
    """MocEngine times a few CTA counts for the streaming kernel on its first large call and keeps the fastest for the
    device; keys are bit-identical for every count, with the calibration on or off."""
    import moc_b200
    from moc_b200 import ops, synthetic
    from moc_b200.engine import MocEngine
    c = 2
    w, we = synthetic.prompt_matrices(c, device=DEV)
    store = moc_b200.RaggedBagStore.synthetic([3000] * 40, c, we, cohort_seed=5, device=DEV)
    ref = ops.score_keys(store.feat, ops.Prompts.pack(w, we))
    for cap in (1, 7, 100, 124, 132, 148, 1000):
        assert torch.equal(ops.score_keys(store.feat, ops.Prompts.pack(w, we), max_ctas=cap), ref), cap
    monkeypatch.setattr(MocEngine, "_TUNE_MIN_ROWS", 100_000)
    monkeypatch.setattr(MocEngine, "_TUNED_CTAS", {})
    eng = MocEngine(w, we, 400, 10)
    assert torch.equal(eng.keys_for(store), ref)
    dev = torch.cuda.current_device()
    if torch.cuda.get_device_properties(dev).multi_processor_count == 148:
        assert MocEngine._TUNED_CTAS.get(dev) in MocEngine._TUNE_CANDIDATES
    monkeypatch.setenv("MOC_SCORE_AUTOTUNE", "0")
    monkeypatch.setattr(MocEngine, "_TUNED_CTAS", {})
    assert torch.equal(eng.keys_for(store), ref) and MocEngine._TUNED_CTAS == {}


@pytest.mark.parametrize("c", [8, 9, 12, 30])
def test_zero_shot_ablation_and_dict_api_across_the_key_layout_change(c):
    """Class counts on both sides of MOC_KEYS_COMPACT_MIN_CLASSES (2C+3 key planes up to 8 classes, C+4 from 9 on: a
    log-sum-exp plane instead of the softmax planes): the four zero-shot poolings, the three ablation combinations, the
    evaluation pass and slide_process' dense planes against the oracle."""
    import moc_b200 as M
    from moc_b200 import ops, synthetic
    from moc_b200.engine import MocEngine
    j, k = 60, 7
    w, we = synthetic.prompt_matrices(c)
    sizes = [900, 1501, 333, 2048]
    bags = [synthetic.make_bag(n, i % c, we, c, seed=700 + 10 * c + i) for i, n in enumerate(sizes)]
    labels = [i % c for i in range(len(sizes))]
    store = M.RaggedBagStore.from_bags(bags, labels, DEV)
    assert ops.num_key_planes(c) == (2 * c + 3 if c < 9 else c + 4)
    eng = MocEngine(w.to(DEV), we.to(DEV), j, k)

    def zs_ref(x, pooling):     # the per-slide body of zs_evaluation (main_moc.py:427-432)
        lo, le = O.score(x, w, we)
        if pooling == "topj":
            return O.topj_pooling(lo, [k])[1][k]
        if pooling == "delta_softmax":
            return O.delta_softmax_pooling(lo, [k])[1][k]
        if pooling == "delta_diff":
            return O.delta_diff_pooling(lo, [k])[1][k]
        return O.bottomk_irrel_pooling(le, [k], coords_list=c)[1][k]

    for pooling in ("topj", "delta_softmax", "delta_diff", "bottomk_irrel"):
        ref = torch.cat([zs_ref(x, pooling) for x in bags], 0)
        close(eng.zero_shot_logits(store, pooling), ref, rtol=1e-3, atol=2e-6)
    for how in ("avg", "sum", "max"):
        ref = torch.cat([O.ablation_logits(x, w, we, c, j, k, how) for x in bags], 0)
        close(eng.ablation_logits(store, how), ref, rtol=1e-3, atol=2e-6)
    oprm = O.SenetParams.init(11)
    prm = ops.HeadParams(oprm.w1.to(DEV), oprm.b1.to(DEV), oprm.w2.to(DEV), oprm.b2.to(DEV))
    ref = torch.cat([O.slide_eval_logits(oprm, x, w, we, c, j, k) for x in bags], 0)
    close(eng.eval_logits(store, prm, check_domain=True), ref, rtol=1e-3, atol=2e-6)
    # the dict API: dense planes of the selected rows, as the reference returns them
    loops_w, loops_we = w.to(DEV), we.to(DEV)
    got = M.slide_process(bags[1].to(DEV), loops_w, loops_we, c, j)
    want = O.slide_process(bags[1], w, we, c, j)
    if got["selected_index"] == want["selected_index"]:
        for name in ("logits_top_classifier", "logits_delta_softmax_classifier", "logits_delta_diff_classifier",
                     "logits_bottomk_irrel_classifier"):
            close(got[name], want[name], rtol=1e-3, atol=2e-6)


def test_ext_class_columns_differ_from_w(golden):
    """zs_evaluation(pooling_func=bottomk_irrel_classifier_pooling) pools (feats @ W_ext)[:, :C] (main_moc.py:428-432):
    with a W_ext whose class columns are not W the engine must score those columns, not reuse the W planes.  Golden
    from the reference's own loops (oracle/make_golden_extfg.py)."""
    import moc_b200 as M
    from moc_b200 import loops
    from moc_b200.engine import MocEngine
    g = golden("zs_extfg")
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]).to(DEV), T(g["W_ext"]).to(DEV)
    loops.set_prompts(w, we)
    args = _args(c, j, k)
    store = M.RaggedBagStore.from_bags([T(g["feat_%d" % i]).float() for i in range(int(g["n_slides"]))],
                                       g["labels"].tolist(), DEV)
    ld = M.BagLoader(M.BagDataset(store))

    def same(d, key, rtol=2e-5):
        ref = g[key]
        np.testing.assert_allclose(d["loss"], ref[0], rtol=rtol, err_msg=key)
        assert d["acc"] == ref[1] and d["auc"] == ref[2], key

    same(M.zs_evaluation(ld, DEV, args), "zs_topj")
    same(M.zs_evaluation(ld, DEV, args, pooling_func=M.delta_softmax_classifier_pooling), "zs_dsoftmax")
    same(M.zs_evaluation(ld, DEV, args, pooling_func=M.delta_diff_classifier_pooling), "zs_ddiff")
    same(M.zs_evaluation(ld, DEV, args, pooling_func=M.bottomk_irrel_classifier_pooling), "zs_bottomk")
    eng = MocEngine(w, we, j, k)
    close(eng.zero_shot_logits(store, "bottomk_irrel"), g["bottomk_logits"])
    assert not np.allclose(eng.zero_shot_logits(store, "topj").cpu().numpy(), g["bottomk_logits"], atol=1e-4)
    model = M.senet(512, 4)
    model.load_state_dict({kk[3:].replace("model_0_", "model.0.").replace("model_2_", "model.2."): T(v)
                           for kk, v in g.items() if kk.startswith("sd_")})
    same(M.evaluation(model.to(DEV), ld, DEV, args), "eval")


@pytest.mark.parametrize("g_size", [2, 4, 6])
def test_dp_microbatch_training_matches_oracle_variant(golden, g_size):
    """train(..., dp_microbatch=G): gradients of G consecutive slides at common parameters, summed, one Adam step -
    against oracle.train_epoch(dp_microbatch=G) on the same bags and masks (single rank: pure accumulation; the
    multi-rank split of the same sum is covered by tests/dist_eval_worker.py)."""
    import moc_b200 as M
    from moc_b200 import loops
    g = golden("loop_c2")
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    w, we = T(g["W"]), T(g["W_ext"])
    bags = [T(g["train_feat_%d" % i]).float() for i in range(int(g["n_train"]))]
    labels = g["train_labels"].tolist()
    rep = int(g["repeat_num"])
    masks = [T(g["mask_%d" % i]) for i in range(rep)]
    oprm = O.SenetParams(*[T(g["sd0_" + n]).clone() for n in ("model_0_weight", "model_0_bias", "model_2_weight", "model_2_bias")])
    st = O.AdamState()
    ref_losses = O.train_epoch(oprm, st, O.BagList(bags, labels, repeat_num=rep), w, we, c, j, k, (), masks,
                               dp_microbatch=g_size)
    assert st.step == -(-rep // g_size)

    loops.set_prompts(w.to(DEV), we.to(DEV))
    model = M.senet(512, 4)
    model.load_state_dict({kk[4:].replace("model_0_", "model.0.").replace("model_2_", "model.2."): T(v)
                           for kk, v in g.items() if kk.startswith("sd0_")})
    model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    tr = M.BagLoader(M.BagDataset(M.RaggedBagStore.from_bags(bags, labels, DEV), repeat_num=rep))
    losses = M.train(model, tr, opt, DEV, _args(c, j, k), masks=masks, dp_microbatch=g_size)
    np.testing.assert_allclose(losses.cpu().numpy(), np.asarray(ref_losses), rtol=5e-5)
    for got, ref in zip(model.parameters_in_order(), oprm.tensors()):
        assert float((got.detach().cpu() - ref).abs().max()) < 2e-5
    # and it is a different trajectory from the reference's one-step-per-slide loop
    one = O.SenetParams(*[T(g["sd0_" + n]).clone() for n in ("model_0_weight", "model_0_bias", "model_2_weight", "model_2_bias")])
    O.train_epoch(one, O.AdamState(), O.BagList(bags, labels, repeat_num=rep), w, we, c, j, k, (), masks)
    assert float((one.w2 - oprm.w2).abs().max()) > 1e-5


def test_graph_captured_training_is_bit_identical_to_eager(golden):
    """train() replays one CUDA graph per few-shot slide (forward + CE + backward + Adam with a device step counter);
    the eager sequence of the same kernels must give bit-identical losses, parameters and optimizer state, also across
    epochs, and the optimizer's host-side step count must follow."""
    import copy
    import moc_b200 as M
    from moc_b200 import loops
    g = golden("loop_c2")
    c, j, k = int(g["C"]), int(g["J"]), int(g["K"])
    loops.set_prompts(T(g["W"]).to(DEV), T(g["W_ext"]).to(DEV))
    bags = [T(g["train_feat_%d" % i]).float() for i in range(int(g["n_train"]))]
    rep = int(g["repeat_num"])
    tr = M.BagLoader(M.BagDataset(M.RaggedBagStore.from_bags(bags, g["train_labels"].tolist(), DEV), repeat_num=rep))
    masks = [T(g["mask_%d" % i]) for i in range(int(g["n_masks"]))]
    model_g = M.senet(512, 4)
    model_g.load_state_dict({kk[4:].replace("model_0_", "model.0.").replace("model_2_", "model.2."): T(v)
                             for kk, v in g.items() if kk.startswith("sd0_")})
    model_g.to(DEV)
    model_e = copy.deepcopy(model_g)
    opt_g = torch.optim.Adam(model_g.parameters(), lr=1e-3, weight_decay=1e-4)
    opt_e = torch.optim.Adam(model_e.parameters(), lr=1e-3, weight_decay=1e-4)
    a_g, a_e = _args(c, j, k), _args(c, j, k)
    a_e.cuda_graph = False
    for e in range(int(g["epochs"])):
        mk = masks[e * rep:(e + 1) * rep]
        lg = M.train(model_g, tr, opt_g, DEV, a_g, masks=mk)
        le = M.train(model_e, tr, opt_e, DEV, a_e, masks=mk)
        assert torch.equal(lg, le)
        np.testing.assert_allclose(lg.cpu().numpy(), g["train_losses_e%d" % e], rtol=5e-5)
        for pg, pe in zip(model_g.parameters(), model_e.parameters()):
            assert torch.equal(pg, pe)
            assert torch.equal(opt_g.state[pg]["exp_avg"], opt_e.state[pe]["exp_avg"])
            assert torch.equal(opt_g.state[pg]["exp_avg_sq"], opt_e.state[pe]["exp_avg_sq"])
            assert int(opt_g.state[pg]["step"]) == int(opt_e.state[pe]["step"]) == (e + 1) * rep
    # an eager step in between (someone else stepping the optimizer) must not desynchronise the device counter
    a_g.cuda_graph = False
    M.train(model_g, tr, opt_g, DEV, a_g, masks=masks[:rep])
    M.train(model_e, tr, opt_e, DEV, a_e, masks=masks[:rep])
    a_g.cuda_graph = True
    lg = M.train(model_g, tr, opt_g, DEV, a_g, masks=masks[:rep])
    le = M.train(model_e, tr, opt_e, DEV, a_e, masks=masks[:rep])
    assert torch.equal(lg, le) and all(torch.equal(pg, pe) for pg, pe in zip(model_g.parameters(), model_e.parameters()))


def test_seeded_run_draws_the_reference_masks():
    """A SEEDED run without explicit masks: the reference's loops consume the CPU generator once per
    ``iter(DataLoader)`` (base seed) and once per training slide (``torch.rand(N) > 0.5``, main_moc.py:330).  Our loops
    mirror both, so after the same ``torch.manual_seed`` a sequence train / evaluation / zs_evaluation / train gives the
    parameters the reference's own functions give over a real torch DataLoader (lifted from oracle/_ref)."""
    import types
    import moc_b200 as M
    from moc_b200 import loops, synthetic
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("the staged reference files (oracle/_ref) are not present")
    c, j, k = 2, 64, 10
    w, we = synthetic.prompt_matrices(c)
    bags, labels = synthetic.make_cohort(4, [300, 350, 280, 320], c, cohort_seed=77)
    rep = 6

    class Split(torch.utils.data.Dataset):           # what Generic_Split looks like to the reference's loops
        def __init__(self, repeat_num):
            self.repeat_num = repeat_num

        def real_len(self):
            return len(bags)

        def __len__(self):
            return self.repeat_num if self.repeat_num else len(bags)

        def __getitem__(self, idx):
            if idx >= len(self):
                raise IndexError
            i = idx % len(bags)
            return bags[i], labels[i], np.zeros((bags[i].size(0), 2), dtype=np.int64), "slide_%d" % i

    args = types.SimpleNamespace(disable_tqdm=True, n_classes=c, topj=j, topk=k, discard_classifiers=[], pretrain="conch",
                                 ablation_study="none", cache_scores=False)
    ref = ref_loader.load()
    ref.set_weights(w, we)
    torch.manual_seed(21)
    model_ref = ref.senet(512, 4)
    init = {kk: v.clone() for kk, v in model_ref.state_dict().items()}
    opt_ref = torch.optim.Adam(model_ref.parameters(), lr=1e-3, weight_decay=1e-4)
    for workers in (0, 1):                             # the iterator's draw is the same with and without a worker
        loader = torch.utils.data.DataLoader(Split(rep), batch_size=1, shuffle=False, num_workers=workers)
        model_ref.load_state_dict(init)
        opt_ref = torch.optim.Adam(model_ref.parameters(), lr=1e-3, weight_decay=1e-4)
        torch.manual_seed(22)
        ref.train(model_ref, loader, opt_ref, "cpu", args)
        ev_ref = ref.evaluation(model_ref, loader, "cpu", args)
        ref.zs_evaluation(loader, "cpu", args)
        ref.train(model_ref, loader, opt_ref, "cpu", args)
        want = {kk: v.clone() for kk, v in model_ref.state_dict().items()}

        loops.set_prompts(w.to(DEV), we.to(DEV))
        model = M.senet(512, 4)
        model.load_state_dict(init)
        model.to(DEV)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
        ours = M.BagLoader(M.BagDataset(M.RaggedBagStore.from_bags(bags, labels, DEV), repeat_num=rep))
        torch.manual_seed(22)
        M.train(model, ours, opt, DEV, args)
        ev = M.evaluation(model, ours, DEV, args)
        M.zs_evaluation(ours, DEV, args)
        M.train(model, ours, opt, DEV, args)
        assert ev["acc"] == ev_ref["acc"] and ev["auc"] == ev_ref["auc"] and abs(ev["loss"] - ev_ref["loss"]) < 5e-5
        for kk, v in model.state_dict().items():
            assert float((v.cpu() - want[kk]).abs().max()) < 2e-5, (workers, kk)
