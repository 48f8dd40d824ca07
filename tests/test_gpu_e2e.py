"""BASELINE.json configs[0]: NSCLC 4-shot MOC train+eval - the CPU oracle end to end against the CUDA path.

Same seeded synthetic cohort (8 train / 49 val / 209 test slides, shapes of splits/nsclc_fewshot/4shots), same
initial gate weights, same half masks; per-epoch loss/acc/auc and the final best_results must agree."""
import types

import numpy as np
import pytest
import torch

from oracle import moc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run_case(c, shot, n_patches, n_val, n_test, topj, topk, epochs, seed):
    import moc_b200 as M
    from moc_b200 import loops, synthetic
    w, we = synthetic.prompt_matrices(c)
    n_train = shot * c
    tr_b, tr_y = synthetic.make_cohort(n_train, n_patches, c, cohort_seed=seed)
    va_b, va_y = synthetic.make_cohort(n_val, n_patches, c, cohort_seed=seed + 1)
    te_b, te_y = synthetic.make_cohort(n_test, n_patches, c, cohort_seed=seed + 2)
    gen = torch.Generator().manual_seed(seed + 3)
    masks = [[torch.rand(tr_b[k % n_train].size(0), generator=gen) > 0.5 for k in range(n_train)] for _ in range(epochs)]

    # ---- oracle (CPU) -----------------------------------------------------------------------------
    torch.set_num_threads(torch.get_num_threads())
    prm = O.SenetParams.init(seed + 4)
    init = prm.clone()
    st = O.AdamState()
    tr = O.BagList(tr_b, tr_y, repeat_num=n_train)
    va, te = O.BagList(va_b, va_y), O.BagList(te_b, te_y)
    ref = {"zs": [O.zs_evaluation(d, w, we, c, topk) for d in (tr, va, te)], "epochs": []}
    best_val, best = 0, None
    for e in range(epochs):
        O.train_epoch(prm, st, tr, w, we, c, topj, topk, (), masks[e])
        row = {"train": O.evaluation(prm, tr, w, we, c, topj, topk), "val": O.evaluation(prm, va, w, we, c, topj, topk)}
        if row["val"]["auc"] > best_val:
            row["test"] = O.evaluation(prm, te, w, we, c, topj, topk)
            best_val, best = row["val"]["auc"], (e, row["test"]["auc"], row["test"]["acc"])
        ref["epochs"].append(row)

    # ---- CUDA path through the reference-shaped loops ------------------------------------------------
    loops.set_prompts(w.to(DEV), we.to(DEV))
    args = types.SimpleNamespace(n_classes=c, topj=topj, topk=topk, discard_classifiers=[], pretrain="conch",
                                 ablation_study="none", cache_scores=True, disable_tqdm=True)
    mk = lambda b, y, rep=None: M.BagLoader(M.BagDataset(M.RaggedBagStore.from_bags(b, y, DEV), repeat_num=rep))
    trl, val, tel = mk(tr_b, tr_y, n_train), mk(va_b, va_y), mk(te_b, te_y)
    model = M.senet(512, 4)
    model.load_state_dict(init.state_dict())
    model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    got = {"zs": [M.zs_evaluation(l, DEV, args) for l in (trl, val, tel)], "epochs": []}
    g_best_val, g_best = 0, None
    for e in range(epochs):
        M.train(model, trl, opt, DEV, args, masks=masks[e])
        row = {"train": M.evaluation(model, trl, DEV, args), "val": M.evaluation(model, val, DEV, args)}
        if row["val"]["auc"] > g_best_val:
            row["test"] = M.evaluation(model, tel, DEV, args)
            g_best_val, g_best = row["val"]["auc"], (e, row["test"]["auc"], row["test"]["acc"])
        got["epochs"].append(row)
    return ref, got, best, g_best, prm, model


def _same(a, b, what):
    assert a["acc"] == b["acc"], "%s acc %r vs %r" % (what, a["acc"], b["acc"])
    assert a["auc"] == b["auc"], "%s auc %r vs %r" % (what, a["auc"], b["auc"])
    assert abs(a["loss"] - b["loss"]) <= 1e-4 * abs(b["loss"]) + 1e-7, "%s loss %r vs %r" % (what, a["loss"], b["loss"])


def test_config1_nsclc_4shot_train_eval():
    ref, got, best, g_best, prm, model = _run_case(c=2, shot=4, n_patches=8000, n_val=49, n_test=209, topj=400,
                                                   topk=10, epochs=25, seed=50)
    for r, g, name in zip(ref["zs"], got["zs"], ("zs_train", "zs_val", "zs_test")):
        _same(g, r, name)
    assert 0.6 < ref["zs"][2]["auc"] < 1.0  # the synthetic cohort is informative, so AUC equality means something
    for e, (r, g) in enumerate(zip(ref["epochs"], got["epochs"])):
        assert set(r) == set(g), "epoch %d: test evaluated on different epochs" % e
        for k in r:
            _same(g[k], r[k], "epoch %d %s" % (e, k))
    assert best == g_best
    for key, t in model.state_dict().items():
        assert (t.cpu() - prm.state_dict()[key]).abs().max().item() < 1e-4, key


def test_config3_rcc_shapes_short():
    """RCC-shaped (C=3, C_ext=7): 3 epochs are enough to cover the three-class AUC path (ovo / macro)."""
    ref, got, best, g_best, _, _ = _run_case(c=3, shot=2, n_patches=3000, n_val=30, n_test=45, topj=400, topk=10,
                                             epochs=3, seed=60)
    for r, g, name in zip(ref["zs"], got["zs"], ("zs_train", "zs_val", "zs_test")):
        _same(g, r, name)
    for e, (r, g) in enumerate(zip(ref["epochs"], got["epochs"])):
        for k in r:
            _same(g[k], r[k], "epoch %d %s" % (e, k))
    assert best == g_best


def test_config4_ebrains30_shapes_short():
    """EBRAINS-30-shaped (C=30, C_ext=34): the tensor-core scoring kernel, 62 selections per slide and the
    30-class ovo/macro AUC inside the full few-shot loop, exact metric equality with the CPU oracle."""
    ref, got, best, g_best, _, _ = _run_case(c=30, shot=1, n_patches=2500, n_val=60, n_test=60, topj=100, topk=10,
                                             epochs=2, seed=70)
    for r, g, name in zip(ref["zs"], got["zs"], ("zs_train", "zs_val", "zs_test")):
        _same(g, r, name)
    for e, (r, g) in enumerate(zip(ref["epochs"], got["epochs"])):
        assert set(r) == set(g)
        for k in r:
            _same(g[k], r[k], "epoch %d %s" % (e, k))
    assert best == g_best


def test_cli_on_h5_dataset_directory(tmp_path):
    """python -m moc_b200.main_moc on a CLAM-style dataset directory: dataset csv + split csv + h5_files/*.h5 (read by
    the native HDF5 reader into the ragged store) + cached prompt matrices, as the reference's driver consumes them
    (main_moc.py:268-293).  The zero-shot metrics it writes must equal the oracle's on the same bags in the
    reference's split order, and a 2-epoch few-shot run must complete and write the reference's output files."""
    import json
    import os
    import subprocess
    import sys
    from moc_b200 import synthetic
    from tests.h5_writer import write_h5
    c, n = 3, 900
    names = ["KICH", "KIRC", "KIRP"]
    w, we = synthetic.prompt_matrices(c)
    root = str(tmp_path)
    os.makedirs(os.path.join(root, "feats", "h5_files"))
    rows, bags = [], {}
    for i in range(21):
        sid, y = "slide_%03d" % i, i % c
        x = synthetic.make_bag(n + 13 * i, y, we, c, seed=900 + i)
        bags[sid] = (x, y)
        rows.append(("case_%d" % i, sid, names[y]))
        write_h5(os.path.join(root, "feats", "h5_files", sid + ".h5"),
                 {"features": x.numpy(), "coords": np.zeros((x.size(0), 2), np.int64)}, batch=256)
    with open(os.path.join(root, "data.csv"), "w") as f:
        f.write("case_id,slide_id,label\n" + "".join("%s,%s,%s\n" % r for r in rows))
    split = {"train": ["slide_%03d" % i for i in (5, 0, 1, 2, 4, 3)], "val": ["slide_%03d" % i for i in range(6, 12)],
             "test": ["slide_%03d" % i for i in range(12, 21)]}
    with open(os.path.join(root, "splits_0.csv"), "w") as f:
        f.write(",train,val,test\n")
        for i in range(9):
            f.write("%d,%s\n" % (i, ",".join(split[k][i] if i < len(split[k]) else "" for k in ("train", "val", "test"))))
    torch.save(w, os.path.join(root, "w.pt"))
    torch.save(we, os.path.join(root, "we.pt"))
    res = os.path.join(root, "results")
    cmd = [sys.executable, "-m", "moc_b200.main_moc", "--dataset", "rcc", "--shot", "2", "--fold", "0", "--topj", "50",
           "--topk", "10", "--epochs", "2", "--seed", "3", "--data_dir", os.path.join(root, "feats"), "--csv",
           os.path.join(root, "data.csv"), "--splits_csv", os.path.join(root, "splits_0.csv"), "--weights",
           os.path.join(root, "w.pt"), "--weights_ext", os.path.join(root, "we.pt"), "--result_dir", res]
    subprocess.run(cmd, check=True, timeout=600, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    zs = json.load(open(os.path.join(res, "zs_results_shot_2_fold_0.json")))
    for key in ("train", "val", "test"):
        ids = sorted(split[key])                       # dataset-csv order = sorted here
        ds = O.BagList([bags[s][0] for s in ids], [bags[s][1] for s in ids], repeat_num=6 if key == "train" else None)
        _same(zs["zs_" + key], O.zs_evaluation(ds, w, we, c, 10), "zs_" + key)
    best = json.load(open(os.path.join(res, "best_results_shot_2_fold_0.json")))
    assert set(best) >= {"zero_shot_test", "best_val", "test_at_best_val", "test_acc_at_best_val", "best_epoch", "best_model_path"}
    assert os.path.exists(os.path.join(res, "best_model_shot_2_fold_0.pt"))


def test_bench_json_contract():
    """bench.py on a small workload: one JSON line with the driver's keys, a live roofline for the streaming kernel,
    an end-to-end number with its copy volumes, a CPU baseline, kernel launches counted."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--slides", "24", "--patches", "6000", "--steps", "3",
                        "--warmup", "3", "--cpu-seconds", "1", "--e2e-steps", "2", "--e2e-host-slides", "8"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "slides_per_sec" and d["n_gpus"] == 1 and d["steps"] == 3 and d["scaling"] == "weak"
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and d["vs_baseline"] is None
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["achieved"] > 0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["algorithmic_bytes_per_launch"] == 24 * 6000 * 2048
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 24 * 6000 * 2048 and e["d2h_bytes_per_step"] == 24 * 2 * 4
    assert e["value"] < d["value"], "the end-to-end number includes the host-to-device copies"
    cb = d["cpu_baseline"]
    from oracle import ref_loader   # the lifted reference when oracle/_ref travelled with the snapshot, else the port
    assert cb["kind"] == ("reference" if ref_loader.available() else "port") and cb["cores"] >= 1 and cb["value"] > 0
    assert e["h2d_ceiling_GBps"] > 0
    assert d["gpu_launches"] >= 3 * 6 and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
