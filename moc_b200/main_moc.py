"""Command line of the MOC few-shot run on B200 - the reference's flags (main_moc.py:29-46) plus the options
needed to run without a CONCH checkpoint or slide data:

    python -m moc_b200.main_moc --fold 0 --shot 4 --topj 400 --topk 10 --dataset nsclc --synthetic
    torchrun --nproc-per-node 8 -m moc_b200.main_moc --dataset ebrains30 --shot 16 --synthetic --n_patches 50000

Real data: ``--data_dir`` with CLAM-style ``h5_files/<slide_id>.h5`` (native reader) or ``pt_files/<slide_id>.pt`` bags, ``--csv`` (slide_id,label) and
``--splits_csv`` (train,val,test columns, as ``splits/*_fewshot/*shots/splits_*.csv``), and ``--weights`` /
``--weights_ext`` pointing at the cached prompt matrices the reference writes to ``models/classifier_weights``.
Outputs keep the reference's names and JSON schema (zs_results_*, best_results_*, best_model_*.pt).
"""
from __future__ import annotations

import argparse
import os

import torch

DATASETS = {
    # name: (classes, label names, synthetic val/test sizes following the shipped 4-shot splits where they exist)
    "nsclc": (2, ["LUAD", "LUSC"], 49, 209),
    "rcc": (3, ["KICH", "KIRC", "KIRP"], 65, 222),
    "ebrains30": (30, None, 120, 400),
}


def get_args(argv=None):
    p = argparse.ArgumentParser(description="Configurations for WSI Training")
    p.add_argument("--fold", type=int, default=0, help="fold number")
    p.add_argument("--shot", type=int, default=1, help="split number")
    p.add_argument("--topj", type=int, default=10, help="topj for classifier selection")
    p.add_argument("--topk", type=int, default=10, help="topk for final pooling")
    p.add_argument("--result_dir", type=str, default="results/moc_train", help="result directory")
    p.add_argument("--dataset", type=str, default="nsclc", choices=sorted(DATASETS), help="dataset name")
    p.add_argument("--pretrain", type=str, default="conch", choices=["conch"], help="pretrain model")
    p.add_argument("--disable_tqdm", action="store_true", help="accepted for compatibility (no progress bars here)")
    p.add_argument("--discard_classifiers", nargs="+", default=[], help="topk, delta_softmax, delta_diff, bottomk")
    p.add_argument("--load_weight", type=bool, default=True, help="load stored classifier weight")
    p.add_argument("--check_zeroshot", type=bool, default=True, help="get zero-shot results")
    p.add_argument("--ablation_study", type=str, default="none", choices=["none", "avg", "sum", "max"])
    p.add_argument("--summary", action="store_true", help="summary results, no training")
    p.add_argument("--summary_dir", type=str, default="")
    # additions
    p.add_argument("--synthetic", action="store_true", help="synthetic CONCH-shaped bags and random prompt matrices")
    p.add_argument("--n_patches", type=int, default=8000)
    p.add_argument("--n_val", type=int, default=None)
    p.add_argument("--n_test", type=int, default=None)
    p.add_argument("--epochs", type=int, default=25, help="the reference hard-codes 25 (main_moc.py:611)")
    p.add_argument("--seed", type=int, default=None, help="torch.manual_seed before model init (reference: unseeded)")
    p.add_argument("--cache_scores", action="store_true", help="score every bag once per run (bit-identical results)")
    p.add_argument("--dp_microbatch", type=int, default=None,
                   help="data-parallel training: sum the gate gradients of this many consecutive slides (spread over the "
                        "ranks, one NCCL all-reduce) per Adam step; changes the trajectory, off by default")
    p.add_argument("--data_dir", type=str, default=None)
    p.add_argument("--csv", type=str, default=None)
    p.add_argument("--splits_csv", type=str, default=None)
    p.add_argument("--weights", type=str, default=None)
    p.add_argument("--weights_ext", type=str, default=None)
    return p.parse_args(argv)


def summarize_results(summary_dir: str, shots=(1, 2, 4, 8), folds=(0, 1, 2, 3, 4)) -> None:
    """``--summary`` (main_moc.py:53-130): per shot, collect the five folds' result files under
    ``<summary_dir>/<shot>_shot`` into ``<summary_dir>/summary_<shot>.csv`` with a trailing "mean" row.  Three file
    shapes are tried in the reference's order: best_results with zero-shot numbers (columns fold, test_auc,
    zs_test_auc, test_acc, zs_test_acc), best_results without them (fold, test_auc, test_acc), ablation results
    (fold, auc, acc); a shot whose files fit none prints "shot <n> summary failed"."""
    import glob
    import json

    import numpy as np
    import pandas as pd

    def load(shot_dir, shot, fold, pattern="best_results"):
        if pattern is None:
            path = glob.glob(os.path.join(shot_dir, "*_shot_%d_fold_%d.json" % (shot, fold)))[0]
        else:
            path = os.path.join(shot_dir, "%s_shot_%d_fold_%d.json" % (pattern, shot, fold))
        with open(path) as f:
            return json.load(f)

    def table(shot_dir, shot, columns, pattern="best_results"):
        cols = {name: [] for name in columns}
        for fold in folds:
            res = load(shot_dir, shot, fold, pattern)
            for name, pick in columns.items():
                cols[name].append(pick(res))
        out = {"fold": list(folds) + ["mean"]}
        for name, vals in cols.items():
            out[name] = vals + [np.mean(vals)]
        return pd.DataFrame(out)

    shapes = [
        ({"test_auc": lambda r: r["test_at_best_val"], "zs_test_auc": lambda r: r["zero_shot_test"]["auc"],
          "test_acc": lambda r: r["test_acc_at_best_val"], "zs_test_acc": lambda r: r["zero_shot_test"]["acc"]}, "best_results"),
        ({"test_auc": lambda r: r["test_at_best_val"], "test_acc": lambda r: r["test_acc_at_best_val"]}, "best_results"),
        ({"auc": lambda r: r["auc"], "acc": lambda r: r["acc"]}, None),
    ]
    print("start summary")
    for shot in shots:
        shot_dir = summary_dir + "/%d_shot" % shot
        summary_file = os.path.join(summary_dir, "summary_%d.csv" % shot)
        for columns, pattern in shapes:
            try:
                if os.path.exists(summary_file):
                    os.remove(summary_file)
                table(shot_dir, shot, columns, pattern).to_csv(summary_file, index=False)
                break
            except Exception:
                continue
        else:
            print("shot %d summary failed" % shot)
    print("end summary")


def _real_stores(args, device, rank=0, world=1):
    from .datasets import Generic_MIL_Dataset
    from .dist import Shard
    n_classes, names, _, _ = DATASETS[args.dataset]
    if names:
        label_dict = {n: i for i, n in enumerate(names)}
    else:   # no label names shipped for this dataset (ebrains30): classes in order of first appearance in the csv
        import pandas as pd
        label_dict = {n: i for i, n in enumerate(pd.read_csv(args.csv, dtype=str)["label"].drop_duplicates())}
    # main_moc.py:268-289
    dataset = Generic_MIL_Dataset(csv_path=args.csv, data_dir=args.data_dir, shuffle=False, seed=1, print_info=True,
                                  label_dict=label_dict, patient_strat=False, ignore=[])
    # h5_files/ as the reference's driver reads them (load_from_h5(True)); pt_files/ when only those exist
    use_h5 = os.path.isdir(os.path.join(args.data_dir, "h5_files"))
    splits = dataset.return_splits(from_id=False, csv_path=args.splits_csv, repeat_num=int(args.shot) * n_classes)
    stores, shards = {}, {}
    for key, sp in zip(("train", "val", "test"), splits):
        if sp is None:
            raise SystemExit("the split file %s leaves a split empty" % args.splits_csv)
        sp.load_full_path(True)
        sp.load_from_h5(use_h5)
        if world > 1 and key != "train":   # eval splits are slide-sharded (LPT on patch count); few-shot bags replicated
            shards[key] = Shard(sp.bag_sizes(), rank, world)
            stores[key] = sp.to_store(device, rows=shards[key].ids)
        else:
            stores[key] = sp.to_store(device)
    return stores, shards


def run(args):
    from . import loops, synthetic
    from .bag_store import BagDataset, BagLoader, RaggedBagStore
    from .dist import Shard, broadcast_parameters, init_from_env, sync_seed
    from .model import senet

    if getattr(args, "summary", False):      # no training, no GPU (main_moc.py:53-130)
        summarize_results(args.summary_dir)
        return None
    rank, local, world = init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("moc_b200 needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    n_classes, _, n_val, n_test = DATASETS[args.dataset]
    args.n_classes = n_classes
    n_val = args.n_val or n_val
    n_test = args.n_test or n_test

    if args.synthetic:
        w, w_ext = synthetic.prompt_matrices(n_classes, device=device)
        n_train = args.shot * n_classes
        sizes = {"train": [args.n_patches] * n_train, "val": [args.n_patches] * n_val, "test": [args.n_patches] * n_test}
        seeds = {"train": 11 + args.fold, "val": 22 + args.fold, "test": 33 + args.fold}
        stores, shards = {}, {}
        for key in ("train", "val", "test"):
            labels = [i % n_classes for i in range(len(sizes[key]))]
            ids = list(range(len(sizes[key])))
            if world > 1 and key != "train":  # eval splits are slide-sharded; few-shot bags are replicated
                shards[key] = Shard(sizes[key], rank, world)
                ids = shards[key].ids
            st = RaggedBagStore.synthetic([sizes[key][i] for i in ids], n_classes, w_ext, device=device,
                                          labels=[labels[i] for i in ids], cohort_seed=seeds[key])
            # slide i of the cohort must be the same bag whichever rank holds it
            if world > 1 and key != "train":
                for k, i in enumerate(ids):
                    synthetic.make_bag(sizes[key][i], labels[i], w_ext, n_classes, synthetic.slide_seed(seeds[key], i),
                                       device=device, out=st.bag(k))
            stores[key] = st
    else:
        if not (args.data_dir and args.csv and args.splits_csv and args.weights and args.weights_ext):
            raise SystemExit("real data needs --data_dir --csv --splits_csv --weights --weights_ext (or use --synthetic)")
        w = torch.load(args.weights, map_location=device).float()
        w_ext = torch.load(args.weights_ext, map_location=device).float()
        stores, shards = _real_stores(args, device, rank, world)

    loops.set_prompts(w, w_ext)
    loaders = {}
    for key, st in stores.items():
        ds = BagDataset(st, repeat_num=int(args.shot) * n_classes if key == "train" else None)
        if key in shards:
            ds.shard = shards[key]
        loaders[key] = BagLoader(ds)

    # replicas must stay identical: one seed for all ranks (rank 0's --seed, or a random one it draws when the run is
    # unseeded like the reference's) fixes the gate's initial weights and the per-step half masks everywhere; the
    # parameters are broadcast as well, so nothing depends on every rank consuming its generator identically up to here
    sync_seed(args.seed)
    model = senet(512, 4).to(device)
    broadcast_parameters(model)
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = loops.main(args, model, optimizer, loaders["train"], loaders["val"], loaders["test"], device,
                     num_epoch=args.epochs, is_main=(rank == 0))
    torch.cuda.synchronize()
    if rank == 0:
        print("zero-shot + %d epochs (train %d steps, eval train/val every epoch, test on improvement): %.2f s"
              % (args.epochs, len(loaders["train"]), time.perf_counter() - t0))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return res


if __name__ == "__main__":
    a = get_args()
    os.makedirs(a.result_dir, exist_ok=True)
    run(a)
