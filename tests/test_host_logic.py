"""CPU-side checks: the C ABI loads and exports every declared symbol, host logic, sharding, gloo collectives."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    from moc_b200 import _lib, build
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "moc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(moc_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libmoc_b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    assert lib.moc_version() >= 100
    assert lib.moc_num_key_planes(3) == 9
    # wide class sets: C+4 planes (log-sum-exp instead of the C softmax planes)
    assert lib.moc_num_key_planes(8) == 19 and lib.moc_num_key_planes(9) == 13 and lib.moc_num_key_planes(30) == 34
    assert [lib.moc_key_plane(3, k) for k in range(6)] == [0, 3, 6, 7, 8, -1]
    assert [lib.moc_key_plane(30, k) for k in range(6)] == [0, -1, 31, 32, 33, 30]
    assert lib.moc_select_capacity(100000, 2, 400) == 2400 and lib.moc_select_capacity(50, 2, 400) == 50
    assert lib.moc_packed_cols(2, 6) == 8 and lib.moc_packed_cols(30, 34) == 36


def test_sass_uses_bulk_copy_engine():
    """The streaming kernel must move patches with the bulk-copy engine (UBLKCP) and wait on mbarriers."""
    from moc_b200 import _lib, build
    build.build()
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out and "SYNCS" in out
    # tensor-core paths: tcgen05.mma (UTC*MMA), accumulators read back (LDTM), TMA tensor copies (UTMALDG), the gate
    # kernel's A operand written to tensor memory (STTM) and the register re-split between warp roles (USETMAXREG)
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "STTM", "USETMAXREG"):
        assert mnemonic in out, mnemonic


def test_ops_fail_loudly_without_cuda():
    from moc_b200 import MocError, slide_process
    from moc_b200 import ops
    with pytest.raises(MocError):
        ops.score_keys(torch.zeros(8, 512), None)
    with pytest.raises(MocError):
        slide_process(torch.zeros(8, 512), torch.zeros(512, 2), torch.zeros(512, 6), 2)
    if not torch.cuda.is_available():
        from moc_b200 import senet
        with pytest.raises(MocError):
            senet(512, 4)(torch.zeros(3, 512))


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: the package, the headers and the
    developer tools must not."""
    for top in ("moc_b200", "include", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                    assert "moc_oracle" not in src, f


def test_flag_masks_follow_reference_quirks():
    from moc_b200 import _lib
    assert _lib.discard_bits(["topk", "bottomk"]) == 9
    assert _lib.discard_bits(["nonsense"]) == 0
    assert _lib.active_bits([], "train") == 15
    assert _lib.active_bits(["topk", "delta_diff"], "train") == 2 | 8
    # evaluation keeps top-k and bottom-k whatever is discarded (main_moc.py:486-492)
    assert _lib.active_bits(["topk", "bottomk", "delta_diff"], "eval") == 1 | 2 | 8


def test_selection_layout_and_synthetic_are_deterministic():
    from moc_b200 import ops, synthetic
    assert ops.selection_layout([0, 50, 20050], 2, 400) == [0, 50, 2450]
    w, we = synthetic.prompt_matrices(3)
    assert w.shape == (512, 3) and we.shape == (512, 7) and torch.equal(we[:, :3], w)
    assert torch.allclose(we.norm(dim=0), torch.ones(7), atol=1e-6)
    a = synthetic.make_bag(257, 1, we, 3, seed=5)
    b = synthetic.make_bag(257, 1, we, 3, seed=5)
    assert torch.equal(a, b) and a.shape == (257, 512)
    assert torch.allclose(a.norm(dim=1), torch.ones(257), atol=1e-5)
    bank, collapsed = synthetic.prompt_bank(3, 64)
    assert bank.shape == (512, 192) and collapsed.shape == (512, 3)
    sizes = synthetic.log_uniform_sizes(1000)
    assert min(sizes) >= 1000 and max(sizes) <= 100000 and 15000 < np.mean(sizes) < 30000


def test_bag_dataset_surface_on_cpu():
    from moc_b200.bag_store import BagDataset, BagLoader, RaggedBagStore
    bags = [torch.randn(n, 512) for n in (3, 5, 2)]
    st = RaggedBagStore.from_bags(bags, [0, 1, 0], device="cpu")
    assert st.total_rows == 10 and st.offsets_h == [0, 3, 8, 10] and torch.equal(st.bag(1), bags[1])
    ds = BagDataset(st, repeat_num=7)
    assert len(ds) == 7 and ds.real_len() == 3
    assert torch.equal(ds[4][0], bags[1]) and ds[4][1] == 1
    with pytest.raises(IndexError):
        ds[7]
    ds.repeat_num = None
    assert len(ds) == 3
    items = list(BagLoader(ds))
    assert len(items) == 3 and items[0][0].shape == (1, 3, 512) and items[2][1].tolist() == [0]


def test_lpt_shards_balance_and_cover():
    from moc_b200 import synthetic
    from moc_b200.dist import lpt_shards
    sizes = synthetic.log_uniform_sizes(500)
    for world in (1, 2, 4, 8):
        shards = lpt_shards(sizes, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(500))
        loads = [sum(sizes[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(sizes)
        assert lpt_shards(sizes, world) == shards


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from moc_b200.dist import Shard, allreduce_sum, init_from_env, barrier_max_ms
rank, local, world = init_from_env("gloo")
sizes = [1000 + 37 * ((i * 7919) %% 101) for i in range(23)]
sh = Shard(sizes, rank, world)
g = torch.Generator().manual_seed(0)
full = torch.randn(23, 3, generator=g)
labels = torch.arange(23) %% 3
ids = torch.tensor(sh.ids, dtype=torch.int64)
out, lab = sh.gather(full[ids], labels[ids])
assert torch.equal(out, full) and torch.equal(lab, labels), "gather mismatch on rank %%d" %% rank
flat = torch.full((33092,), float(rank + 1))
allreduce_sum(flat)
assert float(flat[0]) == sum(range(1, world + 1))
assert barrier_max_ms(float(rank), "cpu") == float(world - 1)
# unseeded replicas (main_moc.py never seeds): ranks whose generators have drifted apart must come out of
# sync_seed() + broadcast_parameters() with identical gate weights and identical half masks
from moc_b200.dist import broadcast_parameters, sync_seed
from moc_b200.model import senet
torch.rand(rank * 3 + 1)
m_before = senet(512, 4)
seed = sync_seed(None)
m = senet(512, 4)
broadcast_parameters(m_before)
mask = (torch.rand(1000) > 0.5).float()
flat = torch.cat([q.detach().flatten() for q in m.parameters()] + [q.detach().flatten() for q in m_before.parameters()] + [mask])
parts = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(parts, flat)
assert all(torch.equal(parts[0], q) for q in parts), "replicas differ after sync_seed/broadcast_parameters"
assert sync_seed(1234) == 1234
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
'''


def test_gloo_world2_gather_and_allreduce(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


def _summary_tree(root):
    import json
    os.makedirs(os.path.join(root, "1_shot"))
    os.makedirs(os.path.join(root, "2_shot"))
    os.makedirs(os.path.join(root, "4_shot"))
    for fold in range(5):
        full = {"zero_shot_test": {"loss": 1.0, "acc": 0.5 + 0.01 * fold, "auc": 0.6 + 0.02 * fold}, "best_val": 0.9,
                "test_at_best_val": 0.7 + 0.03 * fold, "test_acc_at_best_val": 0.65 + 0.01 * fold, "best_epoch": fold}
        json.dump(full, open(os.path.join(root, "1_shot", "best_results_shot_1_fold_%d.json" % fold), "w"))
        nozs = dict(full, zero_shot_test=-1)          # --check_zeroshot off: main_moc.py:599 leaves -1
        json.dump(nozs, open(os.path.join(root, "2_shot", "best_results_shot_2_fold_%d.json" % fold), "w"))
        abl = {"loss": 0.3, "acc": 0.8 - 0.01 * fold, "auc": 0.85 + 0.01 * fold}
        json.dump(abl, open(os.path.join(root, "4_shot", "ablation_results_avg_shot_4_fold_%d.json" % fold), "w"))


def test_cli_summary_mode(tmp_path, capsys):
    """--summary / --summary_dir (main_moc.py:53-130): the three result-file shapes, the trailing mean row, the
    "summary failed" line for a shot without files; compared with the reference's own block where it is present."""
    import types
    import pandas as pd
    from moc_b200 import main_moc
    ours = str(tmp_path / "ours")
    _summary_tree(ours)
    main_moc.run(main_moc.get_args(["--summary", "--summary_dir", ours]))
    out = capsys.readouterr().out
    assert "start summary" in out and "shot 8 summary failed" in out and "end summary" in out
    s1 = pd.read_csv(os.path.join(ours, "summary_1.csv"))
    assert list(s1.columns) == ["fold", "test_auc", "zs_test_auc", "test_acc", "zs_test_acc"]
    assert s1["fold"].tolist() == ["0", "1", "2", "3", "4", "mean"] and abs(s1["test_auc"].iloc[-1] - 0.76) < 1e-12
    assert list(pd.read_csv(os.path.join(ours, "summary_2.csv")).columns) == ["fold", "test_auc", "test_acc"]
    s4 = pd.read_csv(os.path.join(ours, "summary_4.csv"))
    assert list(s4.columns) == ["fold", "auc", "acc"] and abs(s4["acc"].iloc[-1] - 0.78) < 1e-12
    assert not os.path.exists(os.path.join(ours, "summary_8.csv"))

    from oracle import ref_loader
    if not ref_loader.available():
        return
    import ast
    import json
    from glob import glob
    import numpy as np
    path = os.path.join(ref_loader.REFERENCE_ROOT, "main_moc.py")
    tree = ast.parse(open(path).read())
    block = [n for n in tree.body if isinstance(n, ast.If) and isinstance(n.test, ast.Attribute) and n.test.attr == "summary"]
    assert len(block) == 1
    ref = str(tmp_path / "ref")
    _summary_tree(ref)
    glb = {"args": types.SimpleNamespace(summary=True, summary_dir=ref), "os": os, "json": json, "np": np, "pd": pd,
           "glob": glob, "print": print, "exit": lambda *a: None}
    exec(compile(ast.Module(body=block, type_ignores=[]), path, "exec"), glb)
    for shot in (1, 2, 4):
        a = open(os.path.join(ours, "summary_%d.csv" % shot)).read()
        b = open(os.path.join(ref, "summary_%d.csv" % shot)).read()
        assert a == b, shot
    assert not os.path.exists(os.path.join(ref, "summary_8.csv"))


def test_bench_reference_arm_contract():
    """bench.py --impl reference runs on the host alone (no GPU) and prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--slides", "3", "--patches", "500"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "slides_per_sec" and d["unit"] == "slides/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["n_classes"] == 2 and d["config"]["topj"] == 400 and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # a non-zero rank under torchrun exits 0 without work and without output
    r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, cwd=root, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""
