"""Shared comparison rules for the parity tests."""
import numpy as np
import torch

RTOL, ATOL = 1e-3, 1e-6  # north_star: scores and logits within 1e-3 relative (plus an absolute floor at zero)


def full_keys(keys, c):
    """Key planes in the full 2C+3 layout [L | softmax | diff | bg sum | bg max] whatever layout the scoring kernels use
    for c classes (moc_expand_keys: wide class sets store C+5 planes and rebuild the softmax planes on the fly)."""
    from moc_b200 import ops
    return ops.expand_keys(keys, c)


def close(a, b, rtol=RTOL, atol=ATOL):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b)
    tol = rtol * np.abs(b) + atol
    assert (err <= tol).all(), "max abs err %.3e (tol %.3e) at %s" % (
        err.max(), tol.flat[err.argmax()], np.unravel_index(err.argmax(), err.shape))
    return float(err.max()) if err.size else 0.0


def assert_topj_set(got, ref, key, j, largest=True, rtol=RTOL, atol=ATOL):
    """Index sets must be identical except for rows whose key lies within tolerance of the rank-j value."""
    got, ref = set(int(v) for v in got), set(int(v) for v in ref)
    if got == ref:
        return
    key = np.asarray(key, dtype=np.float64)
    j = min(j, key.shape[0])
    srt = np.sort(key)
    thr = srt[-j] if largest else srt[j - 1]
    for r in got ^ ref:
        assert abs(key[r] - thr) <= rtol * abs(thr) + atol, "row %d key %.9g vs threshold %.9g" % (r, key[r], thr)
    assert len(got) == len(ref)


def selection_planes(c, discard=()):
    """(key plane, largest?) of every top-J selection that enters the union (main_moc.py:341-352) in the key layout
    [L_0..L_{C-1} | softmax_0.. | |top1-top2| | sum bg | max bg]."""
    d = set(discard or ())
    planes = []
    if "topk" not in d:
        planes += [(cc, True) for cc in range(c)]
    if "delta_softmax" not in d:
        planes += [(c + cc, True) for cc in range(c)]
    if "delta_diff" not in d:
        planes.append((2 * c, True))
    if "bottomk" not in d:
        planes.append((2 * c + 1, False))
    return planes


def assert_union_set(got, ref, okeys, c, j, discard=(), mask=None, rtol=RTOL, atol=ATOL):
    """``selected_index`` as a set: identical to the reference's, except that a row may be swapped for another when
    both sit within tolerance of the rank-j value of an active selection plane (torch.topk's tie order is
    unspecified).  ``okeys`` are the oracle's key planes [2C+3, N] of the (masked) bag.  Fails otherwise.
    Returns the rows common to both."""
    got_s, ref_s = set(int(v) for v in got), set(int(v) for v in ref)
    if got_s == ref_s:
        return sorted(got_s)
    okeys = np.asarray(okeys, dtype=np.float64)
    if mask is not None:
        okeys = okeys[:, np.asarray(mask, dtype=bool)]
    n = okeys.shape[1]
    jj = min(j, n)
    thr = []
    for plane, largest in selection_planes(c, discard):
        srt = np.sort(okeys[plane])
        thr.append((plane, srt[-jj] if largest else srt[jj - 1]))
    for r in sorted(got_s ^ ref_s):
        near = [abs(okeys[p_, r] - t) <= rtol * abs(t) + atol for p_, t in thr]
        assert any(near), "row %d is in only one of the two selections and is at no selector's rank-%d threshold" % (r, jj)
    assert abs(len(got_s) - len(ref_s)) <= len(thr), (len(got_s), len(ref_s))
    return sorted(got_s & ref_s)


def params_from_golden(g, prefix, device):
    from moc_b200.ops import HeadParams
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[prefix + k])).to(device)
    return HeadParams(t("model_0_weight"), t("model_0_bias"), t("model_2_weight"), t("model_2_bias"))
