"""Shared comparison rules for the parity tests."""
import numpy as np
import torch

RTOL, ATOL = 1e-3, 1e-6  # north_star: scores and logits within 1e-3 relative (plus an absolute floor at zero)


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


def params_from_golden(g, prefix, device):
    from moc_b200.ops import HeadParams
    t = lambda k: torch.from_numpy(np.ascontiguousarray(g[prefix + k])).to(device)
    return HeadParams(t("model_0_weight"), t("model_0_bias"), t("model_2_weight"), t("model_2_bias"))
