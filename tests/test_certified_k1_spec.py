"""Executable statement of the margins the planned tensor-core a2 needs (DESIGN.md "Prepared", profiles/r01_summary.md
round-2 target 1).  When the 16-bit score table comes from an estimate S~ with |S~ - S| <= eps1 instead of the exact
fp32 S, each code moves by at most E = ceil(eps1 * scale) + 1, and
  * probe:  entries with code~ < tau~ - 2E cannot be in a token's top n   (tau~ = n-th largest chunk maximum of code~)
  * a5:     docs whose 16-bit sum is more than (3.25 + 2 (1 + eps1*scale) - 2) * nq + 8 below the M-th largest sum cannot
            make the cut
so collecting / re-checking with those margins keeps every exact winner.  Pure numpy, adversarial perturbations."""
import numpy as np
import pytest


def _codes(S, R, scale):
    t = np.floor(np.float32(S) * np.float32(scale) + np.float32(R * scale))
    return np.clip(t, 0, 65535).astype(np.int64)


def _setup(seed, nq=16, K=4096, dim=32):
    rng = np.random.default_rng(seed)
    Q = rng.standard_normal((nq, dim)).astype(np.float32)
    Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    C = rng.standard_normal((K, dim)).astype(np.float32)
    C /= np.linalg.norm(C, axis=1, keepdims=True)
    C[rng.integers(K, size=64)] = C[rng.integers(K, size=64)]          # duplicates -> exact ties
    S = (Q @ C.T).astype(np.float32)
    R = 1.0001
    scale = 65535.0 / (2 * R)
    return rng, S, R, scale


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("eps1", [0.0, 2.5e-5, 2e-4])
def test_probe_margin_keeps_the_exact_top_n(seed, eps1):
    rng, S, R, scale = _setup(seed)
    nq, K = S.shape
    n, chunk = 8, 256
    E = int(np.ceil(eps1 * scale)) + 1
    # adversarial estimate: push the exact winners down and everything else up by the full eps1
    key = S.astype(np.float64) * 1e6 - np.arange(K)[None, :] * 1e-3          # score desc, index asc
    win = np.argsort(-key, axis=1)[:, :n]
    delta = np.full(S.shape, eps1, np.float64)
    np.put_along_axis(delta, win, -eps1, axis=1)
    delta *= rng.uniform(0.5, 1.0, S.shape)
    ct = _codes(S.astype(np.float64) + delta, R, scale)
    for q in range(nq):
        cmax = ct[q].reshape(-1, chunk).max(1)
        tau = np.sort(cmax)[-n]
        collected = set(np.nonzero(ct[q] >= tau - 2 * E)[0].tolist())
        assert set(win[q].tolist()) <= collected, (seed, eps1, q)
        # and the margin is not vacuous: only a small part of the table is collected
        assert len(collected) < K // 4 or eps1 > 1e-4


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("eps1", [0.0, 2.5e-5, 1e-4])
def test_band_margin_keeps_the_exact_cut(seed, eps1):
    rng, S, R, scale = _setup(seed)
    nq, K = S.shape
    n_docs, M = 600, 40
    docs = [np.unique(rng.integers(K, size=rng.integers(4, 60))) for _ in range(n_docs)]
    exact = np.array([np.float32(sum(np.float32(S[q, d].max()) for q in range(nq))) for d in docs], np.float32)
    delta = rng.uniform(-eps1, eps1, S.shape)
    ct = _codes(S.astype(np.float64) + delta, R, scale)
    L = np.array([sum(int(ct[q, d].max()) for q in range(nq)) for d in docs], np.int64)
    e = eps1 * scale
    W = int(np.ceil((3.25 + 2 * e) * nq + 8))
    cut = np.sort(L)[-M]
    survivors = set(np.nonzero(L >= cut - W)[0].tolist())
    order = np.lexsort((np.arange(n_docs), -exact.astype(np.float64)))          # exact score desc, doc id asc
    assert set(order[:M].tolist()) <= survivors, (seed, eps1)
    if eps1 == 0.0:
        assert W <= 4 * nq + 8           # today's band (k_select_u32: 4 nq + 8) covers the exact-table case
