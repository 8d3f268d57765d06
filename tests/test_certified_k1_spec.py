"""Executable statement of the margins the tensor-core a2 relies on (next-plaid_b200/csrc/k_scores_tc.cuh).  The
16-bit score table is built from an estimate t with |t - e| * scale <= err < 1 code (e = the exact pinned-order
score), so a code moves by at most E = 1 and
  * probe:    an entry of a token's exact top n has code~ >= tau~ - (2E + 1)   (tau~ = n-th largest chunk maximum)
  * a5:       a doc of the exact top M has L~ >= (M-th largest L~) - W,  W = nq (1.004 + 2 err) + nq^2/256 + 4
  * re-check: the code attaining a doc's exact per-token maximum has code~ >= (largest code~ of the doc) - (2E + 1)
and of the two certificates of the MaxSim stage (k_maxsim_tc.cuh): the docs that can reach the top_k, and inside them the
(token, query token) pairs that can hold a per-token maximum.
Pure numpy, adversarial perturbations of the full err; the GPU tests run the kernels themselves."""
import numpy as np
import pytest

ERR = 0.48          # k1_err_codes(128) in engine.cu


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
@pytest.mark.parametrize("err", [0.0, 0.3, ERR, 0.99])
def test_probe_margin_keeps_the_exact_top_n(seed, err):
    rng, S, R, scale = _setup(seed)
    nq, K = S.shape
    n, chunk, margin = 8, 256, 3
    eps1 = err / scale
    # adversarial estimate: push the exact winners down and everything else up by the full error
    key = S.astype(np.float64) * 1e6 - np.arange(K)[None, :] * 1e-3          # score desc, index asc
    win = np.argsort(-key, axis=1)[:, :n]
    delta = np.full(S.shape, eps1, np.float64)
    np.put_along_axis(delta, win, -eps1, axis=1)
    ct = _codes(S.astype(np.float64) + delta, R, scale)
    for q in range(nq):
        cmax = ct[q].reshape(-1, chunk).max(1)
        tau = np.sort(cmax)[-n]
        collected = set(np.nonzero(ct[q] >= tau - margin)[0].tolist())
        assert set(win[q].tolist()) <= collected, (seed, err, q)
        assert len(collected) < K // 8          # the margin is not vacuous


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("err", [0.0, 0.3, ERR])
@pytest.mark.parametrize("nq", [16, 48])
def test_band_keeps_the_exact_cut(seed, err, nq):
    rng, S, R, scale = _setup(seed, nq=nq)
    K = S.shape[1]
    n_docs, M = 600, 40
    docs = [np.unique(rng.integers(K, size=rng.integers(4, 60))) for _ in range(n_docs)]
    exact = np.zeros(n_docs, np.float32)
    for i, d in enumerate(docs):                       # q-ordered fp32 sum of the per-token maxima (search.rs:305-324)
        acc = np.float32(0)
        for q in range(nq):
            acc = np.float32(acc + np.float32(S[q, d].max()))
        exact[i] = acc
    eps1 = err / scale
    # adversarial: docs of the exact top M pushed down, the rest up
    order = np.lexsort((np.arange(n_docs), -exact.astype(np.float64)))
    top = set(order[:M].tolist())
    L = np.zeros(n_docs, np.int64)
    for i, d in enumerate(docs):
        sgn = -1.0 if i in top else 1.0
        ct = _codes(S[:, d].astype(np.float64) + sgn * eps1, R, scale)
        L[i] = ct.max(1).sum()
    W = int(np.ceil(nq * (1.004 + 2 * err) + nq * nq / 256.0 + 4))
    cut = np.sort(L)[-M]
    survivors = set(np.nonzero(L >= cut - W)[0].tolist())
    assert top <= survivors, (seed, err)
    # the engine's integer form of the same band: (ceil(1.004 + 2 err) + 1) * nq + 8
    assert W <= (int(np.ceil(1.004 + 2 * err)) + 1) * nq + 8


@pytest.mark.parametrize("seed", range(6))
def test_recheck_margin_contains_the_exact_argmax(seed):
    rng, S, R, scale = _setup(seed)
    nq, K = S.shape
    eps1 = ERR / scale
    for _ in range(200):
        d = np.unique(rng.integers(K, size=rng.integers(2, 80)))
        for q in range(nq):
            e = S[q, d].astype(np.float64)
            star = int(np.argmax(e))
            delta = np.full(len(d), eps1)
            delta[star] = -eps1
            ct = _codes(e + delta, R, scale)
            assert ct[star] >= ct.max() - 3, (seed, q)


# ------------------------------------------------------------------------------------------
# The MaxSim stage (k_maxsim_tc.cuh): every similarity estimate is within eps of the exact one.
#   * pass 1: a doc of the exact top_k has estimate sum >= (top_k-th largest estimate sum) - 2 nq eps
#   * pass 2: the token holding a (doc, query token) exact maximum has estimate >= (largest estimate of the pair) - 2 eps,
#             so the exact MaxSim of a doc is the sum over q of the maxima of the exact sims over the listed tokens only
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("mode", ["random", "adversarial"])
def test_maxsim_filter_and_pair_band_are_supersets(seed, mode):
    rng = np.random.default_rng(100 + seed)
    n_docs, nq, top_k, eps = 200, 12, 10, 2e-3
    lens = rng.integers(1, 40, n_docs)
    sims = [rng.uniform(-1, 1, (l, nq)).astype(np.float64) for l in lens]
    for d in range(0, n_docs, 7):                     # near ties inside a doc and between docs
        sims[d][-1] = sims[d][0] + rng.uniform(-eps, eps, nq) * 0.5
        if d + 1 < n_docs:
            sims[d + 1] = sims[d][: max(1, min(len(sims[d]), lens[d + 1]))].copy()
    exact = np.array([s.max(0).sum() for s in sims])
    if mode == "random":
        est = [s + rng.uniform(-eps, eps, s.shape) for s in sims]
    else:   # push every per-token winner down and everything else up by the full bound
        est = []
        for s in sims:
            e = s + eps
            e[s.argmax(0), np.arange(nq)] = s.max(0) - eps
            est.append(e)
    est_sum = np.array([e.max(0).sum() for e in est])
    # pass 1
    order = np.argsort(-exact, kind="stable")
    true_top = set(order[:top_k].tolist())
    tau = np.sort(est_sum)[::-1][top_k - 1]
    survivors = set(np.nonzero(est_sum >= tau - 2 * nq * eps - 1e-12)[0].tolist())
    assert true_top <= survivors
    # pass 2 on the survivors: only listed (token, q) pairs are evaluated exactly
    n_pairs = 0
    for d in survivors:
        listed = est[d] >= est[d].max(0)[None, :] - 2 * eps - 1e-12
        n_pairs += int(listed.sum())
        assert listed.any(0).all()
        got = np.where(listed, sims[d], -np.inf).max(0).sum()
        assert got == exact[d]
    if mode == "random":    # the band is narrow: about one pair per (doc, q), not every token
        assert n_pairs < 2.5 * nq * len(survivors)
