"""Cross-check the C oracle's search control flow against an independent numpy restatement of
SURVEY.md appendix A.2/A.3 built from the same pinned primitives, and check the properties the
reference's own integration tests assert (filtering_integration.rs:69-117, ordering)."""
import numpy as np
import pytest


def _order_key_desc(scores):
    """sort key reproducing cmp_score_descending + stability: finite first by total order."""
    s = np.asarray(scores, np.float32)
    fin = np.isfinite(s)
    bits = s.view(np.int32).astype(np.int64)
    key = np.where(bits < 0, bits ^ 0x7FFFFFFF, bits)   # total_cmp key
    key = np.where(fin, key, -(2 ** 40))
    return np.argsort(-key, kind="stable")


def numpy_search(o, ix, q, p, subset=None):
    S = o.centroid_scores(q, ix.centroids)
    K = ix.num_centroids
    nq = q.shape[0]
    batched = p.centroid_batch_size > 0 and K > p.centroid_batch_size
    cells = set()
    if not batched:
        pool = np.arange(K)
        n_probe = p.n_ivf_probe
        if subset is not None:
            el = set()
            for d in subset:
                if 0 <= d < ix.num_documents:
                    el.update(ix.codes[ix.doc_offsets[d]:ix.doc_offsets[d + 1]].tolist())
            pool = np.array(sorted(el), dtype=np.int64)
            if len(pool):
                scaled = p.n_ivf_probe * ix.num_documents // len(subset) if len(subset) else p.n_ivf_probe
                n_probe = min(max(scaled, p.n_ivf_probe), len(pool))
        n = min(n_probe, len(pool))
        for t in range(nq):
            if n == 0:
                break
            row = S[t, pool]
            cells.update(pool[_order_key_desc(row)[:n]].tolist())
        if p.centroid_score_threshold is not None:
            cells = {c for c in cells if S[:, c][np.isfinite(S[:, c])].max(initial=-np.inf)
                     >= p.centroid_score_threshold}
    else:
        B, n = p.centroid_batch_size, p.n_ivf_probe
        fmax = {}
        per_tok = [[] for _ in range(nq)]
        for c0 in range(0, K, B):
            c1 = min(c0 + B, K)
            for t in range(nq):
                row = S[t, c0:c1]
                # running top-n membership at scan time: fewer than n earlier values are >= row[c]
                heap = []
                for lc in range(c1 - c0):
                    s = row[lc]
                    if len(heap) < n:
                        heap.append((s, c0 + lc)); entered = True
                    else:
                        # heap top = min score, ties -> largest index
                        ti = min(range(len(heap)), key=lambda i: (heap[i][0], -heap[i][1]))
                        entered = s > heap[ti][0]
                        if entered:
                            heap[ti] = (s, c0 + lc)
                    if entered:
                        fmax[c0 + lc] = max(fmax.get(c0 + lc, -np.inf), s)
                per_tok[t].extend(heap)
        for t in range(nq):
            ent = sorted(per_tok[t], key=lambda e: (-e[0], e[1]))[:n]
            cells.update(c for _, c in ent)
        if p.centroid_score_threshold is not None:
            cells = {c for c in cells if fmax.get(c, -np.inf) >= p.centroid_score_threshold}
    cand = set()
    for c in cells:
        cand.update(ix.ivf[ix.ivf_offsets[c]:ix.ivf_offsets[c + 1]].tolist())
    if subset is not None:
        cand &= set(int(x) for x in subset)
    cand = np.array(sorted(cand), dtype=np.int64)
    if len(cand) == 0:
        return cand, np.zeros(0, np.float32)
    approx = np.zeros(len(cand), np.float32)
    for i, d in enumerate(cand):
        cd = ix.codes[ix.doc_offsets[d]:ix.doc_offsets[d + 1]]
        m = S[:, cd].max(axis=1)
        acc = np.float32(0)
        for t in range(nq):
            if m[t] > -np.inf:
                acc = np.float32(acc + m[t])
        approx[i] = acc
    o1 = _order_key_desc(approx)
    keep = cand[o1][:p.n_full_scores][:max(p.n_full_scores // 4, p.top_k)]
    exact = np.array([o.maxsim_score(q, o.get_document_embeddings(ix, int(d))) for d in keep],
                     np.float32)
    o2 = _order_key_desc(exact)[:p.top_k]
    return keep[o2], exact[o2]


@pytest.fixture(scope="module")
def small(oracle):
    docs = oracle.synthetic_corpus(600, 24, dim=64, seed=3, ragged=True)
    ix = oracle.create_index(docs, nbits=4, seed=1, num_partitions=64)
    qs, src = oracle.synthetic_queries(docs, 6, nq=8, seed=5)
    return docs, ix, qs, src


@pytest.mark.parametrize("cbs,thr", [(100_000, 0.4), (100_000, None), (16, 0.4), (16, None), (0, 0.4)])
def test_c_search_equals_numpy_restatement(oracle, small, cbs, thr):
    docs, ix, qs, src = small
    p = oracle.SearchParameters(top_k=10, n_ivf_probe=4, n_full_scores=128,
                                centroid_batch_size=cbs, centroid_score_threshold=thr)
    for q in qs:
        r, tr = oracle.search_one(ix, q, p, trace=True)
        ids, sc = numpy_search(oracle, ix, q, p)
        assert tr.variant == int(cbs > 0 and ix.num_centroids > cbs)
        assert r.passage_ids.tolist() == ids.tolist()
        assert np.array_equal(r.scores, sc)


def test_planted_doc_is_top1_and_scores_descend(oracle, small):
    docs, ix, qs, src = small
    p = oracle.SearchParameters(top_k=5, n_ivf_probe=8, n_full_scores=256)
    hits = 0
    for q, s in zip(qs, src):
        r = oracle.search_one(ix, q, p)
        assert len(r.passage_ids) == len(r.scores) <= 5
        assert all(r.scores[i] >= r.scores[i + 1] for i in range(len(r.scores) - 1))
        hits += int(len(r.passage_ids) > 0 and r.passage_ids[0] == s)
    assert hits >= len(qs) - 1


def test_subset_results_within_subset(oracle, small):
    # filtering_integration.rs:69-117: every returned id belongs to the subset
    docs, ix, qs, src = small
    subset = list(range(0, 600, 3))
    for cbs in (100_000, 16):
        p = oracle.SearchParameters(top_k=10, n_ivf_probe=4, n_full_scores=128, centroid_batch_size=cbs)
        for q in qs:
            r = oracle.search_one(ix, q, p, subset=subset)
            assert set(r.passage_ids.tolist()) <= set(subset)
            ids, sc = numpy_search(oracle, ix, q, p, subset=subset)
            assert r.passage_ids.tolist() == ids.tolist() and np.array_equal(r.scores, sc)


def test_empty_and_degenerate_inputs(oracle, small):
    docs, ix, qs, src = small
    p = oracle.SearchParameters(top_k=10)
    r = oracle.search_one(ix, np.zeros((0, ix.dim), np.float32), p)
    assert len(r.passage_ids) == 0
    r = oracle.search_one(ix, qs[0], p, subset=[])
    assert len(r.passage_ids) == 0
    r = oracle.search_one(ix, qs[0], oracle.SearchParameters(top_k=10, centroid_score_threshold=2.0))
    assert len(r.passage_ids) == 0   # every centroid pruned -> empty result (search.rs:439-445)
    big = oracle.SearchParameters(top_k=5000, n_full_scores=64, centroid_score_threshold=None)
    r = oracle.search_one(ix, qs[0], big)
    assert 0 < len(r.passage_ids) <= 64   # take(n_full_scores) caps before take(n_decompress)


def test_index_dir_roundtrip(oracle, small, tmp_path):
    docs, ix, qs, src = small
    oracle.write_index(ix, str(tmp_path / "idx"), chunk_docs=250, merged=True)
    ix2 = oracle.load_index(str(tmp_path / "idx"))
    for a in ("centroids", "bucket_weights", "codes", "residuals", "doc_lengths", "ivf", "ivf_lengths"):
        assert np.array_equal(getattr(ix, a), getattr(ix2, a)), a
    p = oracle.SearchParameters(top_k=10, n_full_scores=128)
    r1, r2 = oracle.search_one(ix, qs[0], p), oracle.search_one(ix2, qs[0], p)
    assert r1.passage_ids.tolist() == r2.passage_ids.tolist()


def test_sharded_oracle_search_equals_unsharded(oracle, small):
    # po_search_sharded (the --impl reference arm at N > 1) == po_search_one on the concatenated index
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sharded_protocol as sp
    docs, ix, qs, src = small
    for G in (2, 3):
        sh = [sp.make_shard(oracle, ix, g, G) for g in range(G)]
        for cbs in (100_000, 16):
            p = oracle.SearchParameters(top_k=10, n_ivf_probe=4, n_full_scores=64, centroid_batch_size=cbs)
            for q in qs:
                r = oracle.search_sharded([s[0] for s in sh], [s[1] for s in sh], q, p)
                w = oracle.search_one(ix, q, p)
                assert r.passage_ids.tolist() == w.passage_ids.tolist() and np.array_equal(r.scores, w.scores)
