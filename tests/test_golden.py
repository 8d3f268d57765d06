"""Committed fixtures (tests/golden/, made by tools/make_golden.py): the reference's known-answer vectors as
data, and a small stored index with the oracle's results.  CPU tests pin the oracle to them; the GPU
test runs the CUDA path through the C-ABI against the same files."""
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(oracle):
    z = np.load(os.path.join(G, "search_small.npz"))
    ix = oracle.Index(z["centroids"], z["bucket_weights"], z["bucket_cutoffs"], z["codes"].astype(np.int64),
                      z["residuals"], z["doc_lengths"].astype(np.int64), z["ivf"].astype(np.int64), z["ivf_lengths"],
                      int(z["nbits"]))
    params = json.load(open(os.path.join(G, "search_small.params.json")))
    return z, ix, params


def test_oracle_matches_reference_kats_from_file(oracle):
    k = json.load(open(os.path.join(G, "kats.json")))
    a = k["maxsim_1p7"]
    assert abs(oracle.maxsim_score(np.array(a["query"], np.float32), np.array(a["doc"], np.float32)) - a["score"]) < a["tol"]
    a = k["maxsim_nan_row_8p0"]
    q = np.tile(np.array(a["query_row"], np.float32), (a["n_query"], 1))
    d = np.tile(np.array(a["doc_row"], np.float32), (a["n_query"], 1))
    d[a["nan_row"], 0] = np.nan
    assert abs(oracle.maxsim_score(q, d) - a["score"]) < a["tol"]
    a = k["rerank_2_1_0"]
    got = [oracle.maxsim_score(np.array(a["query"], np.float32), np.array(d, np.float32)) for d in a["docs"]]
    assert np.allclose(got, a["scores"], atol=a["tol"])
    for name in ("assign_to_centroids", "compress_into_codes"):
        a = k[name]
        assert oracle.compress_into_codes(np.array(a["embeddings"], np.float32),
                                          np.array(a["centroids"], np.float32)).tolist() == a["codes"]
    a = k["byte_layout"]
    cut = np.arange(1, 16, dtype=np.float32)
    assert oracle.quantize_residuals(np.array([[b + 0.5 for b in a["buckets"]]], np.float32), cut, a["nbits"])[0, 0] == a["byte"]
    a = k["quantile"]
    assert np.allclose(oracle.quantiles(np.array(a["values"], np.float32), a["q"]), a["expect"], atol=1e-6)
    a = k["find_outliers"]
    assert oracle.find_outliers(np.array(a["embeddings"], np.float32), np.array(a["centroids"], np.float32),
                                a["threshold_sq"]).tolist() == a["outliers"]
    a = k["defaults"]
    p = oracle.SearchParameters()
    assert (p.batch_size, p.n_full_scores, p.top_k, p.n_ivf_probe, p.centroid_batch_size, p.centroid_score_threshold) == \
        (a["batch_size"], a["n_full_scores"], a["top_k"], a["n_ivf_probe"], a["centroid_batch_size"], a["centroid_score_threshold"])


def test_oracle_reproduces_stored_search_results(oracle):
    z, ix, params = _fixture(oracle)
    assert np.array_equal(oracle.get_document_embeddings(ix, 7), z["decompressed_doc7"])
    assert np.array_equal(oracle.centroid_scores(z["queries"][0], ix.centroids), z["centroid_scores_q0"])
    for pi, kw in enumerate(params):
        p = oracle.SearchParameters(**kw)
        for i, q in enumerate(z["queries"]):
            r = oracle.search_one(ix, q, p)
            n = len(r.passage_ids)
            assert r.passage_ids.tolist() == z[f"ids_{pi}"][i, :n].tolist() and (z[f"ids_{pi}"][i, n:] == -1).all()
            assert np.array_equal(r.scores, z[f"scores_{pi}"][i, :n])
        # the planted source doc ranks first
    p = oracle.SearchParameters(**params[0])
    assert sum(int(oracle.search_one(ix, q, p).passage_ids[0] == s) for q, s in zip(z["queries"], z["source_docs"])) >= 5


@pytest.mark.gpu
def test_gpu_reproduces_stored_search_results(oracle):
    import next_plaid_b200 as npb
    z, ix, params = _fixture(oracle)
    gpu = npb.MmapIndex.from_arrays(ix.centroids, ix.bucket_weights, ix.codes, ix.residuals, ix.doc_lengths, ix.ivf,
                                    ix.ivf_lengths, ix.nbits)
    assert np.array_equal(gpu.get_document_embeddings(7), z["decompressed_doc7"])
    assert np.array_equal(gpu.centroid_scores(z["queries"][0]), z["centroid_scores_q0"])
    for pi, kw in enumerate(params):
        res = gpu.search_batch(list(z["queries"]), npb.SearchParameters(**kw))
        for i, r in enumerate(res):
            n = len(r.passage_ids)
            assert r.passage_ids.tolist() == z[f"ids_{pi}"][i, :n].tolist() and (z[f"ids_{pi}"][i, n:] == -1).all()
            assert np.array_equal(r.scores, z[f"scores_{pi}"][i, :n])
    k = json.load(open(os.path.join(G, "kats.json")))
    a = k["maxsim_1p7"]
    q = np.zeros((2, 32), np.float32); q[:, :4] = a["query"]
    d = np.zeros((3, 32), np.float32); d[:, :4] = a["doc"]
    assert abs(npb.maxsim_scores(q, [d])[0] - a["score"]) < a["tol"]
    gpu.close()
