"""Generate the committed fixtures under tests/golden/.

  kats.json ........... the reference's own known-answer vectors for this path (each with its
                        file:line in /root/reference/next-plaid/src) as data, so the oracle AND the CUDA
                        path are checked against the same literals
  search_small.npz .... a small seeded index in the reference's array layout, queries, and the results of
                        the CPU oracle for several parameter sets (ids + scores).  The reference (Rust)
                        cannot run here, so these are oracle outputs: they pin the oracle against
                        regressions and give the GPU tests a fixture that does not depend on numpy's RNG
                        or BLAS (the index arrays themselves are stored).

    python tools/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

PARAM_SETS = [
    dict(top_k=10, n_ivf_probe=8, n_full_scores=256, centroid_batch_size=100_000, centroid_score_threshold=0.4),
    dict(top_k=10, n_ivf_probe=8, n_full_scores=256, centroid_batch_size=64, centroid_score_threshold=0.4),
    dict(top_k=25, n_ivf_probe=4, n_full_scores=64, centroid_batch_size=100_000, centroid_score_threshold=None),
    dict(top_k=5, n_ivf_probe=16, n_full_scores=128, centroid_batch_size=0, centroid_score_threshold=0.3),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    kats = {
        "maxsim_1p7": {"ref": "maxsim.rs:393-413, search.rs:685-705",
                       "query": [[1, 0, 0, 0], [0, 1, 0, 0]],
                       "doc": [[0.5, 0.5, 0, 0], [0.8, 0.2, 0, 0], [0, 0.9, 0.1, 0]], "score": 1.7, "tol": 1e-5},
        "maxsim_nan_row_8p0": {"ref": "maxsim.rs:498-507", "n_query": 16, "query_row": [1.0, 0.0],
                               "doc_row": [0.5, 0.0], "nan_row": 15, "score": 8.0, "tol": 1e-5},
        "rerank_2_1_0": {"ref": "next-plaid-api/tests/integration_tests.rs:2301-2376",
                         "query": [[1, 0, 0, 0], [0, 1, 0, 0]],
                         "docs": [[[1, 0, 0, 0], [0, 1, 0, 0]], [[1, 0, 0, 0], [0, 0, 1, 0]], [[0, 0, 1, 0], [0, 0, 0, 1]]],
                         "scores": [2.0, 1.0, 0.0], "tol": 1e-2},
        "assign_to_centroids": {"ref": "maxsim.rs:444-477",
                                "centroids": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]],
                                "embeddings": [[0.9, 0.1, 0, 0], [0.1, 0.9, 0, 0], [0, 0.1, 0.9, 0], [0.8, 0.2, 0, 0], [0, 0, 0.8, 0.2]],
                                "codes": [0, 1, 2, 0, 2]},
        "compress_into_codes": {"ref": "codec.rs:637-663", "centroids": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]],
                                "embeddings": [[0.9, 0.1, 0, 0], [0, 0, 0.95, 0.05]], "codes": [0, 2]},
        "byte_layout": {"ref": "codec.rs:384-395, :168-214 (SURVEY 8a a7)", "nbits": 4, "buckets": [13, 10], "byte": 0xB5},
        "quantile": {"ref": "utils.rs:289-294", "values": [1, 2, 3, 4, 5], "q": [0.5, 0.0, 1.0], "expect": [3.0, 1.0, 5.0]},
        "find_outliers": {"ref": "update.rs:1170-1185", "centroids": [[0, 0], [1, 1]],
                          "embeddings": [[0.1, 0.1], [0.9, 0.9], [5.0, 5.0]], "threshold_sq": 1.0, "outliers": [2]},
        "defaults": {"ref": "search.rs:58-69, :708-715", "batch_size": 2000, "n_full_scores": 4096, "top_k": 10,
                     "n_ivf_probe": 8, "centroid_batch_size": 100000, "centroid_score_threshold": 0.4},
    }
    with open(os.path.join(OUT, "kats.json"), "w") as f:
        json.dump(kats, f, indent=1)

    docs = oracle.synthetic_corpus(400, 16, dim=64, seed=77, ragged=True)
    ix = oracle.create_index(docs, nbits=4, seed=7, num_partitions=128)
    qs, src = oracle.synthetic_queries(docs, 6, nq=16, seed=11)
    arrays = dict(centroids=ix.centroids, bucket_weights=ix.bucket_weights, bucket_cutoffs=ix.bucket_cutoffs,
                  codes=ix.codes.astype(np.int32), residuals=ix.residuals, doc_lengths=ix.doc_lengths.astype(np.int32),
                  ivf=ix.ivf.astype(np.int32), ivf_lengths=ix.ivf_lengths, nbits=np.int32(ix.nbits),
                  queries=np.stack(qs).astype(np.float32), source_docs=np.array(src, np.int32))
    for pi, kw in enumerate(PARAM_SETS):
        p = oracle.SearchParameters(**kw)
        ids = np.full((len(qs), p.top_k), -1, np.int32)
        sc = np.zeros((len(qs), p.top_k), np.float32)
        for i, q in enumerate(qs):
            r = oracle.search_one(ix, q, p)
            ids[i, :len(r.passage_ids)] = r.passage_ids
            sc[i, :len(r.scores)] = r.scores
        arrays[f"ids_{pi}"] = ids
        arrays[f"scores_{pi}"] = sc
    arrays["decompressed_doc7"] = oracle.get_document_embeddings(ix, 7)
    arrays["centroid_scores_q0"] = oracle.centroid_scores(qs[0], ix.centroids)
    np.savez_compressed(os.path.join(OUT, "search_small.npz"), **arrays)
    with open(os.path.join(OUT, "search_small.params.json"), "w") as f:
        json.dump(PARAM_SETS, f, indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
