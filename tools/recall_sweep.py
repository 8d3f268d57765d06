"""Sweep (n_ivf_probe, n_full_scores) on the bench workload: recall@top_k against exhaustive exact
MaxSim over the decompressed corpus, and device-timed queries/sec.  Run on a B200:
    python tools/recall_sweep.py --queries 64
Prints one JSON line per setting; the smallest setting with recall >= 0.99 is what bench.py uses."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import next_plaid_b200 as npb
    sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:]]
    extra = [a for a in sys.argv[1:]]
    nq_idx = extra.index("--queries") if "--queries" in extra else -1
    n_queries = int(extra[nq_idx + 1]) if nq_idx >= 0 else 64
    if nq_idx >= 0:
        del extra[nq_idx:nq_idx + 2]
    sys.argv = [sys.argv[0]] + extra
    args = bench.parse_args()
    dev = torch.device("cuda", 0)
    tens = bench.make_index_tensors(args, dev, 0)
    gpu = bench.open_index(npb, tens, args, 0, 0)
    queries, src = bench.make_queries(gpu, args, n_queries, seed=args.seed + 7)
    ex = np.concatenate([gpu.exhaustive_scores(queries[i:i + 32]) for i in range(0, n_queries, 32)], 0)
    truth = [set(np.lexsort((np.arange(ex.shape[1]), -ex[i]))[:args.top_k].tolist()) for i in range(n_queries)]
    planted = float(np.mean([int(np.argmax(ex[i]) == src[i]) for i in range(n_queries)]))
    print(json.dumps({"planted_doc_is_exhaustive_top1": planted, "queries": n_queries}))
    gpu.set_profiling(True)
    for thr in (args.threshold,):
        for n_probe in (2, 4, 8):
            for nfs in (1024, 2048, 4096, 8192):
                p = npb.SearchParameters(top_k=args.top_k, n_ivf_probe=n_probe, n_full_scores=nfs,
                                         centroid_score_threshold=thr)
                hits, ms, cand = [], 0.0, 0
                for i in range(0, n_queries, args.batch):
                    res = gpu.search_batch(queries[i:i + args.batch], p)
                    st, _ = gpu.last_stage_stats()
                    ms += sum(v for k, v in st.items() if k not in ("h2d", "d2h"))
                    cand += gpu.last_work_counters()["n_candidates"]
                    hits += [len(truth[i + j] & set(r.passage_ids.tolist())) / float(args.top_k) for j, r in enumerate(res)]
                print(json.dumps({"threshold": thr, "n_ivf_probe": n_probe, "n_full_scores": nfs,
                                  "recall": float(np.mean(hits)), "min_recall": float(np.min(hits)),
                                  "qps_device": n_queries / (ms * 1e-3), "cand_per_query": cand / n_queries}))


if __name__ == "__main__":
    main()
