"""n_ivf_probe sweep on a bench-style corpus (BASELINE.json configs[4]: 2-bit residuals, 220-token docs, recall@1000,
n_ivf_probe 1..64): recall@top_k against exhaustive exact MaxSim over the decompressed corpus and device-timed
queries/sec per setting.  One GPU holds one shard of the corpus; under torchrun the corpus is doc-sharded as in bench.py.

    python tools/recall_sweep.py --queries 64 --docs-total 6250000 --doclen 220 --nbits 2 --top-k 1000 \
        --n-full-scores 8192 --batch 32 --probes 1,2,4,8,16,32,64                       # one shard of config E
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/recall_sweep.py --queries 64 --docs-total 50000000 --doclen 220 --nbits 2 --top-k 1000 ...

Prints one JSON line per setting (rank 0)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _pop(flag, default, cast):
    if flag in sys.argv:
        i = sys.argv.index(flag)
        v = cast(sys.argv[i + 1])
        del sys.argv[i:i + 2]
        return v
    return default


def main():
    import torch
    import torch.distributed as dist
    import next_plaid_b200 as npb
    n_queries = _pop("--queries", 64, int)
    probes = [int(x) for x in _pop("--probes", "1,2,4,8,16,32,64", str).split(",")]
    nfs_list = [int(x) for x in _pop("--nfs", "", str).split(",") if x]
    args = bench.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    G = bench.corpus_globals(args, dev)
    sh = bench.build_shard(args, G, rank, world, dev)
    per_rank = sh["D"]
    gpu = bench.open_shard(npb, args, G, sh, local, rank * per_rank)
    del sh["codes"]
    if world > 1:
        uid = [npb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.comm_init(uid[0], rank, world)
    queries = bench.make_queries(args, G, dev, n_queries, seed=args.seed + 7)
    # exhaustive ground truth: this shard's top_k per query, merged across shards
    top = []
    for i0 in range(0, n_queries, 16):
        ex = gpu.exhaustive_scores(queries[i0:i0 + 16])
        for i in range(ex.shape[0]):
            o = np.argpartition(-ex[i], min(args.top_k, ex.shape[1] - 1))[:args.top_k]
            top.append([(float(ex[i][j]), int(j) + rank * per_rank) for j in o])
        del ex
    if world > 1:
        alls = [None] * world
        dist.all_gather_object(alls, top)
        top = [[t for part in alls for t in part[i]] for i in range(n_queries)]
    truth = [set(t[1] for t in sorted(top[i], key=lambda t: (-t[0], t[1]))[:args.top_k]) for i in range(n_queries)]
    gpu.set_profiling(True)
    for nfs in (nfs_list or [args.n_full_scores]):
        for n_probe in probes:
            p = npb.SearchParameters(top_k=args.top_k, n_ivf_probe=n_probe, n_full_scores=nfs,
                                     centroid_score_threshold=args.threshold if args.threshold >= 0 else None)
            hits, ms, cand, kept = [], 0.0, 0, 0
            try:
                gpu.search_batch(queries[:args.batch], p)          # warm-up: workspaces of this setting
                for i in range(0, n_queries, args.batch):
                    res = gpu.search_batch(queries[i:i + args.batch], p)
                    ms += gpu.last_call_ms()
                    w = gpu.last_work_counters()
                    cand += w["n_candidates"]
                    kept += w["n_filter_docs"] or w["n_exact_docs"]
                    hits += [len(truth[i + j] & set(r.passage_ids.tolist())) / float(args.top_k) for j, r in enumerate(res)]
            except npb.PlaidError as e:
                if rank == 0:
                    print(json.dumps({"n_ivf_probe": n_probe, "n_full_scores": nfs, "error": str(e)}))
                continue
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0])
            if rank == 0:
                print(json.dumps({"n_gpus": world, "total_docs": args.docs_total, "doclen": args.doclen, "nbits": args.nbits,
                                  "top_k": args.top_k, "threshold": args.threshold if args.threshold >= 0 else None, "n_ivf_probe": n_probe,
                                  "n_full_scores": nfs, "recall_at_k": float(np.mean(hits)),
                                  "min_recall": float(np.min(hits)), "qps_device": n_queries / (ms * 1e-3),
                                  "candidates_per_query_per_gpu": cand / n_queries, "kept_per_query_per_gpu": kept / n_queries,
                                  "queries": n_queries}), flush=True)
    gpu.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
