"""N-rank NCCL check of the doc-sharded CUDA path (run under torchrun on N GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/gpu_sharded_check.py
Every rank opens its shard, joins the communicator and searches the same queries; the result on every
rank must be bit-identical to the CPU oracle searching the UNSHARDED index."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import next_plaid_b200 as npb
    from oracle import oracle
    import sharded_protocol as sp
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    docs = oracle.synthetic_corpus(4000, 40, dim=128, seed=31, ragged=True)
    ix = oracle.create_index(docs, nbits=4, seed=5, num_partitions=512)
    qs, _ = oracle.synthetic_queries(docs, 16, nq=32, seed=6)
    shard, base = sp.make_shard(oracle, ix, rank, world)
    gpu = npb.MmapIndex.from_arrays(shard.centroids, shard.bucket_weights, shard.codes, shard.residuals,
                                    shard.doc_lengths, shard.ivf, shard.ivf_lengths, shard.nbits, device=local,
                                    doc_id_base=base)
    uid = [npb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    gpu.comm_init(uid[0], rank, world)
    bad = 0
    for cbs, subset in ((100_000, None), (128, None), (128, list(range(0, 4000, 3)))):
        pg = npb.SearchParameters(top_k=10, n_ivf_probe=8, n_full_scores=256, centroid_batch_size=cbs)
        po = oracle.SearchParameters(top_k=10, n_ivf_probe=8, n_full_scores=256, centroid_batch_size=cbs)
        res = gpu.search_batch(qs, pg, subset=subset)
        for q, r in zip(qs, res):
            w = oracle.search_one(ix, q, po, subset=subset)
            if r.passage_ids.tolist() != w.passage_ids.tolist() or not np.array_equal(r.scores, w.scores):
                print(f"rank {rank}: search mismatch cbs={cbs} subset={subset is not None}", file=sys.stderr)
                bad += 1
    # data-parallel k-means over NCCL (pb_kmeans_fit_dp): every rank must end with the same unit-norm centroids
    rng = np.random.default_rng(3)
    centers = rng.standard_normal((32, 64)).astype(np.float32)
    centers /= np.linalg.norm(centers, axis=1, keepdims=True)
    prng = np.random.default_rng(100 + rank)
    pts = centers[prng.integers(0, 32, 6000)] + 0.05 * prng.standard_normal((6000, 64)).astype(np.float32)
    pts /= np.linalg.norm(pts, axis=1, keepdims=True)
    uid2 = [npb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid2, src=0)
    cent = npb.kmeans_fit_dp([pts], 32, niters=6, seed=5, device=local, nccl=(uid2[0], rank, world))
    allc = [None] * world
    dist.all_gather_object(allc, cent.tobytes())
    # (Lloyd from a random start need not find every blob: the single-GPU fit of the same seed is the yardstick)
    solo = npb.kmeans_fit(pts, 32, niters=6, seed=5, device=local)
    found = lambda c: ((centers @ c.T).max(1) > 0.98).mean()   # noqa: E731
    km_bad = any(c != allc[0] for c in allc) or np.abs(np.linalg.norm(cent, axis=1) - 1.0).max() > 1e-5 or \
        found(cent) < found(solo) - 0.2
    if km_bad:
        print(f"rank {rank}: k-means check failed: identical={all(c == allc[0] for c in allc)} found={found(cent):.2f} "
              f"solo={found(solo):.2f}", file=sys.stderr)
        bad += 1
    t = torch.tensor([bad], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"sharded check world={world}: {'OK' if t.item() == 0 else 'MISMATCH ' + str(t.item())}")
    gpu.close()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
