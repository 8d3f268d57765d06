"""world_size-2 gloo test of the doc-sharded protocol: sharded search == unsharded search,
bit-identical ids and scores (the property the NCCL path must also satisfy)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    from oracle import oracle
    import sharded_protocol as sp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    docs = oracle.synthetic_corpus(500, 20, dim=32, seed=8, ragged=True)
    ix = oracle.create_index(docs, nbits=4, seed=3, num_partitions=64)
    qs, _ = oracle.synthetic_queries(docs, 5, nq=8, seed=2)
    shard, base = sp.make_shard(oracle, ix, rank, world)
    ok = True
    for cbs, subset in ((100_000, None), (16, None), (16, list(range(0, 500, 3)))):
        p = oracle.SearchParameters(top_k=10, n_ivf_probe=4, n_full_scores=64, centroid_batch_size=cbs)
        for q in qs:
            ids, sc = sp.sharded_search_one(oracle, dist, shard, base, q, p, subset)
            want = oracle.search_one(ix, q, p, subset=subset)
            ok &= ids.tolist() == want.passage_ids.tolist() and np.array_equal(sc, want.scores)
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else "MISMATCH")
    dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
