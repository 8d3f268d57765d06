#!/usr/bin/env python
"""bench.py -- queries/sec of the PLAID search hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU algorithm

One "step" = one batch of `--batch` queries searched through the whole path (centroid scoring ->
probe -> candidates -> approximate score -> cut -> decompress + MaxSim -> top-k) against a synthetic
doc-sharded index resident in HBM.  At N=1 the workload is BASELINE.json configs[1]:
1M docs x 300 tok x 128-d, 4-bit residuals, K = 2^18, batch 32 queries x 32 tokens, top_k = 100.
N > 1: one rank per GPU, each rank owns a 1M-doc shard of an N-million-doc corpus (weak scaling,
configs[2] is 1.25M docs/GPU), queries replicated.

Rank 0 prints ONE JSON line:
  value     whole-job queries/sec with queries already in HBM, timed with CUDA events on the
            library's own stream (pb_last_stage_stats), max over ranks
  e2e       the same metric through the public API with HOST buffers: pinned-host queries in,
            host results out, H2D/D2H inside the timed region
  roofline  the dominant kernel by device time: algorithmic bytes / its CUDA-event time vs the
            measured HBM peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (restatement of the reference's Rust path) timed on this box's host
            cores on a bounded sample of the same queries and index; also the ids/scores parity check

Synthetic index (seed 42).  The corpus is generated directly in the codec domain -- centroid code +
packed residual per token, so a token IS normalise(C[code] + w[bucket]) (codec.rs:455-467) -- because
the f32 corpus (1M x 300 x 128 x 4 B = 154 GB) cannot exist and the GPU build path (k-means + encode)
is a later SURVEY 8 row.  Topic structure makes recall meaningful: every doc belongs to one of
D/256 topics, a topic owns a pool of 512 centroids, a token draws its code from the pool (80 %,
skewed) or uniformly (20 %).  A query is 32 tokens of one doc, each perturbed by 0.15 * unit noise.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=1_000_000, help="documents per GPU")
    ap.add_argument("--doclen", type=int, default=300)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--nbits", type=int, default=4)
    ap.add_argument("--log2k", type=int, default=18)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--nq", type=int, default=32)
    ap.add_argument("--top-k", type=int, default=100)
    ap.add_argument("--n-ivf-probe", type=int, default=8)
    ap.add_argument("--n-full-scores", type=int, default=4096)
    ap.add_argument("--threshold", type=float, default=0.4)
    ap.add_argument("--recall-queries", type=int, default=8)
    ap.add_argument("--cpu-queries", type=int, default=8, help="queries in the CPU-baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--threads", type=int, default=2, help="host threads for the extra concurrent-callers measurement")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--docs-per-topic", type=int, default=1024)
    ap.add_argument("--pool", type=int, default=256, help="centroids per topic pool")
    ap.add_argument("--res-sigma", type=float, default=0.05, help="per-dimension residual scale")
    ap.add_argument("--query-noise", type=float, default=0.15)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# synthetic index in the codec domain, built on the GPU with torch (harness, not product)
# ----------------------------------------------------------------------------------------------
def make_index_tensors(args, device, shard: int):
    import torch
    D, T, dim, nbits, K = args.docs, args.doclen, args.dim, args.nbits, 1 << args.log2k
    g = torch.Generator(device=device)
    g.manual_seed(args.seed)                       # shared across shards: centroids, pools, weights
    cent = torch.randn(K, dim, generator=g, device=device, dtype=torch.float32)
    cent /= cent.norm(dim=1, keepdim=True)
    n_topics = max((D * max(args.gpus, 1)) // args.docs_per_topic, 4)
    P = min(args.pool, K)
    pools = torch.randint(0, K, (n_topics, P), generator=g, device=device, dtype=torch.int32)
    hubs = torch.randint(0, K, (4096,), generator=g, device=device, dtype=torch.int64)   # stop-word-like centroids
    nb = 1 << nbits
    probs = (torch.arange(nb, dtype=torch.float64) + 0.5) / nb
    w = (args.res_sigma * torch.special.ndtri(probs)).to(torch.float32).to(device)   # quantile mid-points of N(0, s^2)
    g.manual_seed(args.seed + 1000 * (shard + 1))   # per-shard documents
    doc_topic = torch.randint(0, n_topics, (D,), generator=g, device=device, dtype=torch.int64)
    N = D * T
    codes = torch.empty(N, dtype=torch.int64, device=device)
    residuals = torch.empty((N, dim * nbits // 8), dtype=torch.uint8, device=device)
    chunk_docs = max(1, min(D, (1 << 25) // T))
    for d0 in range(0, D, chunk_docs):
        d1 = min(D, d0 + chunk_docs)
        n = (d1 - d0) * T
        topic = doc_topic[d0:d1].repeat_interleave(T)
        u = torch.rand(n, generator=g, device=device)
        pidx = (u * u * P).to(torch.int64).clamp_(max=P - 1)
        from_pool = pools[topic, pidx].to(torch.int64)
        rnd = torch.randint(0, K, (n,), generator=g, device=device, dtype=torch.int64)
        sel = torch.rand(n, generator=g, device=device)
        hub = hubs[torch.randint(0, 4096, (n,), generator=g, device=device)]
        codes[d0 * T:d1 * T] = torch.where(sel < 0.7, from_pool, torch.where(sel < 0.9, rnd, hub))
        del sel, hub
        residuals[d0 * T:d1 * T] = torch.randint(0, 256, (n, dim * nbits // 8), generator=g, device=device,
                                                 dtype=torch.uint8)
        del topic, u, pidx, from_pool, rnd
    doc_lengths = torch.full((D,), T, dtype=torch.int64, device=device)
    # IVF: per centroid the ascending unique doc ids (index.rs:479-499)
    keys = torch.empty(N, dtype=torch.int64, device=device)
    for d0 in range(0, D, chunk_docs):
        d1 = min(D, d0 + chunk_docs)
        doc = torch.arange(d0, d1, device=device, dtype=torch.int64).repeat_interleave(T)
        keys[d0 * T:d1 * T] = codes[d0 * T:d1 * T] * D + doc
        del doc
    keys = torch.unique(keys)           # sorted
    ivf = keys % D
    ivf_lengths = torch.bincount(keys // D, minlength=K).to(torch.int32)
    del keys
    torch.cuda.synchronize(device)
    return dict(centroids=cent, bucket_weights=w, codes=codes, residuals=residuals,
                doc_lengths=doc_lengths, ivf=ivf, ivf_lengths=ivf_lengths, K=K, D=D, N=N)


def open_index(npb, t, args, device_index: int, doc_id_base: int):
    return npb.MmapIndex.from_device_pointers(
        args.dim, args.nbits, t["K"], t["D"], t["N"], t["centroids"].data_ptr(), t["bucket_weights"].data_ptr(),
        t["codes"].data_ptr(), t["residuals"].data_ptr(), t["doc_lengths"].data_ptr(), t["ivf"].data_ptr(),
        t["ivf_lengths"].data_ptr(), device=device_index, doc_id_base=doc_id_base)


def make_queries(gpu, args, n_queries: int, seed: int):
    rng = np.random.default_rng(seed)
    D = gpu.num_documents()
    src = rng.integers(0, D, size=n_queries)
    out = []
    for d in src:
        emb = gpu.get_document_embeddings(int(d))
        tok = emb[rng.integers(0, emb.shape[0], size=args.nq)]
        noise = rng.standard_normal(tok.shape).astype(np.float32)
        noise /= np.linalg.norm(noise, axis=1, keepdims=True)
        q = tok + args.query_noise * noise
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        out.append(q.astype(np.float32))
    return out, src


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_index_from_tensors(oracle, t, args):
    return oracle.Index(t["centroids"].cpu().numpy(), t["bucket_weights"].cpu().numpy(), None,
                        t["codes"].cpu().numpy(), t["residuals"].cpu().numpy(), t["doc_lengths"].cpu().numpy(),
                        t["ivf"].cpu().numpy(), t["ivf_lengths"].cpu().numpy(), args.nbits)


def workload_config(args, world):
    return {"workload": f"{args.docs * world // 1000}k docs x {args.doclen} tok x {args.dim}-d, {args.nbits}-bit, "
                        f"K=2^{args.log2k}, batch {args.batch} x {args.nq} query tokens",
            "docs_per_gpu": args.docs, "total_docs": args.docs * world, "doclen": args.doclen, "dim": args.dim,
            "nbits": args.nbits, "num_centroids": 1 << args.log2k, "batch_queries": args.batch,
            "query_tokens": args.nq, "top_k": args.top_k, "n_ivf_probe": args.n_ivf_probe,
            "n_full_scores": args.n_full_scores, "centroid_score_threshold": args.threshold,
            "variant": "batched" if (1 << args.log2k) > 100_000 else "dense",
            "parallelism": f"doc-shard x{world}", "l2": "index (>= 20 GB/GPU) exceeds L2; distinct query batch per step"}


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import next_plaid_b200 as npb
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    t0 = time.time()
    tens = make_index_tensors(args, dev, rank)
    gpu = open_index(npb, tens, args, local, rank * args.docs)
    t_build = time.time() - t0
    if world > 1:   # doc-sharded: the library runs its own NCCL all-gathers; torch only ships the unique id
        uid = [npb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        gpu.comm_init(uid[0], rank, world)
    params = npb.SearchParameters(top_k=args.top_k, n_ivf_probe=args.n_ivf_probe, n_full_scores=args.n_full_scores,
                                  centroid_score_threshold=args.threshold)
    n_batches = max(args.steps + args.warmup, 4)
    n_batches = min(n_batches, 16)
    queries = make_queries(gpu, args, n_batches * args.batch, seed=args.seed + 7)[0] if rank == 0 else None
    if world > 1:   # the same queries on every rank, drawn from shard 0
        box = [queries if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        queries = box[0]
    batches = [queries[i * args.batch:(i + 1) * args.batch] for i in range(n_batches)]

    # ---- recall@top_k against exhaustive exact MaxSim over the decompressed corpus (untimed) ----
    rq = queries[:args.recall_queries]
    recall = None
    if rq:
        ex = gpu.exhaustive_scores(rq)                       # this shard's docs
        res = gpu.search_batch(rq, params)                   # collective when sharded
        top = []
        for i in range(len(rq)):
            o = np.lexsort((np.arange(ex.shape[1]), -ex[i]))[:args.top_k]
            top.append([(float(ex[i][j]), int(j) + rank * args.docs) for j in o])
        if world > 1:
            alls = [None] * world
            dist.all_gather_object(alls, top)
            top = [sorted((t for part in alls for t in part[i]), key=lambda t: (-t[0], t[1]))[:args.top_k]
                   for i in range(len(rq))]
        hits = [len({t[1] for t in top[i]} & set(r.passage_ids.tolist())) / float(args.top_k)
                for i, r in enumerate(res)]
        recall = float(np.mean(hits))
        del ex

    # ---- device-resident timing ("value"): CUDA events on the library's stream ----
    flat = [np.concatenate(b, 0) for b in batches]
    offs = np.arange(args.batch + 1, dtype=np.int64) * args.nq
    d_q = [torch.from_numpy(f).to(dev) for f in flat]
    d_ids = torch.empty((args.batch, args.top_k), dtype=torch.int64, device=dev)
    d_sc = torch.empty((args.batch, args.top_k), dtype=torch.float32, device=dev)
    d_cn = torch.empty((args.batch,), dtype=torch.int32, device=dev)
    gpu.set_profiling(True)
    sampler = ClockSampler(local)       # spans warm-up + both timed regions (nvidia-smi needs ~0.2 s to start)
    for i in range(args.warmup):
        gpu.search_batch_device(d_q[i % n_batches].data_ptr(), offs, params, d_ids.data_ptr(), d_sc.data_ptr(),
                                d_cn.data_ptr())
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    stage_ms = {}
    launches = 0
    work = {}
    dev_ms = 0.0
    for i in range(args.steps):
        gpu.search_batch_device(d_q[(args.warmup + i) % n_batches].data_ptr(), offs, params, d_ids.data_ptr(),
                                d_sc.data_ptr(), d_cn.data_ptr())
        ms, ln = gpu.last_stage_stats()
        for k, v in ms.items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
        dev_ms += sum(ms.values())
        launches += sum(ln.values())
        for k, v in gpu.last_work_counters().items():
            work[k] = work.get(k, 0) + v
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    gpu.set_profiling(False)

    # ---- end-to-end through the public API: pinned host queries in, host results out ----
    pinned = [torch.from_numpy(f).pin_memory() for f in flat]
    import ctypes as C
    idx = sys.modules["next_plaid_b200.index"]
    L = npb.load_library()
    h_ids = np.zeros((args.batch, args.top_k), np.int64)
    h_sc = np.zeros((args.batch, args.top_k), np.float32)
    h_cn = np.zeros(args.batch, np.int32)
    pc = params._c()

    def e2e_step(i):
        st = L.pb_search_batch(gpu._h, C.c_void_p(pinned[i % n_batches].data_ptr()), offs.ctypes.data_as(C.c_void_p),
                               args.batch, C.byref(pc), None, 0, h_ids.ctypes.data_as(C.c_void_p),
                               h_sc.ctypes.data_as(C.c_void_p), h_cn.ctypes.data_as(C.c_void_p))
        if st != 0:
            raise RuntimeError(L.pb_last_error().decode())
    for i in range(args.warmup):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(args.warmup + i)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t1
    clocks = sampler.stop()

    # ---- the reference's deployment model: several host threads share one index (state.rs:24-47).
    #      Same K steps, issued from T threads; reported next to the headline, not instead of it ----
    concurrent = None
    if world == 1 and args.threads > 1:
        def worker(tid):
            bufs = (np.zeros_like(h_ids), np.zeros_like(h_sc), np.zeros_like(h_cn))
            for i in range(tid, args.steps, args.threads):
                st = L.pb_search_batch(gpu._h, C.c_void_p(pinned[(args.warmup + i) % n_batches].data_ptr()),
                                       offs.ctypes.data_as(C.c_void_p), args.batch, C.byref(pc), None, 0,
                                       bufs[0].ctypes.data_as(C.c_void_p), bufs[1].ctypes.data_as(C.c_void_p),
                                       bufs[2].ctypes.data_as(C.c_void_p))
                if st != 0:
                    raise RuntimeError(L.pb_last_error().decode())
        for rep in range(2):          # first repetition creates the extra workspaces
            ths = [threading.Thread(target=worker, args=(t,)) for t in range(args.threads)]
            torch.cuda.synchronize(dev)
            tc0 = time.perf_counter()
            [t.start() for t in ths]
            [t.join() for t in ths]
            torch.cuda.synchronize(dev)
            tc = time.perf_counter() - tc0
        concurrent = {"host_threads": args.threads, "value": args.batch * args.steps / tc, "unit": "queries/s",
                      "ms_per_step": 1e3 * tc / args.steps,
                      "note": "same steps through pb_search_batch (host buffers) from T threads on one handle"}

    # ---- CPU baseline + parity on a bounded sample ----
    cpu = None
    parity = None
    if not args.no_cpu and args.cpu_queries > 0 and world == 1:
        from oracle import oracle
        hix = host_index_from_tensors(oracle, tens, args)
        po = oracle.SearchParameters(top_k=args.top_k, n_ivf_probe=args.n_ivf_probe,
                                     n_full_scores=args.n_full_scores, centroid_score_threshold=args.threshold)
        sample = batches[0][:args.cpu_queries]
        oracle.search_one(hix, sample[0], po)          # warm the page cache / thread pool
        c0 = time.perf_counter()
        cres = [oracle.search_one(hix, q, po) for q in sample]
        c_s = time.perf_counter() - c0
        gres = gpu.search_batch(sample, params)
        same_ids = sum(int(g.passage_ids.tolist() == c.passage_ids.tolist()) for g, c in zip(gres, cres))
        max_ds = max((float(np.abs(g.scores - c.scores).max()) if len(g.scores) == len(c.scores) and len(g.scores)
                      else (0.0 if len(g.scores) == len(c.scores) else float("inf"))) for g, c in zip(gres, cres))
        parity = {"queries": len(sample), "ids_identical": same_ids, "max_abs_score_diff": max_ds}
        cpu = {"value": len(sample) / c_s, "unit": "queries/s", "cores": oracle.lib().po_num_threads(),
               "kind": "port", "sample": f"{len(sample)} queries of batch 0, same index and parameters, "
                                         f"{c_s:.1f} s of wall time (C restatement of the reference, OpenMP)"}
        del hix

    # ---- roofline: every stage's algorithmic traffic / CUDA-event time; the headline block is the
    #      stage with the largest share of the step (DESIGN.md section 4 states the per-unit bytes) ----
    peak, peak_src = measured_peaks()
    kern_stages = {k: v for k, v in stage_ms.items() if k not in ("h2d", "d2h")}
    nq_tot = work.get("n_query_tokens", 0)
    K = 1 << args.log2k
    packed = args.dim * args.nbits // 8
    alg = {
        # C read once per launch, fp32 S written once, 16-bit copy written once
        "centroid_scores": args.steps * K * args.dim * 4 + nq_tot * K * 6,
        "probe": 2 * nq_tot * K * 2,                               # 16-bit S streamed twice (chunk maxima, collect)
        # one u32 code per (candidate, distinct code) + each 16-bit S entry once
        "approx": work.get("n_candidate_tokens", 0) * 4 + nq_tot * K * 2,
        # packed residual + code per token: once for every kept doc in the tensor-core filter (k_exact_tc), once
        # more for the survivors the fp32 kernel scores (k_exact)
        "exact": (work.get("n_exact_tokens", 0) + work.get("n_filter_tokens", 0)) * (packed + 4),
        "cut": work.get("n_candidates", 0) * 8, "candidates": 0, "topk": 0,
    }
    flops = {"centroid_scores": 2.0 * nq_tot * K * args.dim,
             "exact": 2.0 * work.get("n_exact_tokens", 0) * args.nq * args.dim}
    names = {"approx": "k_approx16 (+k_select_u32, k_approx re-check)", "exact": "k_exact_tc (tcgen05 filter) + k_exact (survivors)", "centroid_scores": "k_centroid_scores",
             "probe": "k_chunkmax16 + k_collect16 (+k_tau16, k_topn_merge, k_cells)", "cut": "k_cut", "candidates": "k_mark/k_compact", "topk": "k_topk"}
    per_stage = {}
    for st, ms_tot in kern_stages.items():
        ms1 = ms_tot / max(args.steps, 1)
        gbs = alg.get(st, 0) / max(args.steps, 1) / (ms1 * 1e-3) / 1e9 if ms1 > 0 else 0.0
        ent = {"kernel": names.get(st, st), "ms_per_step": ms1, "algorithmic_bytes_per_step": alg.get(st, 0) / max(args.steps, 1),
               "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        if st in flops and ms1 > 0:
            ent["fp32_tflops"] = flops[st] / max(args.steps, 1) / (ms1 * 1e-3) / 1e12
            ent["frac_of_fp32_fma_peak"] = ent["fp32_tflops"] / 74.4   # 148 SMs x 128 FMA x 2 x 1.965 GHz
        per_stage[st] = ent
    # dram__bytes_read.sum + dram__bytes_write.sum of the stage's main kernel, one launch, from the committed
    # `ncu --set full` captures of this exact workload (profiles/r01_ncu_full_top2_raw.csv for k_approx16 and
    # k_exact_tc, r01_ncu_full_top4_raw.csv for k_centroid_scores); null for other shapes
    ncu_traffic = {"approx": 2.269352e9 + 0.012187e9, "exact": 0.966227e9 + 0.008221e9,
                   "centroid_scores": 0.138181e9 + 1.555811e9}
    default_shape = (args.docs, args.doclen, args.dim, args.nbits, args.log2k, args.batch, args.nq, args.top_k,
                     args.n_ivf_probe, args.n_full_scores) == (1_000_000, 300, 128, 4, 18, 32, 32, 100, 8, 4096)
    dom = max(kern_stages, key=kern_stages.get)
    d = per_stage[dom]
    roof = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": d["frac_of_hbm_peak"], "traffic": ncu_traffic.get(dom) if (default_shape and world == 1) else None,
            "peak_source": peak_src, "ms_per_launch": d["ms_per_step"],
            "algorithmic_bytes_per_launch": d["algorithmic_bytes_per_step"],
            "note": "k_approx16 gathers 64-byte rows of the 16-bit score table out of L2 (86 % hit rate): it is bound by "
                    "the L1 tag / request path (l1tex 79 % busy, profiles/r01_summary.md), not by HBM bytes; "
                    "k_exact and k_centroid_scores are fp32-FMA bound by design (pinned accumulation order, "
                    "DESIGN.md Numerics): see fp32_tflops in roofline_all",
            "fp32_tflops": d.get("fp32_tflops"), "frac_of_fp32_fma_peak": d.get("frac_of_fp32_fma_peak")}

    if world > 1:
        tt = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(tt[0]), float(tt[1])
    qps = args.batch * args.steps / (dev_ms * 1e-3)
    out = {
        "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (codec-domain index, seed 42)",
        "config": workload_config(args, world),
        "recall_at_k": recall, "recall_queries": len(rq),
        "e2e": {"value": args.batch * args.steps / e2e_s, "unit": "queries/s",
                "h2d_bytes_per_step": int(flat[0].nbytes + offs.nbytes),
                "d2h_bytes_per_step": int(h_ids.nbytes + h_sc.nbytes + h_cn.nbytes), "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_all": per_stage,
        "cpu_baseline": cpu, "parity": parity, "concurrent": concurrent,
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
        "work_per_step": {k: v / args.steps for k, v in work.items()},
        "index_build_s": t_build,
    }
    if rank == 0:
        print(json.dumps(out))
    gpu.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation of the path.  The reference is Rust and this image has
    no cargo/rustc, so oracle/_ref cannot exist; the timed code is the C restatement (oracle/), on all
    host threads, same index generator, same queries and parameters as the b200 arm.  With N > 1 the
    corpus is the same N x docs-per-GPU corpus, held in host memory shard by shard."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    import torch
    import next_plaid_b200 as npb
    from oracle import oracle
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    shards, bases, queries = [], [], None
    for g in range(world):
        tens = make_index_tensors(args, dev, g)
        if g == 0:  # the b200 arm draws its queries from shard 0
            gpu = open_index(npb, tens, args, dev.index or 0, 0)
            queries, _ = make_queries(gpu, args, min(max(args.steps + args.warmup, 4), 16) * args.batch, seed=args.seed + 7)
            gpu.close()
        shards.append(host_index_from_tensors(oracle, tens, args))
        bases.append(g * args.docs)
        del tens
        torch.cuda.empty_cache()
    po = oracle.SearchParameters(top_k=args.top_k, n_ivf_probe=args.n_ivf_probe, n_full_scores=args.n_full_scores,
                                 centroid_score_threshold=args.threshold)
    per_step = max(1, args.cpu_queries // world)      # bounded sample: CPU work per query grows with the corpus
    step_q = lambda i: queries[(i * args.batch) % len(queries):][:per_step]   # noqa: E731
    for i in range(args.warmup):
        for q in step_q(i):
            oracle.search_sharded(shards, bases, q, po)
    t0 = time.perf_counter()
    for i in range(args.steps):
        for q in step_q(args.warmup + i):
            oracle.search_sharded(shards, bases, q, po)
    s = time.perf_counter() - t0
    qps = per_step * args.steps / s
    cores = oracle.lib().po_num_threads()
    sample = f"{per_step} queries per step (first of each {args.batch}-query batch), {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (codec-domain index, seed 42)",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
