"""Python face of the CPU oracle (ctypes over oracle/libplaid_oracle.so) plus the numpy-side
restatement of the reference's index builder and on-disk format.

TEST INFRASTRUCTURE ONLY -- see the header of oracle/plaid_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.

Reference lines restated here (relative to /root/reference/next-plaid/src):
  index directory layout ......... index.rs:373-528 (write_index_from_encoded_chunks)
  MmapIndex::load ................ index.rs:1026-1139 (offsets :1089-1110)
  codec training ................. index.rs:182-287 (prepare_codec_artifacts)
  chunk encode ................... index.rs:289-371, :17-40
  K / sampling heuristics ........ kmeans.rs:273-312
  centroid normalisation ......... kmeans.rs:415-419

Parity status of the build path: the k-means iteration lives in fastkmeans-rs 1.0.8 and the
sample shuffles in rand_chacha 0.3.1, neither of which is in /root/reference -> centroid values and
sample membership are "parity unpinned" (SURVEY.md 8c).  Search parity is unaffected: both sides
read the same index files.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libplaid_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/plaid_oracle.c (gcc, recipe in oracle/Makefile)."""
    src = os.path.join(_HERE, "plaid_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


class _POIndex(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("nbits", C.c_int32),
        ("num_centroids", C.c_int64), ("num_docs", C.c_int64),
        ("centroids", C.c_void_p), ("bucket_weights", C.c_void_p),
        ("codes", C.c_void_p), ("residuals", C.c_void_p),
        ("doc_offsets", C.c_void_p), ("ivf", C.c_void_p),
        ("ivf_lengths", C.c_void_p), ("ivf_offsets", C.c_void_p),
    ]


class _POParams(C.Structure):
    _fields_ = [
        ("n_full_scores", C.c_int64), ("top_k", C.c_int64), ("n_ivf_probe", C.c_int64),
        ("centroid_batch_size", C.c_int64), ("has_threshold", C.c_int32),
        ("centroid_score_threshold", C.c_float),
    ]


class _POTrace(C.Structure):
    _fields_ = [
        ("cells", C.c_void_p), ("cells_cap", C.c_int64), ("n_cells", C.c_int64),
        ("candidates", C.c_void_p), ("approx", C.c_void_p), ("cand_cap", C.c_int64),
        ("n_candidates", C.c_int64),
        ("kept", C.c_void_p), ("kept_exact", C.c_void_p), ("kept_cap", C.c_int64),
        ("n_kept", C.c_int64), ("variant", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.po_cmp_score_ascending.argtypes = [C.c_float, C.c_float]
        L.po_cmp_score_ascending.restype = C.c_int
        L.po_is_score_better.argtypes = [C.c_float, C.c_float]
        L.po_is_score_better.restype = C.c_int
        L.po_max_score.argtypes = [C.c_float, C.c_float]
        L.po_max_score.restype = C.c_float
        L.po_dot.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.po_dot.restype = C.c_float
        L.po_sumsq.argtypes = [C.c_void_p, C.c_int]
        L.po_sumsq.restype = C.c_float
        L.po_maxsim.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int]
        L.po_maxsim.restype = C.c_float
        L.po_search_one.restype = C.c_int64
        L.po_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------------------
# reference-shaped parameter / result types (search.rs:27-80)
# --------------------------------------------------------------------------------------

@dataclass
class SearchParameters:
    """search.rs:27-69 (defaults :58-69)."""
    batch_size: int = 2000
    n_full_scores: int = 4096
    top_k: int = 10
    n_ivf_probe: int = 8
    centroid_batch_size: int = 100_000
    centroid_score_threshold: Optional[float] = 0.4

    def _c(self) -> _POParams:
        t = self.centroid_score_threshold
        return _POParams(self.n_full_scores, self.top_k, self.n_ivf_probe,
                         self.centroid_batch_size, 0 if t is None else 1,
                         0.0 if t is None else float(t))


@dataclass
class QueryResult:
    """search.rs:72-80."""
    query_id: int
    passage_ids: np.ndarray
    scores: np.ndarray


@dataclass
class Trace:
    variant: int
    cells: np.ndarray
    candidates: np.ndarray
    approx: np.ndarray
    kept: np.ndarray
    kept_exact: np.ndarray


# --------------------------------------------------------------------------------------
# index in host memory, reference layout (index.rs:995-1016)
# --------------------------------------------------------------------------------------

@dataclass
class Index:
    centroids: np.ndarray        # f32 [K, dim]
    bucket_weights: np.ndarray   # f32 [2^nbits]
    bucket_cutoffs: Optional[np.ndarray]  # f32 [2^nbits - 1]
    codes: np.ndarray            # i64 [N]
    residuals: np.ndarray        # u8  [N, dim*nbits/8]
    doc_lengths: np.ndarray      # i64 [D]
    ivf: np.ndarray              # i64 [sum len]
    ivf_lengths: np.ndarray      # i32 [K]
    nbits: int
    doc_offsets: np.ndarray = field(default=None)   # i64 [D+1]
    ivf_offsets: np.ndarray = field(default=None)   # i64 [K+1]

    def __post_init__(self):
        self.centroids = _f32(self.centroids)
        self.bucket_weights = _f32(self.bucket_weights)
        self.codes = np.ascontiguousarray(self.codes, dtype=np.int64)
        self.residuals = np.ascontiguousarray(self.residuals, dtype=np.uint8)
        self.doc_lengths = np.ascontiguousarray(self.doc_lengths, dtype=np.int64)
        self.ivf = np.ascontiguousarray(self.ivf, dtype=np.int64)
        self.ivf_lengths = np.ascontiguousarray(self.ivf_lengths, dtype=np.int32)
        self.doc_offsets = np.zeros(len(self.doc_lengths) + 1, dtype=np.int64)
        np.cumsum(self.doc_lengths, out=self.doc_offsets[1:])          # index.rs:1107-1110
        self.ivf_offsets = np.zeros(len(self.ivf_lengths) + 1, dtype=np.int64)
        np.cumsum(self.ivf_lengths, out=self.ivf_offsets[1:])          # index.rs:1089-1094

    @property
    def dim(self) -> int:
        return int(self.centroids.shape[1])

    @property
    def num_centroids(self) -> int:
        return int(self.centroids.shape[0])

    @property
    def num_documents(self) -> int:
        return int(len(self.doc_lengths))

    @property
    def num_embeddings(self) -> int:
        return int(self.doc_offsets[-1])

    def _c(self) -> _POIndex:
        return _POIndex(self.dim, self.nbits, self.num_centroids, self.num_documents,
                        _p(self.centroids), _p(self.bucket_weights), _p(self.codes),
                        _p(self.residuals), _p(self.doc_offsets), _p(self.ivf),
                        _p(self.ivf_lengths), _p(self.ivf_offsets))


# --------------------------------------------------------------------------------------
# search
# --------------------------------------------------------------------------------------

def search_one(index: Index, query: np.ndarray, params: SearchParameters,
               subset: Optional[Sequence[int]] = None, trace: bool = False):
    """MmapIndex::search -> search_one_mmap (index.rs:1258, search.rs:327)."""
    L = lib()
    q = _f32(query).reshape(-1, index.dim)
    k = max(int(params.top_k), 1)
    ids = np.zeros(k, dtype=np.int64)
    sc = np.zeros(k, dtype=np.float32)
    ss = None if subset is None else np.ascontiguousarray(subset, dtype=np.int64)
    cix, cp = index._c(), params._c()
    tr = None
    if trace:
        D = index.num_documents
        K = index.num_centroids
        n_keep = max(params.n_full_scores // 4, params.top_k)
        t_cells = np.zeros(K, np.int64)
        t_cand = np.zeros(D, np.int64)
        t_appr = np.zeros(D, np.float32)
        t_kept = np.zeros(max(n_keep, 1), np.int64)
        t_kex = np.zeros(max(n_keep, 1), np.float32)
        tr = _POTrace(_p(t_cells), K, 0, _p(t_cand), _p(t_appr), D, 0, _p(t_kept), _p(t_kex),
                      max(n_keep, 1), 0, 0)
    n = L.po_search_one(C.byref(cix), _p(q), C.c_int(q.shape[0]), C.byref(cp),
                        None if ss is None else _p(ss),
                        C.c_int64(0 if ss is None else len(ss)), C.c_int(0 if ss is None else 1),
                        _p(ids), _p(sc), None if tr is None else C.byref(tr))
    if n < 0:
        raise ValueError("po_search_one rejected its arguments")
    res = QueryResult(0, ids[:n].copy(), sc[:n].copy())
    if trace:
        return res, Trace(tr.variant, t_cells[:tr.n_cells].copy(),
                          t_cand[:tr.n_candidates].copy(), t_appr[:tr.n_candidates].copy(),
                          t_kept[:tr.n_kept].copy(), t_kex[:tr.n_kept].copy())
    return res


def search_sharded(shards: Sequence[Index], bases: Sequence[int], query: np.ndarray, params: SearchParameters):
    """po_search_sharded: the reference search over a corpus held as doc-contiguous shards sharing the
    centroids; equal to search_one on the concatenated index."""
    L = lib()
    L.po_search_sharded.restype = C.c_int64
    q = _f32(query).reshape(-1, shards[0].dim)
    k = max(int(params.top_k), 1)
    ids = np.zeros(k, dtype=np.int64)
    sc = np.zeros(k, dtype=np.float32)
    cs = [s._c() for s in shards]
    arr = (C.POINTER(_POIndex) * len(cs))(*[C.pointer(c) for c in cs])
    b = np.ascontiguousarray(bases, dtype=np.int64)
    cp = params._c()
    n = L.po_search_sharded(arr, _p(b), C.c_int(len(cs)), _p(q), C.c_int(q.shape[0]), C.byref(cp), _p(ids), _p(sc))
    if n < 0:
        raise ValueError("po_search_sharded rejected its arguments")
    return QueryResult(0, ids[:n].copy(), sc[:n].copy())


def search_batch(index: Index, queries: Sequence[np.ndarray], params: SearchParameters,
                 subset: Optional[Sequence[int]] = None) -> List[QueryResult]:
    """MmapIndex::search_batch -> search_many_mmap (index.rs:1279, search.rs:643)."""
    out = []
    for i, q in enumerate(queries):
        r = search_one(index, q, params, subset)
        r.query_id = i                                      # search.rs:661
        out.append(r)
    return out


def exhaustive_scores(index: Index, query: np.ndarray, d0: int = 0, d1: Optional[int] = None):
    """Exact MaxSim of `query` against every doc's decompressed embedding (recall ground truth)."""
    d1 = index.num_documents if d1 is None else d1
    q = _f32(query).reshape(-1, index.dim)
    out = np.zeros(d1 - d0, np.float32)
    cix = index._c()
    lib().po_exhaustive_scores(C.byref(cix), _p(q), C.c_int(q.shape[0]), C.c_int64(d0),
                               C.c_int64(d1), _p(out))
    return out


# --------------------------------------------------------------------------------------
# codec / maxsim primitives
# --------------------------------------------------------------------------------------

def centroid_scores(query: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    q, c = _f32(query), _f32(centroids)
    S = np.zeros((q.shape[0], c.shape[0]), np.float32)
    lib().po_centroid_scores(_p(q), C.c_int(q.shape[0]), _p(c), C.c_int64(c.shape[0]),
                             C.c_int(c.shape[1]), _p(S))
    return S


def maxsim_score(query: np.ndarray, doc: np.ndarray) -> float:
    """maxsim.rs:270."""
    q, d = _f32(query), _f32(doc)
    return float(lib().po_maxsim(_p(q), q.shape[0], _p(d), d.shape[0], q.shape[1]))


def decompress(centroids, bucket_weights, nbits: int, packed: np.ndarray, codes: np.ndarray):
    """ResidualCodec::decompress, codec.rs:423."""
    c, w = _f32(centroids), _f32(bucket_weights)
    pk = np.ascontiguousarray(packed, dtype=np.uint8)
    cd = np.ascontiguousarray(codes, dtype=np.int64)
    out = np.zeros((len(cd), c.shape[1]), np.float32)
    lib().po_decompress(_p(c), C.c_int(c.shape[1]), C.c_int(nbits), _p(w), _p(pk), _p(cd),
                        C.c_int64(len(cd)), _p(out))
    return out


def get_document_embeddings(index: Index, doc_id: int) -> np.ndarray:
    """index.rs:1159-1179."""
    s, e = int(index.doc_offsets[doc_id]), int(index.doc_offsets[doc_id + 1])
    return decompress(index.centroids, index.bucket_weights, index.nbits,
                      index.residuals[s:e], index.codes[s:e])


def quantize_residuals(residuals: np.ndarray, cutoffs: np.ndarray, nbits: int) -> np.ndarray:
    """codec.rs:356."""
    r, ct = _f32(residuals), _f32(cutoffs)
    n, dim = r.shape
    out = np.zeros((n, dim * nbits // 8), np.uint8)
    lib().po_quantize_residuals(_p(r), C.c_int64(n), C.c_int(dim), C.c_int(nbits), _p(ct), _p(out))
    return out


def compress_into_codes(embeddings: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """codec.rs:297."""
    e, c = _f32(embeddings), _f32(centroids)
    out = np.zeros(e.shape[0], np.int64)
    lib().po_compress_into_codes(_p(e), C.c_int64(e.shape[0]), _p(c), C.c_int64(c.shape[0]),
                                 C.c_int(c.shape[1]), _p(out))
    return out


def residuals_of(embeddings: np.ndarray, centroids: np.ndarray, codes: np.ndarray) -> np.ndarray:
    """index.rs:17-40."""
    e, c = _f32(embeddings), _f32(centroids)
    cd = np.ascontiguousarray(codes, dtype=np.int64)
    out = np.zeros_like(e)
    lib().po_residuals(_p(e), C.c_int64(e.shape[0]), _p(c), C.c_int(c.shape[1]), _p(cd), _p(out))
    return out


def quantiles(arr: np.ndarray, qs: Sequence[float]) -> np.ndarray:
    """utils.rs:125."""
    a = _f32(arr).ravel()
    q = np.ascontiguousarray(qs, dtype=np.float64)
    out = np.zeros(len(q), np.float32)
    lib().po_quantiles(_p(a), C.c_int64(len(a)), _p(q), C.c_int(len(q)), _p(out))
    return out


def find_outliers(embeddings: np.ndarray, centroids: np.ndarray, threshold_sq: float) -> np.ndarray:
    """update.rs:490-608: rows whose min squared L2 distance to any centroid exceeds threshold_sq."""
    e, c = _f32(embeddings), _f32(centroids)
    out = np.zeros(max(e.shape[0], 1), np.int64)
    L = lib()
    L.po_find_outliers.restype = C.c_int64
    m = L.po_find_outliers(_p(e), C.c_int64(e.shape[0]), _p(c), C.c_int64(c.shape[0]), C.c_int(c.shape[1]),
                           C.c_float(threshold_sq), _p(out))
    return out[:m].copy()


def byte_reversed_bits_map(nbits: int) -> np.ndarray:
    out = np.zeros(256, np.uint8)
    lib().po_byte_reversed_bits_map(C.c_int(nbits), _p(out))
    return out


def bucket_index_lookup(nbits: int) -> np.ndarray:
    out = np.zeros((256, 8 // nbits), np.uint8)
    lib().po_bucket_index_lookup(C.c_int(nbits), _p(out))
    return out


# --------------------------------------------------------------------------------------
# index build (secondary path; numpy restatement, k-means unpinned)
# --------------------------------------------------------------------------------------

def num_partitions_heuristic(num_documents: int, sample_doclens: Sequence[int]) -> int:
    """kmeans.rs:304-309: K = 2^floor(log2(16*sqrt(avg_sample_doclen * D)))."""
    avg = float(np.sum(sample_doclens)) / max(len(sample_doclens), 1)
    est = avg * num_documents
    return int(2 ** int(np.floor(np.log2(16.0 * np.sqrt(est)))))


def kmeans(samples: np.ndarray, k: int, niters: int = 4, seed: int = 42,
           max_points_per_centroid: int = 256) -> np.ndarray:
    """Lloyd iterations in the shape fastkmeans documents (subsample to k*max_points, random
    data-point init, inner-product-free L2 assignment, mean update, empty clusters keep their
    previous centroid) followed by the L2 normalisation of kmeans.rs:415-419.  PARITY UNPINNED."""
    rng = np.random.default_rng(seed)
    x = _f32(samples)
    if x.shape[0] > k * max_points_per_centroid:
        x = x[rng.choice(x.shape[0], k * max_points_per_centroid, replace=False)]
    k = min(k, x.shape[0])
    cent = x[rng.choice(x.shape[0], k, replace=False)].copy()
    for _ in range(niters):
        d = (x * x).sum(1, keepdims=True) + (cent * cent).sum(1)[None, :] - 2.0 * (x @ cent.T)
        a = d.argmin(1)
        sums = np.zeros_like(cent)
        np.add.at(sums, a, x)
        cnt = np.bincount(a, minlength=k).astype(np.float32)
        nz = cnt > 0
        cent[nz] = sums[nz] / cnt[nz, None]
    n = np.maximum(np.sqrt((cent * cent).sum(1, keepdims=True)), 1e-12)
    return (cent / n).astype(np.float32)


@dataclass
class CodecArtifacts:
    centroids: np.ndarray
    bucket_cutoffs: np.ndarray
    bucket_weights: np.ndarray
    avg_residual: np.ndarray
    cluster_threshold: float


def prepare_codec_artifacts(embeddings: Sequence[np.ndarray], centroids: np.ndarray, nbits: int,
                            seed: int = 42) -> CodecArtifacts:
    """index.rs:182-287.  Sample membership is unpinned (ChaCha8 shuffle not reproduced)."""
    D = len(embeddings)
    total = int(sum(e.shape[0] for e in embeddings))
    sample_count = max(min(int(16.0 * np.sqrt(120.0 * D)), D), 1)           # :195-197
    rng = np.random.default_rng(seed)
    sample = rng.permutation(D)[:sample_count]
    heldout_size = int(min(0.05 * total, 50000.0))                          # :212
    rows, collected = [], 0
    for idx in sample[::-1]:                                                # :216-226
        if collected >= heldout_size:
            break
        e = embeddings[int(idx)]
        take = min(heldout_size - collected, e.shape[0])
        rows.append(e[:take])
        collected += take
    dim = centroids.shape[1]
    heldout = _f32(np.concatenate(rows, 0)) if rows else np.zeros((0, dim), np.float32)
    codes = compress_into_codes(heldout, centroids)
    res = residuals_of(heldout, centroids, codes)
    dist = np.sqrt((res * res).sum(1)).astype(np.float32)
    thr = float(quantiles(dist, [0.75])[0]) if len(dist) else 0.0           # :249-253
    avg_res = np.abs(res).mean(0).astype(np.float32) if len(res) else np.zeros(dim, np.float32)
    n_opt = 1 << nbits
    cut = quantiles(res.ravel(), [i / n_opt for i in range(1, n_opt)])        # :260-270
    wts = quantiles(res.ravel(), [(i + 0.5) / n_opt for i in range(n_opt)])
    return CodecArtifacts(_f32(centroids), cut, wts, avg_res, thr)


def build_ivf(codes: np.ndarray, doc_lengths: np.ndarray, K: int) -> Tuple[np.ndarray, np.ndarray]:
    """index.rs:479-499: per centroid the sorted, de-duplicated doc ids."""
    doc_of_tok = np.repeat(np.arange(len(doc_lengths), dtype=np.int64), doc_lengths)
    key = np.unique(codes.astype(np.int64) * np.int64(len(doc_lengths) + 1) + doc_of_tok)
    cen = key // np.int64(len(doc_lengths) + 1)
    ivf = (key % np.int64(len(doc_lengths) + 1)).astype(np.int64)
    lengths = np.bincount(cen, minlength=K).astype(np.int32)
    return ivf, lengths


def encode_documents(embeddings: Sequence[np.ndarray], art: CodecArtifacts, nbits: int):
    """encode_index_chunk, index.rs:289-371."""
    flat = _f32(np.concatenate(embeddings, 0))
    codes = compress_into_codes(flat, art.centroids)
    res = residuals_of(flat, art.centroids, codes)
    packed = quantize_residuals(res, art.bucket_cutoffs, nbits)
    doclens = np.array([e.shape[0] for e in embeddings], dtype=np.int64)
    return codes, packed, doclens


def create_index(embeddings: Sequence[np.ndarray], nbits: int = 4, seed: int = 42,
                 num_partitions: Optional[int] = None, kmeans_niters: int = 4) -> Index:
    """MmapIndex::create_with_kmeans in memory (index.rs:1392 -> kmeans.rs:261 -> index.rs:551)."""
    D = len(embeddings)
    rng = np.random.default_rng(seed)
    n_samples = min(int(1.0 + 16.0 * np.sqrt(120.0 * D)), D)                # kmeans.rs:273-276
    sidx = rng.permutation(D)[:n_samples]
    samples = np.concatenate([embeddings[int(i)] for i in sidx], 0)
    K = num_partitions or num_partitions_heuristic(D, [embeddings[int(i)].shape[0] for i in sidx])
    K = min(K, samples.shape[0])
    cent = kmeans(samples, K, kmeans_niters, seed)
    art = prepare_codec_artifacts(embeddings, cent, nbits, seed)
    codes, packed, doclens = encode_documents(embeddings, art, nbits)
    ivf, ivf_lengths = build_ivf(codes, doclens, cent.shape[0])
    return Index(cent, art.bucket_weights, art.bucket_cutoffs, codes, packed, doclens, ivf,
                 ivf_lengths, nbits)


# --------------------------------------------------------------------------------------
# on-disk format
# --------------------------------------------------------------------------------------

def write_index(index: Index, path: str, chunk_docs: int = 50_000, merged: bool = False) -> None:
    """Write the reference's index directory (index.rs:394-525).  numpy's NPY v1.0 writer produces
    the same 64-byte-aligned headers the reference reads (mmap.rs:754-1010)."""
    os.makedirs(path, exist_ok=True)
    np.save(os.path.join(path, "centroids.npy"), index.centroids)
    if index.bucket_cutoffs is not None:
        np.save(os.path.join(path, "bucket_cutoffs.npy"), _f32(index.bucket_cutoffs))
    np.save(os.path.join(path, "bucket_weights.npy"), index.bucket_weights)
    np.save(os.path.join(path, "avg_residual.npy"), np.zeros(index.dim, np.float32))
    np.save(os.path.join(path, "cluster_threshold.npy"), np.zeros(1, np.float32))
    np.save(os.path.join(path, "ivf.npy"), index.ivf)
    np.save(os.path.join(path, "ivf_lengths.npy"), index.ivf_lengths)
    D = index.num_documents
    n_chunks = max((D + chunk_docs - 1) // chunk_docs, 1)
    off = 0
    for i in range(n_chunks):
        d0, d1 = i * chunk_docs, min((i + 1) * chunk_docs, D)
        t0, t1 = int(index.doc_offsets[d0]), int(index.doc_offsets[d1])
        np.save(os.path.join(path, f"{i}.codes.npy"), index.codes[t0:t1])
        np.save(os.path.join(path, f"{i}.residuals.npy"), index.residuals[t0:t1])
        with open(os.path.join(path, f"doclens.{i}.json"), "w") as f:
            json.dump([int(x) for x in index.doc_lengths[d0:d1]], f)
        with open(os.path.join(path, f"{i}.metadata.json"), "w") as f:
            json.dump({"num_documents": d1 - d0, "num_embeddings": t1 - t0,
                       "embedding_offset": off}, f, indent=2)
        off += t1 - t0
    with open(os.path.join(path, "plan.json"), "w") as f:
        json.dump({"nbits": index.nbits, "num_chunks": n_chunks}, f, indent=2)
    meta = {"num_chunks": n_chunks, "nbits": index.nbits, "num_partitions": index.num_centroids,
            "num_embeddings": index.num_embeddings,
            "avg_doclen": index.num_embeddings / max(D, 1), "num_documents": D,
            "embedding_dim": index.dim, "next_plaid_compatible": True}
    with open(os.path.join(path, "metadata.json"), "w") as f:
        json.dump(meta, f, indent=2)
    if merged:  # merged_*.npy with `max_len - last_len` zero padding rows (index.rs:1113-1120)
        pad = int(index.doc_lengths.max() - index.doc_lengths[-1]) if D else 0
        np.save(os.path.join(path, "merged_codes.npy"),
                np.concatenate([index.codes, np.zeros(pad, np.int64)]))
        np.save(os.path.join(path, "merged_residuals.npy"),
                np.concatenate([index.residuals,
                                np.zeros((pad, index.residuals.shape[1]), np.uint8)]))


def load_index(path: str) -> Index:
    """MmapIndex::load (index.rs:1026-1139) reading the chunk files directly."""
    with open(os.path.join(path, "metadata.json")) as f:
        meta = json.load(f)
    codes, res, doclens = [], [], []
    for i in range(meta["num_chunks"]):
        codes.append(np.load(os.path.join(path, f"{i}.codes.npy")))
        res.append(np.load(os.path.join(path, f"{i}.residuals.npy")))
        with open(os.path.join(path, f"doclens.{i}.json")) as f:
            doclens.extend(json.load(f))
    cut_p = os.path.join(path, "bucket_cutoffs.npy")
    return Index(np.load(os.path.join(path, "centroids.npy")),
                 np.load(os.path.join(path, "bucket_weights.npy")),
                 np.load(cut_p) if os.path.exists(cut_p) else None,
                 np.concatenate(codes), np.concatenate(res, 0), np.array(doclens, np.int64),
                 np.load(os.path.join(path, "ivf.npy")), np.load(os.path.join(path, "ivf_lengths.npy")),
                 int(meta["nbits"]))


# --------------------------------------------------------------------------------------
# seeded synthetic corpora for tests (clustered unit vectors)
# --------------------------------------------------------------------------------------

def synthetic_corpus(num_docs: int, doclen: int, dim: int = 128, seed: int = 42,
                     ragged: bool = False, n_topics: Optional[int] = None):
    """Topic-clustered unit vectors: i.i.d. random vectors would all fall under the default
    centroid_score_threshold=0.4 (search.rs:54-56) and every search would come back empty."""
    rng = np.random.default_rng(seed)
    n_topics = n_topics or max(num_docs // 16, 4)
    centres = rng.standard_normal((n_topics * 8, dim)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    docs = []
    for d in range(num_docs):
        T = int(rng.integers(max(doclen * 7 // 10, 1), doclen * 13 // 10 + 1)) if ragged else doclen
        topic = int(rng.integers(n_topics))
        pick = topic * 8 + rng.integers(0, 8, size=T)
        g = rng.standard_normal((T, dim)).astype(np.float32) / np.sqrt(dim)
        x = centres[pick] + 0.35 * g
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        docs.append(x.astype(np.float32))
    return docs


def synthetic_queries(docs: Sequence[np.ndarray], n_queries: int, nq: int = 32, seed: int = 7):
    rng = np.random.default_rng(seed)
    dim = docs[0].shape[1]
    out, src = [], []
    for _ in range(n_queries):
        d = int(rng.integers(len(docs)))
        tok = docs[d][rng.integers(0, docs[d].shape[0], size=nq)]
        g = rng.standard_normal((nq, dim)).astype(np.float32) / np.sqrt(dim)
        q = tok + 0.2 * g
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        out.append(q.astype(np.float32))
        src.append(d)
    return out, src
