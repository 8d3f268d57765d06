"""Index-build micro-benchmark (SURVEY 8 a12 / BASELINE config D shape): encode `--tokens` token vectors
against K = 2^log2k centroids (nearest centroid + residual quantise + pack) through pb_codec_encode_chunk.
Prints one JSON line; run under ncu for the kernel-level numbers (profiles/)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=1 << 20)
    ap.add_argument("--log2k", type=int, default=18)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--nbits", type=int, default=4)
    ap.add_argument("--check", type=int, default=2048, help="tokens verified against the CPU oracle")
    a = ap.parse_args()
    import next_plaid_b200 as npb
    rng = np.random.default_rng(42)
    K = 1 << a.log2k
    cent = rng.standard_normal((K, a.dim), dtype=np.float32)
    cent /= np.linalg.norm(cent, axis=1, keepdims=True)
    emb = cent[rng.integers(0, K, a.tokens)] + 0.35 * rng.standard_normal((a.tokens, a.dim), dtype=np.float32) / np.sqrt(a.dim)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    nopt = 1 << a.nbits
    cut = np.quantile((emb[:4096] - cent[:4096]).ravel() * 0.2, [i / nopt for i in range(1, nopt)]).astype(np.float32)
    codec = npb.ResidualCodec(a.nbits, cent, cut)
    codec.encode_chunk(emb[:4096])          # warm-up (module load, allocations)
    t0 = time.perf_counter()
    codes, packed = codec.encode_chunk(emb)
    dt = time.perf_counter() - t0
    st = codec.last_assign_stats()
    out = {"tokens": a.tokens, "num_centroids": K, "dim": a.dim, "nbits": a.nbits, "seconds_e2e_host_buffers": dt,
           "tokens_per_s_e2e": a.tokens / dt, "assign_pair_rate_e2e": a.tokens * K / dt,
           "tflops_equiv_e2e": 2.0 * a.tokens * K * a.dim / dt / 1e12, "assign_stats": st}
    if a.check:
        from oracle import oracle
        idx = rng.choice(a.tokens, a.check, replace=False)
        want = oracle.compress_into_codes(emb[idx], cent)
        out["oracle_check"] = {"tokens": int(a.check), "codes_identical": bool(np.array_equal(want, codes[idx]))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
