"""Pin the CPU oracle against the reference's own known-answer unit tests.

Each test names the reference test it lifts (paths relative to /root/reference/next-plaid/src).
These are the only fixed vectors the reference holds for the search path: its
search tests use unseeded random data and assert properties only.
"""
import numpy as np
import pytest


def test_maxsim_basic_1p7(oracle):
    # maxsim.rs:393-413 test_maxsim_score_basic == search.rs:685-705 test_colbert_score
    q = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    d = np.array([[0.5, 0.5, 0, 0], [0.8, 0.2, 0, 0], [0, 0.9, 0.1, 0]], np.float32)
    assert abs(oracle.maxsim_score(q, d) - 1.7) < 1e-5


def test_maxsim_ignores_non_finite_rows_8p0(oracle):
    # maxsim.rs:498-507
    q = np.tile(np.array([1.0, 0.0], np.float32), (16, 1))
    d = np.tile(np.array([0.5, 0.0], np.float32), (16, 1))
    d[15] = [np.nan, 0.0]
    assert abs(oracle.maxsim_score(q, d) - 8.0) < 1e-5


def test_rerank_api_scores_2_1_0(oracle):
    # next-plaid-api/tests/integration_tests.rs:2301-2376 (exact rerank scores 2.0 / 1.0 / 0.0)
    q = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    docs = [np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32),
            np.array([[1, 0, 0, 0], [0, 0, 1, 0]], np.float32),
            np.array([[0, 0, 1, 0], [0, 0, 0, 1]], np.float32)]
    got = [oracle.maxsim_score(q, d) for d in docs]
    assert np.allclose(got, [2.0, 1.0, 0.0], atol=1e-2)


def test_score_order_non_finite_last(oracle):
    # search.rs:717-726 test_cmp_score_descending_places_non_finite_scores_last
    L = oracle.lib()
    import functools
    xs = [1.0, float("inf"), 0.5, float("nan")]
    xs.sort(key=functools.cmp_to_key(lambda a, b: L.po_cmp_score_ascending(b, a)))
    assert xs[0] == 1.0 and xs[1] == 0.5
    assert not np.isfinite(xs[2]) and not np.isfinite(xs[3])


def test_score_replacement_finite_beats_non_finite(oracle):
    # search.rs:728-734 and :736-742
    L = oracle.lib()
    assert L.po_is_score_better(1.0, float("nan"))
    assert L.po_is_score_better(1.0, float("inf"))
    assert not L.po_is_score_better(float("nan"), 1.0)
    assert not L.po_is_score_better(float("inf"), 1.0)
    assert L.po_max_score(float("nan"), 1.0) == 1.0
    assert L.po_max_score(1.0, float("nan")) == 1.0
    assert L.po_max_score(float("inf"), 1.0) == 1.0
    assert L.po_max_score(1.0, float("inf")) == 1.0


def test_assign_to_centroids(oracle):
    # maxsim.rs:444-477 test_assign_to_centroids -> [0,1,2,0,2]
    c = np.eye(3, 4, dtype=np.float32)
    e = np.array([[0.9, 0.1, 0, 0], [0.1, 0.9, 0, 0], [0, 0.1, 0.9, 0], [0.8, 0.2, 0, 0],
                  [0, 0, 0.8, 0.2]], np.float32)
    assert oracle.compress_into_codes(e, c).tolist() == [0, 1, 2, 0, 2]


def test_compress_into_codes(oracle):
    # codec.rs:637-663 -> [0, 2]
    c = np.eye(3, 4, dtype=np.float32)
    e = np.array([[0.9, 0.1, 0, 0], [0, 0, 0.95, 0.05]], np.float32)
    assert oracle.compress_into_codes(e, c).tolist() == [0, 2]


def test_compress_into_codes_ignores_nan(oracle):
    # codec.rs:733-752: a NaN centroid is never the argmax when a finite choice exists -> 1
    c = np.array([[np.nan, 0], [1, 0], [0, 1]], np.float32)
    e = np.array([[1, 0]], np.float32)
    assert oracle.compress_into_codes(e, c).tolist() == [1]


def test_compress_last_max_wins_on_exact_tie(oracle):
    # codec.rs:329-337: Iterator::max_by returns the last maximum
    c = np.array([[1, 0], [1, 0], [0, 1]], np.float32)
    assert oracle.compress_into_codes(np.array([[1, 0]], np.float32), c).tolist() == [1]


def test_quantize_decompress_roundtrip_4bit(oracle):
    # codec.rs:666-730
    dim = 8
    cent = np.zeros((4, dim), np.float32)
    cut = np.array([(i / 16.0 - 0.5) * 2.0 for i in range(1, 16)], np.float32)
    wts = np.array([((i + 0.5) / 16.0 - 0.5) * 2.0 for i in range(16)], np.float32)
    res = np.array([[-0.9, -0.7, -0.5, -0.3, 0.0, 0.3, 0.5, 0.9],
                    [-0.8, -0.4, 0.0, 0.4, 0.8, -0.6, 0.2, 0.6]], np.float32)
    packed = oracle.quantize_residuals(res, cut, 4)
    assert packed.shape == (2, dim * 4 // 8)
    rec = oracle.decompress(cent, wts, 4, packed, np.zeros(2, np.int64))
    for i in range(2):
        for j in range(dim):
            if abs(res[i, j]) > 0.2:
                assert (res[i, j] > 0) == (rec[i, j] > 0) or abs(rec[i, j]) < 0.1
    # decompressed rows are unit norm (codec.rs:464-467)
    assert np.allclose(np.linalg.norm(rec, axis=1), 1.0, atol=1e-6)


def test_byte_layout_probe(oracle):
    # buckets [13, 10] <-> byte 0xB5 (first dim in the high bits, each bucket
    # bit-reversed within its field): packer codec.rs:384-395, LUTs :168-214
    cut = np.arange(1, 16, dtype=np.float32)      # bucket(v) = #{cutoffs < v}
    res = np.array([[13.5, 10.5]], np.float32)
    assert oracle.quantize_residuals(res, cut, 4)[0, 0] == 0xB5
    rev, look = oracle.byte_reversed_bits_map(4), oracle.bucket_index_lookup(4)
    assert look[rev[0xB5]].tolist() == [13, 10]
    # 2-bit: buckets [3,0,1,2] -> bit-reversed fields 11 00 10 01
    cut2 = np.arange(1, 4, dtype=np.float32)
    res2 = np.array([[3.5, 0.5, 1.5, 2.5]], np.float32)
    b = oracle.quantize_residuals(res2, cut2, 2)[0, 0]
    assert b == 0b11001001
    assert oracle.bucket_index_lookup(2)[oracle.byte_reversed_bits_map(2)[b]].tolist() == [3, 0, 1, 2]


def test_quantile_values(oracle):
    # utils.rs:289-294
    a = np.array([1, 2, 3, 4, 5], np.float32)
    got = oracle.quantiles(a, [0.5, 0.0, 1.0])
    assert np.allclose(got, [3.0, 1.0, 5.0], atol=1e-6)


def test_search_parameters_defaults(oracle):
    # search.rs:708-715
    p = oracle.SearchParameters()
    assert (p.batch_size, p.n_full_scores, p.top_k, p.n_ivf_probe) == (2000, 4096, 10, 8)
    assert p.centroid_score_threshold == 0.4 and p.centroid_batch_size == 100_000


def test_pinned_dot_matches_float64_to_1e5(oracle):
    # the reference's own bar for any sgemm order is 1e-5 (maxsim.rs:412)
    rng = np.random.default_rng(0)
    q = rng.standard_normal((32, 128)).astype(np.float32)
    c = rng.standard_normal((500, 128)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    S = oracle.centroid_scores(q, c)
    assert np.abs(S - q.astype(np.float64) @ c.astype(np.float64).T).max() < 1e-5
    # and the vectorised routine equals the scalar pinned dot bit for bit
    L = oracle.lib()
    import ctypes as C
    for (i, j) in [(0, 0), (3, 17), (31, 499)]:
        s = L.po_dot(q[i].ctypes.data_as(C.c_void_p), c[j].ctypes.data_as(C.c_void_p), 128)
        assert np.float32(s) == S[i, j]


def test_find_outliers_kat(oracle):
    # update.rs:1170-1185 test_find_outliers: centroids (0,0), (1,1); embeddings near each and one at
    # (5,5); threshold_sq = 1.0 -> only row 2 is an outlier
    c = np.array([[0, 0], [1, 1]], np.float32)
    e = np.array([[0.1, 0.1], [0.9, 0.9], [5.0, 5.0]], np.float32)
    assert oracle.find_outliers(e, c, 1.0).tolist() == [2]
    # borderline rows take the f64 recheck (update.rs:592-599): exactly on the threshold is not an outlier
    c2 = np.zeros((1, 4), np.float32)
    e2 = np.array([[1, 0, 0, 0], [1.0000001, 0, 0, 0], [1.01, 0, 0, 0]], np.float32)
    assert oracle.find_outliers(e2, c2, 1.0).tolist() == [1, 2]
    assert oracle.find_outliers(np.zeros((0, 4), np.float32), c2, 1.0).tolist() == []
