/*
 * plaid_oracle.c -- CPU restatement of the next-plaid PLAID search / codec path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (next-plaid_b200/) may
 * link, import or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and there only as the
 * checker or as the timed CPU arm.
 *
 * What it restates (paths relative to /root/reference/next-plaid/src):
 *   score ordering ............ search.rs:110-133 (dup. maxsim.rs:23-35, codec.rs:23-30)
 *   dense search .............. search.rs:327-516
 *   batched search ............ search.rs:140-254 (probe), :259-302 (sparse approx), :521-640
 *   candidates ................ index.rs:1142-1156
 *   approximate score ......... search.rs:305-324
 *   decompress ................ codec.rs:423-470 (LUTs :168-214), index.rs:1159-1179
 *   MaxSim .................... maxsim.rs:270-315
 *   nearest centroid .......... codec.rs:297-343
 *   quantize + pack ........... codec.rs:356-411
 *   quantiles ................. utils.rs:94-149
 *
 * The reference cannot be built here (no cargo/rustc, no network), so this file is
 * pinned by the reference's own known-answer unit tests (tests/test_oracle_kats.py
 * lists each with its file:line).
 *
 * Floating point.  Every contraction in the reference goes through ndarray `.dot`
 * (matrixmultiply 0.3.10 or a BLAS): the fp32 accumulation order is implementation
 * defined and the reference's tests only pin it to 1e-5.  This oracle PINS one
 * order, which is a valid instance of the reference's algorithm and is the order
 * the CUDA kernels use, so GPU-vs-oracle comparisons are bit-exact:
 *     dot(a,b)   = acc=+0; for j=0..dim-1: acc = fmaf(a[j], b[j], acc)
 *     sumsq(row) = 32 interleaved partial sums (element j belongs to lane (j/4)%32,
 *                  accumulated with fmaf in increasing j) combined by the xor
 *                  butterfly 16,8,4,2,1 (p[l] += p[l^m])
 * Everything else (centroid + weight add, sqrt, divide, the q-ordered score sums)
 * is a single IEEE op in the reference and is reproduced as such.
 * Compile with -ffp-contract=off so nothing else is fused.
 *
 * Unspecified-in-the-reference behaviour pinned here: exact score ties at the
 * n_ivf_probe boundary (select_nth_unstable / heap-array order) resolve to the lower
 * centroid index.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* score ordering: search.rs:110-133                                          */
/* ------------------------------------------------------------------------- */

static inline int32_t f32_total_key(float x) {
    /* f32::total_cmp: sign-magnitude bits -> two's complement order */
    int32_t b;
    memcpy(&b, &x, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}

/* returns <0, 0, >0 like Ordering::{Less,Equal,Greater}; search.rs:110-117 */
PO_EXPORT int po_cmp_score_ascending(float a, float b) {
    int fa = isfinite(a), fb = isfinite(b);
    if (fa && fb) {
        int32_t ka = f32_total_key(a), kb = f32_total_key(b);
        return (ka > kb) - (ka < kb);
    }
    if (fa && !fb) return 1;
    if (!fa && fb) return -1;
    return 0;
}

/* search.rs:123-125 */
PO_EXPORT int po_is_score_better(float candidate, float current) {
    return po_cmp_score_ascending(candidate, current) > 0;
}

/* search.rs:127-133 */
PO_EXPORT float po_max_score(float a, float b) {
    return po_is_score_better(b, a) ? b : a;
}

/* ------------------------------------------------------------------------- */
/* pinned contractions                                                        */
/* ------------------------------------------------------------------------- */

PO_EXPORT float po_dot(const float *a, const float *b, int dim) {
    float acc = 0.0f;
    for (int j = 0; j < dim; ++j) acc = fmaf(a[j], b[j], acc);
    return acc;
}

PO_EXPORT float po_sumsq(const float *v, int dim) {
    float p[32];
    for (int l = 0; l < 32; ++l) p[l] = 0.0f;
    for (int j = 0; j < dim; ++j) {
        int l = (j >> 2) & 31;
        p[l] = fmaf(v[j], v[j], p[l]);
    }
    for (int m = 16; m >= 1; m >>= 1) {
        float t[32];
        for (int l = 0; l < 32; ++l) t[l] = p[l] + p[l ^ m];
        memcpy(p, t, sizeof(p));
    }
    return p[0];
}

/* S[q][c] = dot(Q[q], C[c]); search.rs:345 / :174 / :268.  Row-major [nq][K]. */
PO_EXPORT void po_centroid_scores(const float *Q, int nq, const float *C, int64_t K, int dim,
                                  float *S) {
    if (nq <= 0 || K <= 0) return;
    /* transpose Q so the q loop vectorises; every lane still runs the pinned order */
    float *Qt = (float *)malloc((size_t)dim * nq * sizeof(float));
    for (int q = 0; q < nq; ++q)
        for (int j = 0; j < dim; ++j) Qt[(size_t)j * nq + q] = Q[(size_t)q * dim + j];
#pragma omp parallel
    {
        float *acc = (float *)malloc((size_t)nq * sizeof(float));
#pragma omp for schedule(static)
        for (int64_t c = 0; c < K; ++c) {
            const float *cr = C + (size_t)c * dim;
            for (int q = 0; q < nq; ++q) acc[q] = 0.0f;
            for (int j = 0; j < dim; ++j) {
                const float cj = cr[j];
                const float *qt = Qt + (size_t)j * nq;
                for (int q = 0; q < nq; ++q) acc[q] = fmaf(qt[q], cj, acc[q]);
            }
            for (int q = 0; q < nq; ++q) S[(size_t)q * K + c] = acc[q];
        }
        free(acc);
    }
    free(Qt);
}

/* ------------------------------------------------------------------------- */
/* codec: codec.rs                                                            */
/* ------------------------------------------------------------------------- */

static inline uint32_t bitrev_n(uint32_t v, int nbits) {
    uint32_t r = 0;
    for (int k = 0; k < nbits; ++k)
        if (v & (1u << k)) r |= 1u << (nbits - 1 - k);
    return r;
}

/* byte_reversed_bits_map, codec.rs:168-196 */
PO_EXPORT void po_byte_reversed_bits_map(int nbits, uint8_t *out256) {
    uint32_t mask = (1u << nbits) - 1;
    for (int i = 0; i < 256; ++i) {
        uint32_t val = (uint32_t)i, out = 0;
        int pos = 8;
        while (pos >= nbits) {
            uint32_t segment = (val >> (pos - nbits)) & mask;
            out |= bitrev_n(segment, nbits);
            if (pos > nbits) out <<= nbits;
            pos -= nbits;
        }
        out256[i] = (uint8_t)out;
    }
}

/* bucket_weight_indices_lookup, codec.rs:198-214: table[byte][kpb-1-k] = (byte >> k*nbits) & mask */
PO_EXPORT void po_bucket_index_lookup(int nbits, uint8_t *out /* [256][8/nbits] */) {
    int kpb = 8 / nbits;
    uint32_t mask = (1u << nbits) - 1;
    for (int b = 0; b < 256; ++b)
        for (int k = kpb - 1; k >= 0; --k)
            out[b * kpb + (kpb - 1 - k)] = (uint8_t)(((uint32_t)b >> (k * nbits)) & mask);
}

/* decompress one token row; codec.rs:443-467 */
static void decompress_row(const float *centroid, const uint8_t *packed, int packed_dim, int dim,
                           int nbits, const float *weights, const uint8_t *revmap,
                           const uint8_t *lookup, float *out) {
    int kpb = 8 / nbits;
    int r = 0;
    for (int b = 0; b < packed_dim; ++b) {
        const uint8_t *idx = lookup + (size_t)revmap[packed[b]] * kpb;
        for (int k = 0; k < kpb; ++k) {
            if (r < dim) {
                out[r] = centroid[r] + weights[idx[k]];
                ++r;
            }
        }
    }
    for (; r < dim; ++r) out[r] = 0.0f; /* Array2::zeros for dims the packed row does not cover */
    float norm = sqrtf(po_sumsq(out, dim));
    if (!(norm >= 1e-12f)) norm = 1e-12f; /* f32::max(1e-12): NaN.max(x) = x */
    for (int j = 0; j < dim; ++j) out[j] = out[j] / norm;
}

/* ResidualCodec::decompress, codec.rs:423-470.  codes are i64 as stored on disk. */
PO_EXPORT void po_decompress(const float *centroids, int dim, int nbits, const float *weights,
                             const uint8_t *packed, const int64_t *codes, int64_t n, float *out) {
    uint8_t revmap[256], lookup[256 * 8];
    int packed_dim = dim * nbits / 8;
    po_byte_reversed_bits_map(nbits, revmap);
    po_bucket_index_lookup(nbits, lookup);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        decompress_row(centroids + (size_t)codes[i] * dim, packed + (size_t)i * packed_dim,
                       packed_dim, dim, nbits, weights, revmap, lookup, out + (size_t)i * dim);
}

/* quantize_residuals, codec.rs:356-411: bucket = #{cutoffs < v}; bits LSB-first into an
 * MSB-first bit stream */
PO_EXPORT void po_quantize_residuals(const float *residuals, int64_t n, int dim, int nbits,
                                     const float *cutoffs, uint8_t *packed) {
    int ncut = (1 << nbits) - 1;
    int packed_dim = dim * nbits / 8;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint8_t *row = packed + (size_t)i * packed_dim;
        memset(row, 0, (size_t)packed_dim);
        int bit_idx = 0;
        for (int j = 0; j < dim; ++j) {
            float v = residuals[(size_t)i * dim + j];
            unsigned bucket = 0;
            for (int c = 0; c < ncut; ++c) bucket += (v > cutoffs[c]);
            for (int b = 0; b < nbits; ++b) {
                uint8_t bit = (bucket >> b) & 1u;
                row[bit_idx / 8] |= (uint8_t)(bit << (7 - (bit_idx % 8)));
                ++bit_idx;
            }
        }
    }
}

/* compress_into_codes_cpu, codec.rs:297-343: argmax_c dot(x, C_c) under the score order;
 * Iterator::max_by keeps the LAST maximum on exact ties. */
PO_EXPORT void po_compress_into_codes(const float *emb, int64_t n, const float *C, int64_t K,
                                      int dim, int64_t *codes) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        const float *x = emb + (size_t)i * dim;
        int64_t best = 0;
        float best_s = 0.0f;
        for (int64_t c = 0; c < K; ++c) {
            float s = po_dot(x, C + (size_t)c * dim, dim);
            if (c == 0 || po_cmp_score_ascending(s, best_s) >= 0) {
                best = c;
                best_s = s;
            }
        }
        codes[i] = best;
    }
}

/* compress_and_residuals_cpu, index.rs:17-40 */
PO_EXPORT void po_residuals(const float *emb, int64_t n, const float *C, int dim,
                            const int64_t *codes, float *res) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < dim; ++j)
            res[(size_t)i * dim + j] = emb[(size_t)i * dim + j] - C[(size_t)codes[i] * dim + j];
}

/* ------------------------------------------------------------------------- */
/* MaxSim: maxsim.rs:270-315                                                  */
/* ------------------------------------------------------------------------- */

PO_EXPORT float po_maxsim(const float *Q, int nq, const float *D, int64_t T, int dim) {
    /* both reference branches (simple loop for nq*T<256, GEMM + simd_max otherwise) reduce to:
     * per query token the best FINITE similarity; rows with no finite similarity add nothing */
    float total = 0.0f;
    for (int q = 0; q < nq; ++q) {
        float m = -INFINITY;
        for (int64_t t = 0; t < T; ++t) {
            float s = po_dot(Q + (size_t)q * dim, D + (size_t)t * dim, dim);
            if (po_is_score_better(s, m)) m = s;
        }
        if (isfinite(m)) total += m;
    }
    return total;
}

/* ------------------------------------------------------------------------- */
/* index view + search                                                        */
/* ------------------------------------------------------------------------- */

typedef struct {
    int32_t dim;
    int32_t nbits;
    int64_t num_centroids;
    int64_t num_docs;
    const float *centroids;      /* [K][dim]  centroids.npy */
    const float *bucket_weights; /* [2^nbits] bucket_weights.npy */
    const int64_t *codes;        /* [N]       merged_codes.npy */
    const uint8_t *residuals;    /* [N][dim*nbits/8] merged_residuals.npy */
    const int64_t *doc_offsets;  /* [D+1]     index.rs:1107-1110 */
    const int64_t *ivf;          /* ivf.npy */
    const int32_t *ivf_lengths;  /* [K] ivf_lengths.npy */
    const int64_t *ivf_offsets;  /* [K+1]     index.rs:1089-1094 */
} po_index;

typedef struct {
    int64_t n_full_scores;
    int64_t top_k;
    int64_t n_ivf_probe;
    int64_t centroid_batch_size;
    int32_t has_threshold;
    float centroid_score_threshold;
} po_params;

/* optional stage outputs for stage-level parity tests (any pointer may be NULL) */
typedef struct {
    int64_t *cells;      /* capacity cells_cap, ascending */
    int64_t cells_cap;
    int64_t n_cells;
    int64_t *candidates; /* capacity cand_cap, ascending doc ids */
    float *approx;       /* capacity cand_cap, approx score per candidate */
    int64_t cand_cap;
    int64_t n_candidates;
    int64_t *kept;       /* capacity kept_cap: docs sent to exact scoring, in approx-rank order */
    float *kept_exact;   /* capacity kept_cap */
    int64_t kept_cap;
    int64_t n_kept;
    int32_t variant;     /* out: 0 dense, 1 batched */
} po_trace;

typedef struct {
    float s;
    int64_t id;
    int64_t pos;
} scored;

static int cmp_scored_desc_stable(const void *a, const void *b) {
    const scored *x = (const scored *)a, *y = (const scored *)b;
    int c = po_cmp_score_ascending(y->s, x->s); /* cmp_score_descending(a,b)=asc(b,a) */
    if (c) return c;
    return (x->pos > y->pos) - (x->pos < y->pos); /* sort_by is stable */
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* "better" for probe selection with the pinned tie rule: higher score, then lower index */
static inline int probe_better(float sa, int64_t ca, float sb, int64_t cb) {
    int c = po_cmp_score_ascending(sa, sb);
    if (c) return c > 0;
    return ca < cb;
}

/* top-n of a row under (score desc, index asc); out gets n indices (unordered semantics) */
static void top_n_of_row(const float *row, const int64_t *ids, int64_t len, int64_t n,
                         int64_t *out) {
    /* simple bounded insertion list: n is small (n_ivf_probe) relative to len */
    float *bs = (float *)malloc((size_t)n * sizeof(float));
    int64_t cnt = 0;
    for (int64_t i = 0; i < len; ++i) {
        int64_t c = ids ? ids[i] : i;
        float s = ids ? row[c] : row[i];
        if (cnt == n && !probe_better(s, c, bs[cnt - 1], out[cnt - 1])) continue;
        int64_t p = cnt < n ? cnt : n - 1;
        while (p > 0 && probe_better(s, c, bs[p - 1], out[p - 1])) {
            bs[p] = bs[p - 1];
            out[p] = out[p - 1];
            --p;
        }
        bs[p] = s;
        out[p] = c;
        if (cnt < n) ++cnt;
    }
    free(bs);
}

/* dense probe: search.rs:350-428.  S row-major [nq][K]. Returns malloc'd ascending cells. */
static int64_t *probe_dense(const po_index *ix, const float *S, int nq, const po_params *p,
                            const int64_t *subset, int64_t n_subset, int has_subset,
                            int64_t *n_cells_out) {
    int64_t K = ix->num_centroids;
    uint8_t *sel = (uint8_t *)calloc((size_t)K + 1, 1);
    int64_t *elig = NULL, n_elig = 0;
    int64_t n_probe_eff = p->n_ivf_probe;
    if (has_subset) {
        /* eligible centroids = codes of the subset docs; search.rs:350-364 */
        uint8_t *em = (uint8_t *)calloc((size_t)K + 1, 1);
        for (int64_t i = 0; i < n_subset; ++i) {
            int64_t d = subset[i];
            if (d < 0 || d >= ix->num_docs) continue; /* `as usize` of a negative id is >= len */
            for (int64_t t = ix->doc_offsets[d]; t < ix->doc_offsets[d + 1]; ++t)
                em[ix->codes[t]] = 1;
        }
        elig = (int64_t *)malloc((size_t)(K + 1) * sizeof(int64_t));
        for (int64_t c = 0; c < K; ++c)
            if (em[c]) elig[n_elig++] = c;
        free(em);
        if (n_elig > 0) { /* search.rs:370-382, u64 integer arithmetic */
            uint64_t scaled = n_subset > 0
                                  ? (uint64_t)p->n_ivf_probe * (uint64_t)ix->num_docs / (uint64_t)n_subset
                                  : (uint64_t)p->n_ivf_probe;
            if (scaled < (uint64_t)p->n_ivf_probe) scaled = (uint64_t)p->n_ivf_probe;
            if (scaled > (uint64_t)n_elig) scaled = (uint64_t)n_elig;
            n_probe_eff = (int64_t)scaled;
        }
    }
    int64_t pool = has_subset ? n_elig : K;
    int64_t n_probe = n_probe_eff < pool ? n_probe_eff : pool;
    if (n_probe > 0) {
        int64_t *top = (int64_t *)malloc((size_t)n_probe * sizeof(int64_t));
        for (int q = 0; q < nq; ++q) {
            top_n_of_row(S + (size_t)q * K, has_subset ? elig : NULL, pool, n_probe, top);
            for (int64_t i = 0; i < n_probe; ++i) sel[top[i]] = 1;
        }
        free(top);
    }
    free(elig);
    int64_t *cells = (int64_t *)malloc((size_t)(K + 1) * sizeof(int64_t));
    int64_t n = 0;
    for (int64_t c = 0; c < K; ++c) {
        if (!sel[c]) continue;
        if (p->has_threshold) { /* search.rs:417-425: true max over all query tokens */
            float m = -INFINITY;
            int have = 0;
            for (int q = 0; q < nq; ++q) {
                float s = S[(size_t)q * K + c];
                /* Iterator::max_by keeps the last max; value-wise only -0/+0 and the identity of
                 * a non-finite could differ, neither changes `m >= threshold` */
                if (!have || po_cmp_score_ascending(s, m) >= 0) m = s;
                have = 1;
            }
            if (!have) m = -INFINITY;
            if (!(m >= p->centroid_score_threshold)) continue;
        }
        cells[n++] = c;
    }
    free(sel);
    *n_cells_out = n;
    return cells;
}

/* batched probe: search.rs:140-254 (literal heap simulation per slab and token) */
typedef struct {
    float s;
    int64_t c;
} hent;

/* index of the heap top of BinaryHeap<(Reverse<OrdF32>, usize)>: min score, ties -> largest c */
static int64_t heap_top(const hent *h, int64_t n) {
    int64_t t = 0;
    for (int64_t i = 1; i < n; ++i) {
        int c = po_cmp_score_ascending(h[i].s, h[t].s);
        if (c < 0 || (c == 0 && h[i].c > h[t].c)) t = i;
    }
    return t;
}

static int cmp_hent_c(const void *a, const void *b) {
    const hent *x = (const hent *)a, *y = (const hent *)b;
    return (x->c > y->c) - (x->c < y->c);
}

static int64_t *probe_batched(const po_index *ix, const float *S, int nq, const po_params *p,
                              int64_t *n_cells_out) {
    int64_t K = ix->num_centroids, B = p->centroid_batch_size, n_probe = p->n_ivf_probe;
    int64_t n_slabs = (K + B - 1) / B;
    /* final_max_scores: HashMap<usize,f32>; present[] marks membership */
    float *final_max = (float *)malloc((size_t)(K + 1) * sizeof(float));
    uint8_t *present = (uint8_t *)calloc((size_t)K + 1, 1);
    hent *fheap = (hent *)malloc((size_t)nq * (size_t)(n_probe + 1) * sizeof(hent));
    int64_t *fcnt = (int64_t *)calloc((size_t)nq + 1, sizeof(int64_t));
    float *slab_max = (float *)malloc((size_t)B * sizeof(float));
    uint8_t *slab_has = (uint8_t *)malloc((size_t)B);
    hent *lheap = (hent *)malloc((size_t)(n_probe + 1) * sizeof(hent));
    for (int64_t sl = 0; sl < n_slabs; ++sl) {
        int64_t c0 = sl * B, c1 = c0 + B < K ? c0 + B : K, len = c1 - c0;
        /* :174 computes this slab's scores; they are the same pinned dots as S[q][c0..c1) */
        memset(slab_has, 0, (size_t)len);
        for (int q = 0; q < nq; ++q) {
            int64_t cnt = 0;
            const float *row = S + (size_t)q * K + c0;
            for (int64_t lc = 0; lc < len; ++lc) { /* :177-198 */
                float s = row[lc];
                int pushed = 0;
                if (cnt < n_probe) {
                    lheap[cnt].s = s;
                    lheap[cnt].c = c0 + lc;
                    ++cnt;
                    pushed = 1;
                } else if (cnt > 0) {
                    int64_t t = heap_top(lheap, cnt);
                    if (po_is_score_better(s, lheap[t].s)) {
                        lheap[t].s = s;
                        lheap[t].c = c0 + lc;
                        pushed = 1;
                    }
                }
                if (pushed) {
                    if (slab_has[lc]) slab_max[lc] = po_max_score(slab_max[lc], s);
                    else {
                        slab_max[lc] = s;
                        slab_has[lc] = 1;
                    }
                }
            }
            /* merge this slab's heap for token q into the final heap, :212-225 (pinned: ascending c) */
            qsort(lheap, (size_t)cnt, sizeof(hent), cmp_hent_c);
            hent *fh = fheap + (size_t)q * (size_t)(n_probe + 1);
            for (int64_t i = 0; i < cnt; ++i) {
                if (fcnt[q] < n_probe) fh[fcnt[q]++] = lheap[i];
                else if (fcnt[q] > 0) {
                    int64_t t = heap_top(fh, fcnt[q]);
                    if (po_is_score_better(lheap[i].s, fh[t].s)) fh[t] = lheap[i];
                }
            }
        }
        for (int64_t lc = 0; lc < len; ++lc) { /* :226-231, plain f32::max */
            if (!slab_has[lc]) continue;
            int64_t c = c0 + lc;
            if (present[c]) final_max[c] = fmaxf(final_max[c], slab_max[lc]);
            else {
                final_max[c] = slab_max[lc];
                present[c] = 1;
            }
        }
    }
    uint8_t *sel = (uint8_t *)calloc((size_t)K + 1, 1);
    for (int q = 0; q < nq; ++q)
        for (int64_t i = 0; i < fcnt[q]; ++i) sel[fheap[(size_t)q * (size_t)(n_probe + 1) + i].c] = 1;
    int64_t *cells = (int64_t *)malloc((size_t)(K + 1) * sizeof(int64_t));
    int64_t n = 0;
    for (int64_t c = 0; c < K; ++c) {
        if (!sel[c]) continue;
        if (p->has_threshold) { /* :243-251 */
            float m = present[c] ? final_max[c] : -INFINITY;
            if (!(m >= p->centroid_score_threshold)) continue;
        }
        cells[n++] = c;
    }
    free(sel);
    free(lheap);
    free(slab_has);
    free(slab_max);
    free(fcnt);
    free(fheap);
    free(present);
    free(final_max);
    *n_cells_out = n;
    return cells;
}

/* get_candidates, index.rs:1142-1156 */
static int64_t *get_candidates(const po_index *ix, const int64_t *cells, int64_t n_cells,
                               int64_t *n_out) {
    int64_t total = 0;
    for (int64_t i = 0; i < n_cells; ++i)
        if (cells[i] < ix->num_centroids) total += ix->ivf_lengths[cells[i]];
    int64_t *cand = (int64_t *)malloc((size_t)(total + 1) * sizeof(int64_t));
    int64_t n = 0;
    for (int64_t i = 0; i < n_cells; ++i) {
        int64_t c = cells[i];
        if (c >= ix->num_centroids) continue;
        memcpy(cand + n, ix->ivf + ix->ivf_offsets[c], (size_t)ix->ivf_lengths[c] * sizeof(int64_t));
        n += ix->ivf_lengths[c];
    }
    qsort(cand, (size_t)n, sizeof(int64_t), cmp_i64);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i)
        if (m == 0 || cand[m - 1] != cand[i]) cand[m++] = cand[i];
    *n_out = m;
    return cand;
}

/* One query.  Returns the number of results written (<= top_k), or -1 on bad arguments.
 * search.rs:327 (dispatch :337), dense :345-516, batched :521-640. */
PO_EXPORT int64_t po_search_one(const po_index *ix, const float *Q, int nq, const po_params *p,
                                const int64_t *subset, int64_t n_subset, int has_subset,
                                int64_t *out_ids, float *out_scores, po_trace *tr) {
    int64_t K = ix->num_centroids;
    int dim = ix->dim;
    if (p->n_ivf_probe < 1 || p->top_k < 0) return -1;
    int batched = p->centroid_batch_size > 0 && K > p->centroid_batch_size; /* :337 */
    if (tr) {
        tr->variant = batched;
        tr->n_cells = tr->n_candidates = tr->n_kept = 0;
    }
    /* S for every centroid.  The batched variant only ever reads S[q][c] for codes of candidate
     * docs (sparse map, :259-272) and those entries are the same pinned dot products. */
    float *S = (float *)malloc((size_t)(nq > 0 ? nq : 1) * (size_t)K * sizeof(float));
    po_centroid_scores(Q, nq, ix->centroids, K, dim, S);

    int64_t n_cells = 0;
    int64_t *cells = batched ? probe_batched(ix, S, nq, p, &n_cells)
                             : probe_dense(ix, S, nq, p, subset, n_subset, has_subset, &n_cells);
    if (tr && tr->cells) {
        tr->n_cells = n_cells;
        for (int64_t i = 0; i < n_cells && i < tr->cells_cap; ++i) tr->cells[i] = cells[i];
    }
    int64_t n_cand = 0;
    int64_t *cand = get_candidates(ix, cells, n_cells, &n_cand);
    free(cells);
    if (has_subset) { /* :434-437 / :542-545 */
        int64_t *ss = (int64_t *)malloc((size_t)(n_subset + 1) * sizeof(int64_t));
        memcpy(ss, subset, (size_t)n_subset * sizeof(int64_t));
        qsort(ss, (size_t)n_subset, sizeof(int64_t), cmp_i64);
        int64_t m = 0;
        for (int64_t i = 0; i < n_cand; ++i)
            if (bsearch(&cand[i], ss, (size_t)n_subset, sizeof(int64_t), cmp_i64)) cand[m++] = cand[i];
        n_cand = m;
        free(ss);
    }
    if (n_cand == 0) { /* :439-445 */
        free(cand);
        free(S);
        return 0;
    }
    /* approximate scores, :305-324 (== :275-302 for the batched variant) */
    scored *ap = (scored *)malloc((size_t)n_cand * sizeof(scored));
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_cand; ++i) {
        int64_t d = cand[i];
        float score = 0.0f;
        for (int q = 0; q < nq; ++q) {
            float m = -INFINITY;
            const float *row = S + (size_t)q * K;
            for (int64_t t = ix->doc_offsets[d]; t < ix->doc_offsets[d + 1]; ++t) {
                float cs = row[ix->codes[t]];
                if (cs > m) m = cs;
            }
            if (m > -INFINITY) score += m;
        }
        ap[i].s = score;
        ap[i].id = d;
        ap[i].pos = i;
    }
    if (tr && tr->candidates) {
        tr->n_candidates = n_cand;
        for (int64_t i = 0; i < n_cand && i < tr->cand_cap; ++i) {
            tr->candidates[i] = cand[i];
            if (tr->approx) tr->approx[i] = ap[i].s;
        }
    }
    free(cand);
    free(S);
    qsort(ap, (size_t)n_cand, sizeof(scored), cmp_scored_desc_stable); /* :460 */
    int64_t n_top = n_cand < p->n_full_scores ? n_cand : p->n_full_scores; /* :461-465 */
    int64_t n_dec = p->n_full_scores / 4 > p->top_k ? p->n_full_scores / 4 : p->top_k; /* :468 */
    int64_t n_keep = n_top < n_dec ? n_top : n_dec;
    if (n_keep == 0) { /* :471-477 */
        free(ap);
        return 0;
    }
    /* exact scores, :481-493 */
    scored *ex = (scored *)malloc((size_t)n_keep * sizeof(scored));
#pragma omp parallel
    {
        float *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < n_keep; ++i) {
            int64_t d = ap[i].id;
            int64_t t0 = ix->doc_offsets[d], T = ix->doc_offsets[d + 1] - t0;
            if (T > cap) {
                free(buf);
                cap = T;
                buf = (float *)malloc((size_t)cap * dim * sizeof(float));
            }
            po_decompress(ix->centroids, dim, ix->nbits, ix->bucket_weights,
                          ix->residuals + (size_t)t0 * (size_t)(dim * ix->nbits / 8), ix->codes + t0,
                          T, buf);
            ex[i].s = po_maxsim(Q, nq, buf, T, dim);
            ex[i].id = d;
            ex[i].pos = i;
        }
        free(buf);
    }
    if (tr && tr->kept) {
        tr->n_kept = n_keep;
        for (int64_t i = 0; i < n_keep && i < tr->kept_cap; ++i) {
            tr->kept[i] = ex[i].id;
            if (tr->kept_exact) tr->kept_exact[i] = ex[i].s;
        }
    }
    free(ap);
    qsort(ex, (size_t)n_keep, sizeof(scored), cmp_scored_desc_stable); /* :496 */
    int64_t n_res = p->top_k < n_keep ? p->top_k : n_keep;                /* :499 */
    for (int64_t i = 0; i < n_res; ++i) {
        out_ids[i] = ex[i].id;
        out_scores[i] = ex[i].s;
    }
    free(ex);
    return n_res;
}

/* search_many_mmap, search.rs:643-675.  Queries are concatenated; q_off has n_queries+1 entries
 * (token offsets).  out_* are [n_queries][top_k]; out_counts[n_queries]. */
PO_EXPORT int po_search_batch(const po_index *ix, const float *Q, const int64_t *q_off,
                              int64_t n_queries, const po_params *p, const int64_t *subset,
                              int64_t n_subset, int has_subset, int64_t *out_ids, float *out_scores,
                              int32_t *out_counts) {
    for (int64_t b = 0; b < n_queries; ++b) {
        int nq = (int)(q_off[b + 1] - q_off[b]);
        int64_t r = po_search_one(ix, Q + (size_t)q_off[b] * ix->dim, nq, p, subset, n_subset,
                                  has_subset, out_ids + (size_t)b * p->top_k,
                                  out_scores + (size_t)b * p->top_k, NULL);
        out_counts[b] = r < 0 ? 0 : (int32_t)r; /* parallel mode: a failed query is an empty result */
    }
    return 0;
}

/* exhaustive exact MaxSim of one query against docs [d0,d1) of the decompressed corpus
 * (ground truth for recall; no reference analogue beyond maxsim.rs:270 + codec.rs:423) */
PO_EXPORT void po_exhaustive_scores(const po_index *ix, const float *Q, int nq, int64_t d0,
                                    int64_t d1, float *out) {
    int dim = ix->dim;
#pragma omp parallel
    {
        float *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t d = d0; d < d1; ++d) {
            int64_t t0 = ix->doc_offsets[d], T = ix->doc_offsets[d + 1] - t0;
            if (T > cap) {
                free(buf);
                cap = T;
                buf = (float *)malloc((size_t)cap * dim * sizeof(float));
            }
            po_decompress(ix->centroids, dim, ix->nbits, ix->bucket_weights,
                          ix->residuals + (size_t)t0 * (size_t)(dim * ix->nbits / 8), ix->codes + t0,
                          T, buf);
            out[d - d0] = po_maxsim(Q, nq, buf, T, dim);
        }
        free(buf);
    }
}

/* quantiles, utils.rs:125-149 (sort, position q*(n-1) in f64, lo*(1-w)+hi*w with w cast to f32) */
static int cmp_f32_partial(const void *a, const void *b) {
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

PO_EXPORT void po_quantiles(const float *arr, int64_t n, const double *qs, int nqs, float *out) {
    if (n == 0) {
        for (int i = 0; i < nqs; ++i) out[i] = 0.0f;
        return;
    }
    float *s = (float *)malloc((size_t)n * sizeof(float));
    memcpy(s, arr, (size_t)n * sizeof(float));
    qsort(s, (size_t)n, sizeof(float), cmp_f32_partial);
    for (int i = 0; i < nqs; ++i) {
        double idx = qs[i] * (double)(n - 1);
        int64_t lo = (int64_t)floor(idx), hi = (int64_t)ceil(idx);
        if (lo == hi) out[i] = s[lo];
        else {
            float w = (float)(idx - (double)lo);
            out[i] = s[lo] * (1.0f - w) + s[hi] * w;
        }
    }
    free(s);
}

PO_EXPORT int po_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* update path: find_outliers, update.rs:427-608                               */
/* ------------------------------------------------------------------------- */

/* squared_norm, update.rs:427-449: four partial sums, plain mul + add (Rust does not fuse) */
static float squared_norm_ref(const float *row, int dim) {
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int i = 0;
    while (i + 4 <= dim) {
        s0 += row[i] * row[i];
        s1 += row[i + 1] * row[i + 1];
        s2 += row[i + 2] * row[i + 2];
        s3 += row[i + 3] * row[i + 3];
        i += 4;
    }
    float total = s0 + s1 + s2 + s3;
    while (i < dim) {
        total += row[i] * row[i];
        ++i;
    }
    return total;
}

/* min_distance_sq_precise, update.rs:456-473 */
static float min_distance_sq_precise(const float *row, const float *C, int64_t K, int dim) {
    float m = INFINITY;
    for (int64_t c = 0; c < K; ++c) {
        double d2 = 0.0;
        for (int d = 0; d < dim; ++d) {
            double diff = (double)row[d] - (double)C[(size_t)c * dim + d];
            d2 += diff * diff;
        }
        m = fminf(m, (float)d2);
    }
    return m;
}

/* find_outliers, update.rs:490-608.  The per-(row, centroid) dot is a sequential mul+add over the
 * dimension in both the tiled and the tail loop, so the tiling does not change any value.
 * out_idx must hold n entries; returns the number of outliers (ascending row indices). */
PO_EXPORT int64_t po_find_outliers(const float *emb, int64_t n, const float *C, int64_t K, int dim,
                                   float threshold_sq, int64_t *out_idx) {
    if (n == 0 || K == 0) return 0;
    float *cn = (float *)malloc((size_t)K * sizeof(float));
    for (int64_t c = 0; c < K; ++c) cn[c] = squared_norm_ref(C + (size_t)c * dim, dim);
    uint8_t *flag = (uint8_t *)calloc((size_t)n, 1);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < n; ++r) {
        const float *row = emb + (size_t)r * dim;
        const float en = squared_norm_ref(row, dim);
        float md = INFINITY;
        for (int64_t c = 0; c < K; ++c) {
            const float *cen = C + (size_t)c * dim;
            float dot = 0.0f;
            for (int d = 0; d < dim; ++d) dot += row[d] * cen[d]; /* -ffp-contract=off: not fused */
            float dist = en + cn[c] - 2.0f * dot;
            md = fminf(md, dist);
        }
        float band = fmaxf(fabsf(threshold_sq), 1.0f) * 1e-5f; /* OUTLIER_THRESHOLD_RECHECK_REL_EPS */
        float fin = (fabsf(md - threshold_sq) <= band) ? min_distance_sq_precise(row, C, K, dim) : md;
        flag[r] = fin > threshold_sq;
    }
    int64_t m = 0;
    for (int64_t r = 0; r < n; ++r)
        if (flag[r]) out_idx[m++] = r;
    free(flag);
    free(cn);
    return m;
}

/* ------------------------------------------------------------------------- */
/* The same search over a corpus held as several doc-contiguous shards that share the centroids      */
/* (bench.py --impl reference --gpus N: the N x 1M-doc corpus of the sharded GPU run, on the CPU).   */
/* Identical in result to po_search_one on the concatenated index: S and the probe are computed     */
/* once, candidates / approximate scores per shard, then the GLOBAL cut (search.rs:460-469), exact   */
/* scores and the stable final sort (search.rs:496).  No subset support.                             */
/* ------------------------------------------------------------------------- */
PO_EXPORT int64_t po_search_sharded(const po_index *const *shards, const int64_t *bases, int n_shards,
                                    const float *Q, int nq, const po_params *p, int64_t *out_ids,
                                    float *out_scores) {
    if (n_shards < 1 || p->n_ivf_probe < 1 || p->top_k < 0) return -1;
    const po_index *ix0 = shards[0];
    int64_t K = ix0->num_centroids;
    int dim = ix0->dim;
    int batched = p->centroid_batch_size > 0 && K > p->centroid_batch_size;
    float *S = (float *)malloc((size_t)(nq > 0 ? nq : 1) * (size_t)K * sizeof(float));
    po_centroid_scores(Q, nq, ix0->centroids, K, dim, S);
    int64_t n_cells = 0;
    int64_t *cells = batched ? probe_batched(ix0, S, nq, p, &n_cells)
                             : probe_dense(ix0, S, nq, p, NULL, 0, 0, &n_cells);
    /* candidates of every shard, global ids ascending because shards are doc-contiguous and ordered */
    int64_t total = 0;
    int64_t **cand = (int64_t **)malloc((size_t)n_shards * sizeof(int64_t *));
    int64_t *ncand = (int64_t *)malloc((size_t)n_shards * sizeof(int64_t));
    for (int s = 0; s < n_shards; ++s) {
        cand[s] = get_candidates(shards[s], cells, n_cells, &ncand[s]);
        total += ncand[s];
    }
    free(cells);
    if (total == 0) {
        for (int s = 0; s < n_shards; ++s) free(cand[s]);
        free(cand);
        free(ncand);
        free(S);
        return 0;
    }
    scored *ap = (scored *)malloc((size_t)total * sizeof(scored));
    int64_t off = 0;
    for (int s = 0; s < n_shards; ++s) {
        const po_index *ix = shards[s];
        const int64_t *cs = cand[s];
        const int64_t n = ncand[s], base = bases[s];
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            int64_t d = cs[i];
            float score = 0.0f;
            for (int q = 0; q < nq; ++q) {
                float m = -INFINITY;
                const float *row = S + (size_t)q * K;
                for (int64_t t = ix->doc_offsets[d]; t < ix->doc_offsets[d + 1]; ++t) {
                    float v = row[ix->codes[t]];
                    if (v > m) m = v;
                }
                if (m > -INFINITY) score += m;
            }
            ap[off + i].s = score;
            ap[off + i].id = base + d;
            ap[off + i].pos = off + i;
        }
        off += n;
        free(cand[s]);
    }
    free(cand);
    free(ncand);
    free(S);
    qsort(ap, (size_t)total, sizeof(scored), cmp_scored_desc_stable);
    int64_t n_top = total < p->n_full_scores ? total : p->n_full_scores;
    int64_t n_dec = p->n_full_scores / 4 > p->top_k ? p->n_full_scores / 4 : p->top_k;
    int64_t n_keep = n_top < n_dec ? n_top : n_dec;
    if (n_keep == 0) {
        free(ap);
        return 0;
    }
    scored *ex = (scored *)malloc((size_t)n_keep * sizeof(scored));
#pragma omp parallel
    {
        float *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < n_keep; ++i) {
            int64_t g = ap[i].id;
            int s = n_shards - 1;
            while (s > 0 && bases[s] > g) --s;
            const po_index *ix = shards[s];
            int64_t d = g - bases[s];
            int64_t t0 = ix->doc_offsets[d], T = ix->doc_offsets[d + 1] - t0;
            if (T > cap) {
                free(buf);
                cap = T;
                buf = (float *)malloc((size_t)cap * dim * sizeof(float));
            }
            po_decompress(ix->centroids, dim, ix->nbits, ix->bucket_weights,
                          ix->residuals + (size_t)t0 * (size_t)(dim * ix->nbits / 8), ix->codes + t0, T, buf);
            ex[i].s = po_maxsim(Q, nq, buf, T, dim);
            ex[i].id = g;
            ex[i].pos = i;
        }
        free(buf);
    }
    free(ap);
    qsort(ex, (size_t)n_keep, sizeof(scored), cmp_scored_desc_stable);
    int64_t n_res = p->top_k < n_keep ? p->top_k : n_keep;
    for (int64_t i = 0; i < n_res; ++i) {
        out_ids[i] = ex[i].id;
        out_scores[i] = ex[i].s;
    }
    free(ex);
    return n_res;
}
