// k_k1tc.cuh -- a2 on the tensor cores: diagnostic stages (PB_K1_TC_DIAG).
// Part of kernels.cuh (included from there, in order; not a standalone header).
// ==========================================================================================
// a2 on the tensor cores, stage 1 (diagnostic, PB_K1_TC_DIAG=1; not on the product path yet).
// The 16-bit score table from a 3-product split-fp16 UMMA GEMM: x = xh + xl (xh = fp16(x), xl = fp16(x - xh)),
// S~ = qh.ch + qh.cl + ql.ch accumulated in fp32 in TMEM.  M = 128 centroids = TMEM lanes, N = 128 rows of the
// QS-padded query layout (row = b*QS + q), so a thread's accumulator row is a run of ST16[b][c][.] rows.
// The engine runs it next to k_centroid_scores and reports the largest code difference
// (pb_work_counters.k1_tc_max_code_diff): the measured input for the certified consumers of profiles/r01_summary.md.
// grid = ceil(K/128) CTAs, 192 threads: warps 0-3 epilogue, warp 4 bulk-copy loader, warp 5 MMA issuer.
// ==========================================================================================
// fp16 hi/lo split of `n` rows into UMMA tile order (128-row tiles, K-major core matrices); rows >= n stay zero
__global__ void k_rows_to_f16_split_tiles(const float *__restrict__ X, long long n, int dim, __half *__restrict__ Xh,
                                          __half *__restrict__ Xl) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t tile_elems = (size_t)128 * dim;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nw) {
        const size_t tbase = (size_t)(r >> 7) * tile_elems;
        const int rr = (int)(r & 127);
        for (int j = lane; j < dim; j += 32) {
            const float v = X[(size_t)r * dim + j];
            const __half h = __float2half_rn(v);
            const size_t o = tbase + (size_t)((j >> 3) * 16 + (rr >> 3)) * 64 + (rr & 7) * 8 + (j & 7);
            Xh[o] = h;
            Xl[o] = __float2half_rn(v - __half2float(h));
        }
    }
}

// the same for the query rows in the QS-padded layout (row = b*QS + q, rows q >= nq are zero)
__global__ void k_query_split_tiles(const float *__restrict__ Q, const int *__restrict__ q_off, int B, int QS, int dim,
                                    __half *__restrict__ Qh, __half *__restrict__ Ql) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t tile_elems = (size_t)128 * dim;
    const long long n = (((long long)B * QS + 127) / 128) * 128;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nw) {
        const long long b = r / QS;
        const int q = (int)(r - b * QS);
        const bool real = b < B && q < q_off[b + 1] - q_off[b];
        const size_t tbase = (size_t)(r >> 7) * tile_elems;
        const int rr = (int)(r & 127);
        for (int j = lane; j < dim; j += 32) {
            const float v = real ? Q[(size_t)(q_off[b] + q) * dim + j] : 0.0f;
            const __half h = __float2half_rn(v);
            const size_t o = tbase + (size_t)((j >> 3) * 16 + (rr >> 3)) * 64 + (rr & 7) * 8 + (j & 7);
            Qh[o] = h;
            Ql[o] = __float2half_rn(v - __half2float(h));
        }
    }
}

template <int DIM>
__global__ void __launch_bounds__(192, 1)
k_scores16_tc(const __half *__restrict__ Ch, const __half *__restrict__ Cl, long long K, const __half *__restrict__ Qh,
              const __half *__restrict__ Ql, int n_groups, int B, int QS, const int *__restrict__ q_off,
              const float2 *__restrict__ qrange, unsigned short *__restrict__ ST16, int *__restrict__ qflag) {
    extern __shared__ __align__(128) unsigned char smem_k1[];
    constexpr int KSTEPS = DIM / 16;
    constexpr uint32_t T_BYTES = 128 * DIM * 2;  // one 128-row fp16 tile
    constexpr uint32_t LBO = 16 * 128, SBO = 128;
    unsigned char *Ah = smem_k1, *Al = Ah + T_BYTES;  // this CTA's centroid tile, hi and lo
    unsigned char *Bs = Al + T_BYTES;                 // 2 stages x (hi, lo) query-row tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(Bs + 4 * T_BYTES);
    uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6, *abar = bars + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 9);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long c0 = (long long)blockIdx.x * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 128);
        }
        mbar_init(abar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (w == 4) {
        // ---------------- loader ----------------
        if (lane == 0) {
            mbar_expect_tx(abar, 2 * T_BYTES);
            bulk_g2s(Ah, reinterpret_cast<const unsigned char *>(Ch) + (size_t)blockIdx.x * T_BYTES, T_BYTES, abar);
            bulk_g2s(Al, reinterpret_cast<const unsigned char *>(Cl) + (size_t)blockIdx.x * T_BYTES, T_BYTES, abar);
            for (int g = 0; g < n_groups; ++g) {
                const int st = g & 1;
                mbar_wait(&empty[st], (uint32_t)(((g >> 1) & 1) ^ 1));
                mbar_expect_tx(&full[st], 2 * T_BYTES);
                bulk_g2s(Bs + (size_t)(2 * st) * T_BYTES, reinterpret_cast<const unsigned char *>(Qh) + (size_t)g * T_BYTES, T_BYTES, &full[st]);
                bulk_g2s(Bs + (size_t)(2 * st + 1) * T_BYTES, reinterpret_cast<const unsigned char *>(Ql) + (size_t)g * T_BYTES, T_BYTES, &full[st]);
            }
        }
    } else if (w == 5) {
        // ---------------- MMA issuer: 3 products per k-step into one fp32 accumulator ----------------
        // instruction descriptor: c = f32, a = b = f16 (format 0), K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mbar_wait(abar, 0);
        for (int g = 0; g < n_groups; ++g) {
            const int st = g & 1, acc = g & 1;
            mbar_wait(&full[st], (uint32_t)((g >> 1) & 1));
            mbar_wait(&tempty[acc], (uint32_t)(((g >> 1) & 1) ^ 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t ah = smem_u32(Ah), al = smem_u32(Al);
                const uint32_t bh = smem_u32(Bs + (size_t)(2 * st) * T_BYTES), bl = smem_u32(Bs + (size_t)(2 * st + 1) * T_BYTES);
#pragma unroll
                for (int s = 0; s < KSTEPS; ++s) {
                    const u64 dah = tc_smem_desc(ah + s * 2 * LBO, LBO, SBO), dal = tc_smem_desc(al + s * 2 * LBO, LBO, SBO);
                    const u64 dbh = tc_smem_desc(bh + s * 2 * LBO, LBO, SBO), dbl = tc_smem_desc(bl + s * 2 * LBO, LBO, SBO);
                    tc_mma_bf16(tmem_base + acc * 128, dah, dbh, idesc, s > 0 ? 1u : 0u);
                    tc_mma_bf16(tmem_base + acc * 128, dah, dbl, idesc, 1u);
                    tc_mma_bf16(tmem_base + acc * 128, dal, dbh, idesc, 1u);
                }
                tc_commit(&empty[st]);   // query tiles consumed
                tc_commit(&tfull[acc]);  // accumulators ready
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: thread = centroid row; 128 columns = 128 padded query rows ----------------
        const long long c = c0 + threadIdx.x;
        for (int g = 0; g < n_groups; ++g) {
            const int acc = g & 1;
            mbar_wait(&tfull[acc], (uint32_t)((g >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t rr[32];
                tc_ld32(tmem_base + ((uint32_t)(32 * w) << 16) + acc * 128 + cb * 32, rr);
#pragma unroll
                for (int sub = 0; sub < 4; ++sub) {
                    const long long row0 = (long long)g * 128 + cb * 32 + sub * 8;  // 8 query rows of one query (QS % 8 == 0)
                    const int b = (int)(row0 / QS), q = (int)(row0 - (long long)b * QS);
                    if (b >= B || c >= K) continue;
                    const int nq = q_off[b + 1] - q_off[b];
                    const float2 rg = qrange[b];  // (R*scale, scale)
                    uint32_t cd[8];
                    bool real_bad = false;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float t = floorf(__fmaf_rn(__uint_as_float(rr[sub * 8 + i]), rg.y, rg.x));
                        real_bad |= (q + i < nq) && !(t >= 0.0f && t <= 65535.0f);
                        cd[i] = (uint32_t)fminf(fmaxf(t, 0.0f), 65535.0f);
                    }
                    if (real_bad) atomicOr(&qflag[b], 1);
                    *reinterpret_cast<uint4 *>(ST16 + ((size_t)b * K + c) * QS + q) =
                        make_uint4(cd[0] | (cd[1] << 16), cd[2] | (cd[3] << 16), cd[4] | (cd[5] << 16), cd[6] | (cd[7] << 16));
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (w == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

// largest |a - b| over the codes of real query tokens (diagnostic)
__global__ void k_diff16(const unsigned short *__restrict__ a, const unsigned short *__restrict__ b, const int *__restrict__ q_off,
                         long long K, int QS, int *__restrict__ out_max) {
    const int bq = blockIdx.y, nq = q_off[bq + 1] - q_off[bq];
    int best = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < K * QS; i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % QS);
        if (q >= nq) continue;
        const size_t o = (size_t)bq * K * QS + i;
        best = max(best, abs((int)a[o] - (int)b[o]));
    }
    best = __reduce_max_sync(PB_FULL, best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(out_max, best);
}

// ------------------------------------------------------------------------------------------
// a2 on the tensor cores, stage 2 building block (diagnostic under PB_K1_TC_DIAG=1): exact pinned-order score
// rows for a LIST of centroids per query -- the sparse fp32 pass that will serve the consumers which need exact
// values (probe winners, cells, the a5 re-check) once the dense table comes from k_scores16_tc.
// Same FFMA2 tile as k_centroid_scores<., true>; the centroid rows are gathered with cp.async.
// out row = list position (compact = 1: OUT[b][cap][QS]) or the centroid id (compact = 0: ST[b][K][QS]).
// grid = (ceil(cap/128), B), 128 threads.
// ------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128, 2)
k_exact_rows(const float *__restrict__ Qi, const int *__restrict__ q_off, int QS, const float *__restrict__ C, long long K,
             const uint32_t *__restrict__ list, const int *__restrict__ list_n, int cap, int compact,
             float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = DIM + 4, G = DIM / 4;
    float *Vs = smem;                      // [128][LD] gathered centroid rows
    float *Qs = smem + PB_TOK_TILE * LD;   // 16 interleaved row pairs
    const int b = blockIdx.y, n = min(list_n[b], cap), i0 = blockIdx.x * PB_TOK_TILE;
    if (i0 >= n) return;
    const int nv = min(PB_TOK_TILE, n - i0);
    const uint32_t *lst = list + (size_t)b * cap + i0;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < PB_TOK_TILE * G; idx += blockDim.x) {
        const int r = idx / G, g = idx - r * G;
        if (r < nv) cp_async16(Vs + r * LD + 4 * g, C + (size_t)lst[r] * DIM + 4 * g);
        else *reinterpret_cast<float4 *>(Vs + r * LD + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int nq = q_off[b + 1] - q_off[b];
    for (int qb = 0; qb < nq; qb += PB_Q_TILE) {
        load_pairs_async<DIM>(Qs, Qi + ((size_t)b * QS + qb) * DIM, min(PB_Q_TILE, QS - qb) / 2, PB_Q_TILE / 2);
        cp_async_wait_all();
        __syncthreads();
        if (qb + 8 * w < ((nq + 7) & ~7)) {
            float acc[8][4];
            tile_dots_f2<DIM>(Qs + 4 * w * 2 * DIM, Vs + lane * LD, acc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = lane + 32 * k;
                if (i < nv) {
                    const size_t row = compact ? (size_t)b * cap + i0 + i : (size_t)b * K + lst[i];
                    float4 *dst = reinterpret_cast<float4 *>(out + row * QS + qb + 8 * w);
                    dst[0] = make_float4(acc[0][k], acc[1][k], acc[2][k], acc[3][k]);
                    dst[1] = make_float4(acc[4][k], acc[5][k], acc[6][k], acc[7][k]);
                }
            }
        }
        __syncthreads();
    }
}

// number of 32-bit words that differ between OUT[b][i][q] and ST[b][list[i]][q] (diagnostic; 0 expected)
__global__ void k_cmp_rows(const float *__restrict__ ST, const float *__restrict__ OUT, const int *__restrict__ q_off, long long K,
                           int QS, const uint32_t *__restrict__ list, const int *__restrict__ list_n, int cap,
                           int *__restrict__ mismatches) {
    const int b = blockIdx.y, n = min(list_n[b], cap), nq8 = ((q_off[b + 1] - q_off[b]) + 7) & ~7;
    int bad = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)n * QS; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / QS), q = (int)(t - (long long)i * QS);
        if (q >= nq8) continue;
        const uint32_t x = __float_as_uint(ST[((size_t)b * K + list[(size_t)b * cap + i]) * QS + q]);
        const uint32_t y = __float_as_uint(OUT[((size_t)b * cap + i) * QS + q]);
        bad += x != y;
    }
    bad = __reduce_add_sync(PB_FULL, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

// ------------------------------------------------------------------------------------------
// a2 on the tensor cores, stage 2 (PB_K1_TC=1, off by default, unmeasured): the consumers of S when the dense
// table is the 16-bit one from k_scores16_tc and exact fp32 rows exist only where k_exact_rows put them.
//   k_collect16_tc  a3: thresholds lowered by the code margin, exact selection keys from pinned-order dots
//   k_sel_list      a3: the selected centroids of a query as a list (for k_exact_rows -> their exact rows)
//   k_cells_tc      a3: k_cells whose slab-prefix scan (batched variant) ranks on the 16-bit table and settles the
//                       codes within the margin by exact dots
//   k_mark_codes    a5: bitmap of the distinct codes of the docs that get the exact re-check (k_compact turns it
//                       into the list k_exact_rows consumes)
// Both generated from their exact-table twins in k_probe.cuh, which stay untouched.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_collect16_tc(const unsigned short *__restrict__ ST16, const float *__restrict__ Q, const int *__restrict__ q_off,
               const float *__restrict__ C, int dim, int code_margin, long long K, int QS, int n_chunks,
            const uint32_t *__restrict__ tau, int cap, int *__restrict__ counts, u64 *__restrict__ list,
            int *__restrict__ fallback) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, b = blockIdx.y;
    const int chunk = blockIdx.x * 4 + w;
    if (chunk >= n_chunks || *fallback) return;
    const int GQ = QS >> 3, g = lane & (GQ - 1);
    const long long c0 = (long long)chunk * 1024;
    const int rows = (int)min(1024ll, K - c0);
    const uint4 *base = reinterpret_cast<const uint4 *>(ST16 + ((size_t)b * K + c0) * QS);
    const int total = rows * GQ;
    // this lane's 8 thresholds as packed halfwords; padding rows (tau = 65536) never match
    uint32_t t2[4], live[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        // thresholds lowered by the margin: an estimate-built code may sit up to code_margin/2 off its exact value
        uint32_t a = tau[(size_t)b * QS + 8 * g + 2 * e], c = tau[(size_t)b * QS + 8 * g + 2 * e + 1];
        if (a < 65536u) a = a > (uint32_t)code_margin ? a - (uint32_t)code_margin : 0u;
        if (c < 65536u) c = c > (uint32_t)code_margin ? c - (uint32_t)code_margin : 0u;
        t2[e] = min(a, 65535u) | (min(c, 65535u) << 16);
        live[e] = (a < 65536u ? 0xffffu : 0u) | (c < 65536u ? 0xffff0000u : 0u);
    }
    for (int i0 = lane; i0 < total; i0 += 8 * 32) {
        uint4 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (i0 + 32 * e < total) ? __ldg(base + i0 + 32 * e) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const uint32_t hx = __vcmpgeu2(v[e].x, t2[0]) & live[0], hy = __vcmpgeu2(v[e].y, t2[1]) & live[1];
            const uint32_t hz = __vcmpgeu2(v[e].z, t2[2]) & live[2], hw = __vcmpgeu2(v[e].w, t2[3]) & live[3];
            if ((hx | hy | hz | hw) == 0u || i0 + 32 * e >= total) continue;  // the common case
            const long long c = c0 + (i0 + 32 * e) / GQ;
            const uint32_t hits[4] = {hx, hy, hz, hw};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (!((hits[j >> 1] >> (16 * (j & 1))) & 1u)) continue;
                const int q = 8 * g + j;
                const int slot = atomicAdd(&counts[(size_t)b * QS + q], 1);
                if (slot < cap) {  // the exact selection key: pinned-order dot of the query token and the centroid
                    const float *qrow = Q + (size_t)(q_off[b] + q) * dim, *crow = C + (size_t)c * dim;
                    float s = 0.0f;
                    for (int d = 0; d < dim; ++d) s = __fmaf_rn(qrow[d], crow[d], s);
                    list[((size_t)b * QS + q) * cap + slot] = ((u64)score_key_asc(s) << 32) | (uint32_t)(~(uint32_t)c);
                }
                else atomicOr(fallback, 1);
            }
        }
    }
}

// grid = B, 256 threads: list[b][0..n_list) = centroid ids of the non-empty selection keys (duplicates allowed)
__global__ void k_sel_list(const u64 *__restrict__ sel, const int *__restrict__ q_off, int QS, int n, int cap,
                           uint32_t *__restrict__ list, int *__restrict__ list_n) {
    __shared__ int fill;
    const int b = blockIdx.x, nq = q_off[b + 1] - q_off[b];
    if (threadIdx.x == 0) fill = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nq * n; i += blockDim.x) {
        const u64 k = sel[(size_t)b * QS * n + i];
        if (k != 0ull) {
            const int pos = atomicAdd(&fill, 1);
            if (pos < cap) list[(size_t)b * cap + pos] = (uint32_t)(~(uint32_t)k);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) list_n[b] = min(fill, cap);
}

// grid = (CTAs, B): one warp per listed doc, one bit per distinct code
__global__ void __launch_bounds__(256)
k_mark_codes(const uint32_t *__restrict__ docs, long long stride, const int *__restrict__ n_docs,
             const uint32_t *__restrict__ ucodes, const long long *__restrict__ udoc_off, uint32_t *__restrict__ bits,
             long long W) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int n = n_docs[b];
    uint32_t *bm = bits + (size_t)b * W;
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += gridDim.x * (blockDim.x >> 5)) {
        const uint32_t d = docs[(size_t)b * stride + i];
        for (long long t = udoc_off[d] + lane; t < udoc_off[d + 1]; t += 32) {
            const uint32_t c = ucodes[t];
            atomicOr(&bm[c >> 5], 1u << (c & 31));
        }
    }
}

__global__ void __launch_bounds__(256)
k_cells_tc(const u64 *__restrict__ sel, const float *__restrict__ ST, const int *__restrict__ q_off,
        long long K, int QS, int n, int cells_cap, int has_thr, float thr, int batched,
        long long slab, uint32_t *__restrict__ cells, int *__restrict__ n_cells,
        const unsigned short *__restrict__ ST16, const float2 *__restrict__ qrange, int code_margin,
        const float *__restrict__ Q, const float *__restrict__ C, int dim) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    const int nq = q_off[b + 1] - q_off[b];
    const int total = nq * n;
    const int P = next_pow2(max(total, 1));
    u64 *s = reinterpret_cast<u64 *>(smem_raw);  // [P] sort buffer, then unique list
    int *flags = reinterpret_cast<int *>(s + P);  // [P]
    __shared__ int scan_tmp[33];
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        u64 v = ~0ull;
        if (i < total) {
            u64 k = sel[(size_t)b * QS * n + i];  // rows q < nq are the first nq*n entries
            if (k != 0ull) v = (u64)(uint32_t)(~(uint32_t)k);  // centroid id
        }
        s[i] = v;
    }
    __syncthreads();
    bitonic_sort_u64(s, P);
    // unique
    int nu = 0;
    for (int base = 0; base < P; base += blockDim.x) {
        int i = base + threadIdx.x;
        int f = (i < P && s[i] != ~0ull && (i == 0 || s[i - 1] != s[i])) ? 1 : 0;
        int tot;
        int pos = block_exclusive_scan(f, scan_tmp, &tot);
        u64 v = i < P ? s[i] : 0;
        __syncthreads();
        if (f) reinterpret_cast<uint32_t *>(flags)[nu + pos] = (uint32_t)v;  // stage ids in flags
        nu += tot;
        __syncthreads();
    }
    // move unique ids to the front of s (as u32 in the low half), flags reused below
    for (int i = threadIdx.x; i < nu; i += blockDim.x) s[i] = reinterpret_cast<uint32_t *>(flags)[i];
    __syncthreads();
    // threshold, one warp per unique centroid
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float *STb = ST + (size_t)b * K * QS;
    for (int u = w; u < nu; u += nwarps) {
        const uint32_t c = (uint32_t)s[u];
        int keep = 1;
        if (has_thr) {
            const float *row = STb + (size_t)c * QS;
            if (!batched) {
                uint32_t best = 0u;
                for (int q = lane; q < nq; q += 32) best = max(best, score_key_asc(row[q]));
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) best = max(best, __shfl_xor_sync(PB_FULL, best, m));
                // Iterator::max_by keeps the last maximum: all non-finite -> the last token's value
                float mval = best ? key_to_score(best) : (nq > 0 ? row[nq - 1] : -INFINITY);
                keep = (mval >= thr);
            } else {
                // m1 = best finite score among tokens that selected c (they entered their slab heap).
                // Non-finite scores are not tracked here: with NaN/Inf centroid scores only the
                // dense variant's threshold is reproduced exactly (DESIGN.md "Limits").
                uint32_t best = 0u;
                for (int q = lane; q < nq; q += 32) {
                    const u64 *sq = sel + ((size_t)b * QS + q) * n;
                    bool is_sel = false;
                    for (int i = 0; i < n; ++i)
                        if (sq[i] != 0ull && (uint32_t)(~(uint32_t)sq[i]) == c) is_sel = true;
                    if (is_sel) best = max(best, score_key_asc(row[q]));
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) best = max(best, __shfl_xor_sync(PB_FULL, best, m));
                float m1 = best ? key_to_score(best) : -INFINITY;
                keep = (m1 >= thr);
                if (!keep) {
                    // another token may have recorded a score >= thr for c while scanning its slab
                    const long long s0 = (long long)(c / slab) * slab;
                    for (int q = 0; q < nq && !keep; ++q) {
                        const float v = row[q];
                        const uint32_t kv = score_key_asc(v);
                        if (!(kv != 0u && v >= thr)) continue;  // finite and over the threshold
                        // entered iff fewer than n earlier slab entries are "not worse" than v
                        // only the rows of selected centroids are exact in ST here: rank the slab prefix on the 16-bit
                        // table, and settle the codes within the margin of v's own code by an exact pinned-order dot
                        const float2 rg = qrange[b];
                        const int kv16 = (int)fminf(fmaxf(floorf(__fmaf_rn(v, rg.y, rg.x)), 0.0f), 65535.0f);
                        const unsigned short *col16 = ST16 + (size_t)b * K * QS + q;
                        const float *qrow = Q + (size_t)(q_off[b] + q) * dim;
                        int cnt = 0;
                        for (long long c2 = s0 + lane; c2 < (long long)c; c2 += 32) {
                            const int cd2 = (int)col16[(size_t)c2 * QS];
                            if (cd2 > kv16 + code_margin) ++cnt;
                            else if (cd2 + code_margin >= kv16) {
                                const float *crow = C + (size_t)c2 * dim;
                                float s2 = 0.0f;
                                for (int j = 0; j < dim; ++j) s2 = __fmaf_rn(qrow[j], crow[j], s2);
                                cnt += (score_key_asc(s2) >= kv) ? 1 : 0;
                            }
                        }
#pragma unroll
                        for (int m = 16; m >= 1; m >>= 1) cnt += __shfl_xor_sync(PB_FULL, cnt, m);
                        if (cnt < n) keep = 1;
                    }
                }
            }
        }
        if (lane == 0) flags[u] = keep;
    }
    __syncthreads();
    // ordered compaction
    int outn = 0;
    for (int base = 0; base < nu; base += blockDim.x) {
        int i = base + threadIdx.x;
        int f = (i < nu) ? flags[i] : 0;
        int tot;
        int pos = block_exclusive_scan(f, scan_tmp, &tot);
        if (f && outn + pos < cells_cap) cells[(size_t)b * cells_cap + outn + pos] = (uint32_t)s[i];
        outn += tot;
    }
    if (threadIdx.x == 0) n_cells[b] = min(outn, cells_cap);
}
