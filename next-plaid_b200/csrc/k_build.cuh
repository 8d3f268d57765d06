// k_build.cuh -- index build: assignment (exact fp32 and tcgen05 certified), quantise + pack, k-means.
// Part of kernels.cuh (included from there, in order; not a standalone header).
// ==========================================================================================
// Index-build path (secondary): nearest-centroid assignment, residual quantisation
// and bit packing, Lloyd k-means.
// ==========================================================================================

// compress_into_codes (codec.rs:297-343): code = argmax_c dot(x, C_c) in the score order, the LAST
// maximum winning exact ties (Iterator::max_by).  One CTA = 64 tokens resident in shared memory,
// all centroid tiles streamed through a double-buffered 128-row tile (cp.async); 8 warps, each
// 8 tokens x 4 centroids per lane with the pinned sequential-j FMA, running best key
// (score_key << 32 | c) per token row in registers.  `bias` (optional, k-means only) is added to the
// score before ranking: argmin ||x - c||^2 == argmax (x.c - |c|^2 / 2).
template <int DIM>
__global__ void __launch_bounds__(256, 1)
k_assign(const float *__restrict__ X, long long n, const float *__restrict__ C, long long K,
         const float *__restrict__ bias, long long *__restrict__ codes_i64, uint32_t *__restrict__ codes_u32) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = DIM + 4;
    constexpr int NB = DIM <= 128 ? 2 : 1;      // the double-buffered tile does not fit at dim 256
    float *Vs0 = smem;                          // NB x [128][LD] centroid tiles
    float *Xs = smem + NB * PB_TOK_TILE * LD;   // [64][LD] tokens
    const long long x0 = (long long)blockIdx.x * 64;
    const int nx = (int)min(64ll, n - x0);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_rows_padded_async<DIM>(Xs, X + (size_t)x0 * DIM, nx, 64);
    const long long n_tiles = (K + PB_TOK_TILE - 1) / PB_TOK_TILE;
    load_rows_padded_async<DIM>(Vs0, C, (int)min((long long)PB_TOK_TILE, K), PB_TOK_TILE);
    u64 best[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) best[i] = 0ull;
    for (long long t = 0; t < n_tiles; ++t) {
        const int buf = NB == 2 ? (int)(t & 1) : 0;
        if (NB == 1 && t > 0) {
            __syncthreads();  // everyone finished with the previous tile
            const long long c1 = t * PB_TOK_TILE;
            load_rows_padded_async<DIM>(Vs0, C + (size_t)c1 * DIM, (int)min((long long)PB_TOK_TILE, K - c1), PB_TOK_TILE);
        }
        cp_async_wait_all();
        __syncthreads();  // tile t (and Xs) landed; everyone finished with tile t-1's buffer
        if (NB == 2 && t + 1 < n_tiles) {
            const long long c1 = (t + 1) * PB_TOK_TILE;
            load_rows_padded_async<DIM>(Vs0 + (buf ^ 1) * PB_TOK_TILE * LD, C + (size_t)c1 * DIM,
                                        (int)min((long long)PB_TOK_TILE, K - c1), PB_TOK_TILE);
        }
        float acc[8][4];
        tile_dots<DIM>(Xs + 8 * w * LD, Vs0 + buf * PB_TOK_TILE * LD + lane * LD, acc);
        const long long c0 = t * PB_TOK_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long c = c0 + lane + 32 * k;
            if (c < K) {
                const float bs = bias ? bias[c] : 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float sc = bias ? acc[i][k] + bs : acc[i][k];
                    const u64 key = ((u64)score_key_asc(sc) << 32) | (uint32_t)c;
                    best[i] = key >= best[i] ? key : best[i];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const u64 b = warp_max_u64(best[i]);
        const long long tok = x0 + 8 * w + i;
        if (lane == 0 && tok < n) {
            if (codes_i64) codes_i64[tok] = (long long)(uint32_t)b;
            if (codes_u32) codes_u32[tok] = (uint32_t)b;
        }
    }
}

// residual = x - C[code] (index.rs:17-40), bucket = #{cutoffs < v} (codec.rs:386), bits LSB-first into
// an MSB-first stream (codec.rs:389-395) == per value the bit-reversed bucket, first dim in the high
// bits.  One warp per token, lane = float4 group.
template <int DIM>
__global__ void __launch_bounds__(256)
k_quantize_pack(const float *__restrict__ X, long long n, const float *__restrict__ C,
                const long long *__restrict__ codes, const float *__restrict__ cutoffs, int nbits,
                uint8_t *__restrict__ packed_out, float *__restrict__ residual_out) {
    __shared__ float cut[256];
    const int ncut = (1 << nbits) - 1;
    for (int i = threadIdx.x; i < ncut; i += blockDim.x) cut[i] = cutoffs[i];
    __syncthreads();
    constexpr int G = DIM / 4;
    const int packed = DIM * nbits / 8;
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n; t += nw) {
        const float *cen = C + (size_t)codes[t] * DIM;
        uint8_t *prow = packed_out ? packed_out + (size_t)t * packed : nullptr;
        for (int g0 = 0; g0 < G; g0 += 32) {
            const int g = g0 + lane;
            uint32_t bits = 0;  // this lane's 4*nbits bits, MSB-first
            if (g < G) {
                const float4 x = reinterpret_cast<const float4 *>(X + (size_t)t * DIM)[g];
                const float4 c = reinterpret_cast<const float4 *>(cen)[g];
                float v[4] = {__fsub_rn(x.x, c.x), __fsub_rn(x.y, c.y), __fsub_rn(x.z, c.z), __fsub_rn(x.w, c.w)};
                if (residual_out) reinterpret_cast<float4 *>(residual_out + (size_t)t * DIM)[g] = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t bucket = 0;
                    for (int c2 = 0; c2 < ncut; ++c2) bucket += (v[e] > cut[c2]) ? 1u : 0u;
                    uint32_t rev = 0;
                    for (int b2 = 0; b2 < nbits; ++b2) rev |= ((bucket >> b2) & 1u) << (nbits - 1 - b2);
                    bits = (bits << nbits) | rev;
                }
            }
            if (!prow) continue;
            if (nbits == 8) {
                if (g < G) {  // 4 bytes, first dim first
                    prow[4 * g] = (uint8_t)(bits >> 24);
                    prow[4 * g + 1] = (uint8_t)(bits >> 16);
                    prow[4 * g + 2] = (uint8_t)(bits >> 8);
                    prow[4 * g + 3] = (uint8_t)bits;
                }
            } else if (nbits == 4) {
                if (g < G) {
                    prow[2 * g] = (uint8_t)(bits >> 8);
                    prow[2 * g + 1] = (uint8_t)bits;
                }
            } else if (nbits == 2) {
                if (g < G) prow[g] = (uint8_t)bits;
            } else {  // nbits == 1: two lanes share a byte
                const uint32_t other = __shfl_down_sync(PB_FULL, bits, 1);
                if (g < G && (g & 1) == 0) prow[g >> 1] = (uint8_t)((bits << 4) | (other & 15u));
            }
        }
    }
}

// ---- Lloyd k-means (kmeans.rs:261-422 wraps fastkmeans-rs 1.0.8, whose source is not in the
// reference tree: PARITY UNPINNED, statistical tests only) ----
__global__ void k_half_sqnorm(const float *__restrict__ C, long long K, int dim, float *__restrict__ bias) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < K; c += nw) {
        float p = 0.0f;
        for (int j = lane; j < dim; j += 32) {
            const float v = C[(size_t)c * dim + j];
            p = fmaf(v, v, p);
        }
        for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(PB_FULL, p, m);
        if (lane == 0) bias[c] = -0.5f * p;
    }
}

__global__ void k_accumulate(const float *__restrict__ X, long long n, int dim, const uint32_t *__restrict__ codes,
                             float *__restrict__ sums, float *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n; t += nw) {
        const uint32_t c = codes[t];
        for (int j = lane; j < dim; j += 32) atomicAdd(&sums[(size_t)c * dim + j], X[(size_t)t * dim + j]);
        if (lane == 0) atomicAdd(&counts[c], 1.0f);
    }
}

// new centroid = mean of its points; an empty cluster keeps its previous centroid
__global__ void k_update_centroids(float *__restrict__ C, long long K, int dim, const float *__restrict__ sums,
                                   const float *__restrict__ counts) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < K * dim; i += (long long)gridDim.x * blockDim.x) {
        const float cnt = counts[i / dim];
        if (cnt > 0.0f) C[i] = sums[i] / cnt;
    }
}

// row /= max(||row||, 1e-12)  (kmeans.rs:415-419)
__global__ void k_normalize_rows(float *__restrict__ C, long long K, int dim) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < K; c += nw) {
        float p = 0.0f;
        for (int j = lane; j < dim; j += 32) {
            const float v = C[(size_t)c * dim + j];
            p = fmaf(v, v, p);
        }
        for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(PB_FULL, p, m);
        const float nrm = fmaxf(sqrtf(p), 1e-12f);
        for (int j = lane; j < dim; j += 32) C[(size_t)c * dim + j] /= nrm;
    }
}

__global__ void k_gather_rows(const float *__restrict__ X, const long long *__restrict__ idx, long long K, int dim,
                              float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < K * dim; i += (long long)gridDim.x * blockDim.x)
        out[i] = X[(size_t)idx[i / dim] * dim + (i % dim)];
}

// ==========================================================================================
// tcgen05 certified filter for nearest-centroid assignment (index-build path).
//
// The exact kernel above spends 128 fp32 FMAs per (token, centroid) pair.  Here an fp16 UMMA
// (tcgen05.mma.kind::f16, fp32 accumulators in TMEM) scores every pair and the epilogue keeps the 4 best
// centroids per token.  |s_tc - s_exact| <= eps = (2^-10 + 2^-22) |x| max|c| + 2^-24 sqrt(dim) (|x| + max|c|) + 1e-5
// (two fp16 roundings per product, Cauchy-Schwarz; the absolute spacing of fp16 subnormals; fp32 accumulation
// slack), so if the 4th best tensor-core score is more than 2*eps below the best, the true argmax is among the
// first three; those are re-scored in the pinned fp32 order and ranked with the reference's tie rule.  (bf16
// operands, round 1: eps = 2^-7 |x| max|c| -- on k-means centroids of real data, where a token has several
// centroids within 0.01 of its best, nearly every token failed the certificate: 99.98 % exact fallback on the
// clustered benchmark corpus.  fp16 narrows the band 8x; values past the fp16 range become inf and are flagged.)  Tokens that cannot be certified
// (near ties, non-finite values) go through k_assign.  The result is therefore bit-identical to
// compress_into_codes_cpu while ~98 % of the arithmetic runs on the tensor cores.
//
// One CTA = 256 tokens (two UMMA M = 128 tiles sharing every 128-centroid tile), 320 threads: warps
// 0-7 epilogue (one TMEM lane = one token each), warp 8 loader (cp.async, 3-stage ring), warp 9 MMA issuer.
// Operands sit in shared memory in the canonical K-major no-swizzle layout (8 rows x 16 bytes core
// matrices; SBO = 128 B between row groups, LBO = rows/8 * 128 B between the two 8-element K
// chunks of one MMA).
// ==========================================================================================
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#define PB_TC_M 128
#define PB_TC_N 128
#define PB_TC_STAGES 3

PB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PB_DEV void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
PB_DEV void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PB_DEV void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!done);
}
PB_DEV void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier
PB_DEV void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
PB_DEV void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PB_DEV void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PB_DEV void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
PB_DEV void tc_mma_bf16(uint32_t tmem_c, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64))
PB_DEV u64 tc_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (u64)((saddr >> 4) & 0x3fffu) | ((u64)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((u64)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
PB_DEV void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// f32 rows -> fp16 (round to nearest even; the array type says bf16 for history, the bits are fp16) in UMMA tile order
// + the L2 norm of every row.
// Tile order: blocks of 128 rows, each block stored exactly as the kernel wants it in shared memory --
// K-major canonical no-swizzle layout, byte (kc*16 + r/8)*128 + (r%8)*16 + 2*e for row r, 16-byte K chunk
// kc, element e -- so one cp.async.bulk (TMA 1-D copy) moves a whole operand tile.  The array is padded
// with zero rows to a multiple of 128.
__global__ void k_rows_to_bf16(const float *__restrict__ X, long long n, int dim, __nv_bfloat16 *__restrict__ Xb,
                               float *__restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const size_t tile_elems = (size_t)128 * dim;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nw) {
        float p = 0.0f;
        const size_t tbase = (size_t)(r >> 7) * tile_elems;
        const int rr = (int)(r & 127);
        for (int j = lane; j < dim; j += 32) {
            const float v = X[(size_t)r * dim + j];
            const int kc = j >> 3, e = j & 7;
            reinterpret_cast<__half *>(Xb)[tbase + (size_t)(kc * 16 + (rr >> 3)) * 64 + (rr & 7) * 8 + e] = __float2half_rn(v);
            p = fmaf(v, v, p);
        }
        for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(PB_FULL, p, m);
        if (lane == 0) norms[r] = sqrtf(p);
    }
}

// BIAS: rank x.c + bias[c] instead of x.c (k-means: bias = -|c|^2 / 2 turns the maximum into the L2-nearest centroid);
// the encode path instantiates BIAS = false.
template <int DIM, bool BIAS>
__global__ void __launch_bounds__(320, 1)
k_assign_tc(const __nv_bfloat16 *__restrict__ Xb, long long n, const __nv_bfloat16 *__restrict__ Cb, long long K,
            float *__restrict__ top_s /* [n][4] */, uint32_t *__restrict__ top_i /* [n][4] */,
            const float *__restrict__ bias /* [ceil(K/128)*128] or NULL */) {
    // 256 tokens per CTA = two UMMA M=128 operand tiles that share every centroid tile (halves the L2
    // traffic per token); N = 128 centroids per tile; TMEM = 2 buffers x 2 halves x 128 fp32 columns.
    // warps 0-7 epilogue (warp w: token half w/4, TMEM lanes 32*(w%4)..), warp 8 loader, warp 9 MMA issuer.
    extern __shared__ __align__(1024) unsigned char smem_tc[];
    constexpr int KC = DIM / 8;            // 16-byte K chunks per row
    constexpr int KSTEPS = DIM / 16;       // UMMA K = 16 for bf16
    constexpr uint32_t A_BYTES = PB_TC_M * DIM * 2, B_BYTES = PB_TC_N * DIM * 2;   // per 128-row tile
    constexpr uint32_t LBO = (128 / 8) * 128, SBO = 128;
    unsigned char *As = smem_tc;                 // 2 tiles (token halves)
    unsigned char *Bs = smem_tc + 2 * A_BYTES;   // PB_TC_STAGES tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(Bs + PB_TC_STAGES * B_BYTES);
    uint64_t *full = bars, *empty = bars + PB_TC_STAGES, *tfull = bars + 2 * PB_TC_STAGES, *tempty = tfull + 2;
    uint64_t *abar = tempty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(abar + 1);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long x0 = (long long)blockIdx.x * (2 * PB_TC_M);
    const long long n_tiles = (K + PB_TC_N - 1) / PB_TC_N;

    if (threadIdx.x == 0) {
        for (int i = 0; i < PB_TC_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 256);
        }
        mbar_init(abar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (w == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // A tiles: this CTA's two 128-token tiles, one bulk copy each (the bf16 array is stored in tile order)
    if (threadIdx.x == 0) {
        mbar_expect_tx(abar, 2 * A_BYTES);
        bulk_g2s(As, reinterpret_cast<const unsigned char *>(Xb) + (size_t)(2 * blockIdx.x) * A_BYTES, A_BYTES, abar);
        bulk_g2s(As + A_BYTES, reinterpret_cast<const unsigned char *>(Xb) + (size_t)(2 * blockIdx.x + 1) * A_BYTES, A_BYTES, abar);
    }
    const uint32_t tmem_base = *tmem_slot;

    if (w == 8) {
        // ---------------- loader: one elected lane, one 32 KB bulk copy per centroid tile ----------------
        if (lane == 0) {
            for (long long t = 0; t < n_tiles; ++t) {
                const int st = (int)(t % PB_TC_STAGES);
                mbar_wait(&empty[st], (uint32_t)(((t / PB_TC_STAGES) & 1) ^ 1));
                mbar_expect_tx(&full[st], B_BYTES);
                bulk_g2s(Bs + (size_t)st * B_BYTES, reinterpret_cast<const unsigned char *>(Cb) + (size_t)t * B_BYTES, B_BYTES,
                         &full[st]);
            }
        }
    } else if (w == 9) {
        // ---------------- MMA issuer ----------------
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=f16 [7,10)=0,
        // b=f16 [10,13)=0, both K-major, N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(PB_TC_N >> 3) << 17) | ((uint32_t)(PB_TC_M >> 4) << 24);
        mbar_wait(abar, 0);  // token tiles landed
        for (long long t = 0; t < n_tiles; ++t) {
            const int st = (int)(t % PB_TC_STAGES), acc = (int)(t & 1);
            mbar_wait(&full[st], (uint32_t)((t / PB_TC_STAGES) & 1));
            mbar_wait(&tempty[acc], (uint32_t)(((t >> 1) & 1) ^ 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs + (size_t)st * B_BYTES);
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int s = 0; s < KSTEPS; ++s) {
                        const u64 ad = tc_smem_desc(a0 + half * A_BYTES + s * 2 * LBO, LBO, SBO);
                        const u64 bd = tc_smem_desc(b0 + s * 2 * LBO, LBO, SBO);
                        tc_mma_bf16(tmem_base + acc * 256 + half * PB_TC_N, ad, bd, idesc, s > 0 ? 1u : 0u);
                    }
                tc_commit(&empty[st]);   // B tile consumed
                tc_commit(&tfull[acc]);  // accumulators ready
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: thread = token row, running top-4 over all centroids ----------------
        float s0 = -INFINITY, s1 = -INFINITY, s2 = -INFINITY, s3 = -INFINITY;
        uint32_t i0 = 0xffffffffu, i1 = 0xffffffffu, i2 = 0xffffffffu, i3 = 0xffffffffu;
        const int half = w >> 2, lg = w & 3;
        for (long long t = 0; t < n_tiles; ++t) {
            const int acc = (int)(t & 1);
            mbar_wait(&tfull[acc], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            const long long c0 = t * PB_TC_N;
            const bool edge = c0 + PB_TC_N > K;  // the (zero-filled) columns past K must not be ranked
#pragma unroll 1
            for (int cb = 0; cb < PB_TC_N / 32; ++cb) {
                uint32_t rr[32];
                tc_ld32(tmem_base + ((uint32_t)(32 * lg) << 16) + acc * 256 + half * PB_TC_N + cb * 32, rr);
                if (BIAS) {  // warp-uniform addresses: one broadcast load per column
#pragma unroll
                    for (int j = 0; j < 32; ++j) rr[j] = __float_as_uint(__uint_as_float(rr[j]) + __ldg(bias + c0 + cb * 32 + j));
                }
                // one max tree per 32 columns; the insertion path runs only when the batch can matter
                float m = __uint_as_float(rr[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(rr[j]));
                if (m > s3 || edge) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float v = __uint_as_float(rr[j]);
                        const uint32_t c = (uint32_t)(c0 + cb * 32 + j);
                        if (v > s3 && c < (uint32_t)K) {  // NaN never enters
                            if (v > s0) { s3 = s2; i3 = i2; s2 = s1; i2 = i1; s1 = s0; i1 = i0; s0 = v; i0 = c; }
                            else if (v > s1) { s3 = s2; i3 = i2; s2 = s1; i2 = i1; s1 = v; i1 = c; }
                            else if (v > s2) { s3 = s2; i3 = i2; s2 = v; i2 = c; }
                            else { s3 = v; i3 = c; }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
        const long long tok = x0 + half * PB_TC_M + 32 * lg + lane;
        if (tok < n) {
            reinterpret_cast<float4 *>(top_s)[tok] = make_float4(s0, s1, s2, s3);
            reinterpret_cast<uint4 *>(top_i)[tok] = make_uint4(i0, i1, i2, i3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (w == 9) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// certification + exact re-scoring of the shortlist; uncertified tokens are flagged for k_assign
__global__ void k_assign_certify(const float *__restrict__ X, long long n, int dim, const float *__restrict__ C,
                                 const float *__restrict__ xnorm, float cmax, int c_finite,
                                 const float *__restrict__ top_s, const uint32_t *__restrict__ top_i,
                                 long long *__restrict__ codes, int *__restrict__ n_fallback,
                                 long long *__restrict__ fallback_list) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const float4 s = reinterpret_cast<const float4 *>(top_s)[t];
        const uint4 id = reinterpret_cast<const uint4 *>(top_i)[t];
        const float xn = xnorm[t];
        // (2^-10 + 2^-22) |x| max|c| (fp16 unit roundoff 2^-11, twice) + subnormal spacing + accumulation slack
        const float eps = 0.00097680f * xn * cmax + 5.9604645e-8f * sqrtf((float)dim) * (xn + cmax) + 1e-5f;
        // certified iff everything is finite, four candidates exist and the 4th is out of the band
        bool ok = c_finite && xn < 1e18f && (s.x > -1e30f) && (s.x < 1e30f) && id.w != 0xffffffffu && (s.w < s.x - 2.0f * eps);
        if (ok) {
            const float sv[3] = {s.x, s.y, s.z};
            const uint32_t iv[3] = {id.x, id.y, id.z};
            u64 best = 0ull;
            for (int j = 0; j < 3; ++j) {
                if (sv[j] < s.x - 2.0f * eps) continue;  // cannot be the argmax
                const float *c = C + (size_t)iv[j] * dim;
                const float *x = X + (size_t)t * dim;
                float acc = 0.0f;
                for (int d = 0; d < dim; ++d) acc = __fmaf_rn(x[d], c[d], acc);  // pinned order
                const u64 key = ((u64)score_key_asc(acc) << 32) | iv[j];
                best = key >= best ? key : best;
            }
            codes[t] = (long long)(uint32_t)best;
        } else {
            const int slot = atomicAdd(n_fallback, 1);
            fallback_list[slot] = t;
        }
    }
}

__global__ void k_gather_rows_i64(const float *__restrict__ X, const long long *__restrict__ idx, long long m, int dim,
                                  float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m * dim; i += (long long)gridDim.x * blockDim.x)
        out[i] = X[(size_t)idx[i / dim] * dim + (i % dim)];
}
__global__ void k_scatter_codes(const long long *__restrict__ src, const long long *__restrict__ idx, long long m,
                                long long *__restrict__ dst) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x)
        dst[idx[i]] = src[i];
}

// ------------------------------------------------------------------------------------------
// Inverted file from the codes (index.rs:850-873: code -> sorted unique doc ids).  The per-doc distinct code lists
// (ucodes, built by k_unique_codes for a5) already hold the (doc, code) relation; k_ivf_pairs emits one 64-bit key
// (code << 32 | doc) per distinct pair -- entries equal to their predecessor are the padding -- a radix sort on
// the keys orders them by centroid, then doc id; duplicates (docs longer than PB_UCODE_MAX keep a raw code list) are
// dropped by a unique pass.  Offsets = one lower_bound per centroid.
// ------------------------------------------------------------------------------------------
__global__ void k_ivf_pairs(const uint32_t *__restrict__ ucodes, const long long *__restrict__ udoc_off, long long D,
                            u64 *__restrict__ keys, unsigned long long *__restrict__ n_keys) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long d = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < D; d += nw) {
        const long long t0 = udoc_off[d], t1 = udoc_off[d + 1];
        for (long long t = t0; t < t1; t += 32) {
            const long long i = t + lane;
            const bool real = i < t1 && (i == t0 || ucodes[i] != ucodes[i - 1]);
            const unsigned bal = __ballot_sync(PB_FULL, real);
            unsigned long long base = 0;
            if (lane == 0 && bal) base = atomicAdd(n_keys, (unsigned long long)__popc(bal));
            base = shfl_u64(base, 0);
            if (real) keys[base + __popc(bal & ((1u << lane) - 1u))] = ((u64)ucodes[i] << 32) | (uint32_t)d;
        }
    }
}

__global__ void k_ivf_from_keys(const u64 *__restrict__ keys, long long n, uint32_t *__restrict__ ivf) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        ivf[i] = (uint32_t)keys[i];
}

// ivf_off[c] = first position whose key has centroid >= c (c = 0..K)
__global__ void k_ivf_offsets(const u64 *__restrict__ keys, long long n, long long K, long long *__restrict__ ivf_off) {
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c <= K; c += (long long)gridDim.x * blockDim.x) {
        const u64 want = (u64)c << 32;
        long long lo = 0, hi = n;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (keys[mid] < want) lo = mid + 1; else hi = mid;
        }
        ivf_off[c] = lo;
    }
}

// export in the reference's dtypes: ivf.npy <i8 (global doc ids), ivf_lengths.npy <i4
__global__ void k_ivf_export(const uint32_t *__restrict__ ivf, const long long *__restrict__ ivf_off, long long n, long long K,
                             long long doc_id_base, long long *__restrict__ out_ivf, int *__restrict__ out_len) {
    const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (out_ivf)
        for (long long i = i0; i < n; i += stride) out_ivf[i] = (long long)ivf[i] + doc_id_base;
    if (out_len)
        for (long long c = i0; c < K; c += stride) out_len[c] = (int)(ivf_off[c + 1] - ivf_off[c]);
}

// codec training (index.rs:240-258): L2 norm of every residual row; per-dimension mean of |residual|
__global__ void k_residual_stats(const float *__restrict__ R, long long n, int dim, float *__restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nw) {
        float p = 0.0f;
        for (int j = lane; j < dim; j += 32) {
            const float v = R[(size_t)r * dim + j];
            p = fmaf(v, v, p);
        }
        for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(PB_FULL, p, m);
        if (lane == 0) norms[r] = sqrtf(p);
    }
}
__global__ void k_column_abs_mean(const float *__restrict__ R, long long n, int dim, float *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    double acc = 0.0;  // one thread per dimension, rows in order: deterministic
    for (long long r = 0; r < n; ++r) acc += (double)fabsf(R[(size_t)r * dim + j]);
    out[j] = n > 0 ? (float)(acc / (double)n) : 0.0f;
}

// out[i] = stage[0][i] + stage[1][i] + ... in rank order (the in-process all-reduce: identical bits on every rank)
__global__ void k_sum_ranks(const float *__restrict__ stage, int world, long long count, float *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        float acc = stage[i];
        for (int r = 1; r < world; ++r) acc += stage[(size_t)r * count + i];
        out[i] = acc;
    }
}

// k-means assignment from the tensor-core shortlist: the best (score + bias) wins, no exact re-score (the iteration is
// parity-unpinned; bf16 rounding only moves points that sit between two centroids)
__global__ void k_take_top1(const uint32_t *__restrict__ top_i, long long n, uint32_t *__restrict__ codes) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const uint32_t c = top_i[4 * t];
        codes[t] = c == 0xffffffffu ? 0u : c;
    }
}
// bias[c] = -|c|^2 / 2 for c < K, 0 for the padding columns of the last tile
__global__ void k_half_sqnorm_padded(const float *__restrict__ C, long long K, long long Kpad, int dim, float *__restrict__ bias) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < Kpad; c += nw) {
        float p = 0.0f;
        if (c < K)
            for (int j = lane; j < dim; j += 32) {
                const float v = C[(size_t)c * dim + j];
                p = fmaf(v, v, p);
            }
        for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(PB_FULL, p, m);
        if (lane == 0) bias[c] = -0.5f * p;
    }
}
