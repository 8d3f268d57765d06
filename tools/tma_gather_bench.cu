// tma_gather_bench.cu -- how fast can an SM pull random 64-byte (or 32-byte) rows of an L2-resident table?
// The a5 first pass (k_approx16) gathers one 2*QS-byte row of a 16 MiB score table per (candidate, distinct code):
// 374 M rows per step at BASELINE config B.  Through the LSU that costs one L1 wavefront per row and 2 L2 sectors;
// this tool measures the same gather (i) through the LSU exactly as k_approx16 issues it and (ii) through the TMA
// (`cp.async.bulk.tensor.2d ... tile::gather4`: 4 rows per instruction, L2 -> shared memory, no L1 tag / wavefront),
// both followed by the packed u16 max-reduce, so the a5 design question "would gather4 lift the ceiling?" has a number.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/tma_gather_bench.cu -o gpurun_out/tma_gather_bench
//   gpurun_out/tma_gather_bench [row_bytes=64] [log2_rows=18] [million_gathers=256] [box_rows=1]
// Prints one JSON line per mode.  box_rows: second box dimension of the tensor map (1 is what tile::gather4 wants; the
// argument exists to probe).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, uint64_t *bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}

// ---- LSU: the k_approx16 access pattern.  A warp takes 64 indices per step, a row is (row_bytes/16) lanes x 16 bytes.
template <int ROW_BYTES>
__global__ void __launch_bounds__(256, 4)
k_lsu(const char *__restrict__ table, const uint32_t *__restrict__ idx, long long n_per_warp, uint32_t idx_mask,
      uint32_t *__restrict__ out) {
    constexpr int LPR = ROW_BYTES / 16;     // lanes per row
    constexpr int RPI = 32 / LPR;           // rows per load instruction
    const int lane = threadIdx.x & 31, r = lane / LPR, sl = lane % LPR;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    long long base = warp * n_per_warp;
    for (long long t = 0; t < n_per_warp; t += 64) {
        const uint32_t c0 = idx[(base + t + lane) & idx_mask], c1 = idx[(base + t + 32 + lane) & idx_mask];
        uint4 v[64 / RPI];
#pragma unroll
        for (int e = 0; e < 64 / RPI; ++e) {
            const int src = e * RPI + r;
            const uint32_t c = __shfl_sync(0xffffffffu, src < 32 ? c0 : c1, src & 31);
            v[e] = *reinterpret_cast<const uint4 *>(table + (size_t)c * ROW_BYTES + 16 * sl);
        }
#pragma unroll
        for (int e = 0; e < 64 / RPI; ++e) {
            m0 = __vmaxu2(m0, v[e].x);
            m1 = __vmaxu2(m1, v[e].y);
            m2 = __vmaxu2(m2, v[e].z);
            m3 = __vmaxu2(m3, v[e].w);
        }
    }
    out[(size_t)warp * 32 + lane] = m0 ^ m1 ^ m2 ^ m3;
}

// ---- TMA gather4: warp 0 of the CTA issues, warps 1..7 reduce.  STAGES ring buffers of ROWS_PER_STAGE rows each.
template <int ROW_BYTES, int STAGES, int ROWS_PER_STAGE>
__global__ void __launch_bounds__(256, 1)
k_tma(const __grid_constant__ CUtensorMap map, const uint32_t *__restrict__ idx, long long n_per_cta, uint32_t idx_mask,
      uint32_t *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *buf = smem;                                        // [STAGES][ROWS_PER_STAGE][ROW_BYTES]
    uint64_t *full = reinterpret_cast<uint64_t *>(buf + (size_t)STAGES * ROWS_PER_STAGE * ROW_BYTES);
    uint64_t *empty = full + STAGES;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 7);   // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long base = (long long)blockIdx.x * n_per_cta;
    const long long n_stages = n_per_cta / ROWS_PER_STAGE;
    if (w == 0) {
        for (long long it = 0; it < n_stages; ++it) {
            const int s = (int)(it % STAGES);
            mbar_wait(&empty[s], (uint32_t)(((it / STAGES) & 1) ^ 1));
            if (lane == 0) mbar_expect_tx(&full[s], ROWS_PER_STAGE * ROW_BYTES);
            __syncwarp();
            // every lane issues its share of the stage's gather4 instructions (4 rows each)
            for (int g = lane; g < ROWS_PER_STAGE / 4; g += 32) {
                const long long p = base + it * ROWS_PER_STAGE + 4 * g;
                const uint4 r4 = *reinterpret_cast<const uint4 *>(idx + (p & idx_mask));
                tma_gather4(buf + ((size_t)s * ROWS_PER_STAGE + 4 * g) * ROW_BYTES, &map, &full[s], 0, (int)r4.x, (int)r4.y,
                            (int)r4.z, (int)r4.w);
            }
        }
    } else {
        uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        constexpr int LPR = ROW_BYTES / 16;
        const int cw = w - 1;  // 0..6
        for (long long it = 0; it < n_stages; ++it) {
            const int s = (int)(it % STAGES);
            mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
            const unsigned char *st = buf + (size_t)s * ROWS_PER_STAGE * ROW_BYTES;
            // 16-byte pieces of the stage, spread over the 7 consumer warps
            for (int piece = cw * 32 + lane; piece < ROWS_PER_STAGE * LPR; piece += 7 * 32) {
                const uint4 v = *reinterpret_cast<const uint4 *>(st + (size_t)piece * 16);
                // keep the per-column structure of the real kernel: piece % LPR is the column group
                if ((piece % LPR) & 1) {
                    m2 = __vmaxu2(m2, v.x ^ v.z);
                    m3 = __vmaxu2(m3, v.y ^ v.w);
                } else {
                    m0 = __vmaxu2(m0, v.x ^ v.z);
                    m1 = __vmaxu2(m1, v.y ^ v.w);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        out[((size_t)blockIdx.x * 7 + cw) * 32 + lane] = m0 ^ m1 ^ m2 ^ m3;
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROW_BYTES>
static int run(int log2_rows, long long mgathers, int box_rows) {
    const long long K = 1ll << log2_rows;
    const size_t table_bytes = (size_t)K * ROW_BYTES;
    char *table;
    CK(cudaMalloc(&table, table_bytes));
    std::vector<uint32_t> h((size_t)table_bytes / 4);
    uint64_t s = 88172645463325252ull;
    for (auto &x : h) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        x = (uint32_t)s;
    }
    CK(cudaMemcpy(table, h.data(), table_bytes, cudaMemcpyHostToDevice));
    const uint32_t n_idx = 1u << 24;
    std::vector<uint32_t> hi(n_idx);
    for (auto &x : hi) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        x = (uint32_t)(s >> 20) & (uint32_t)(K - 1);
    }
    uint32_t *idx, *out;
    CK(cudaMalloc(&idx, (size_t)n_idx * 4));
    CK(cudaMemcpy(idx, hi.data(), (size_t)n_idx * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, 1 << 24));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const long long total = mgathers * 1000000ll;
    // ---- LSU ----
    {
        const int ctas = sms * 8, warps = ctas * 8;
        const long long per_warp = (total / warps) / 64 * 64;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            k_lsu<ROW_BYTES><<<ctas, 256>>>(table, idx, per_warp, n_idx - 1, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaGetLastError());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double rows = (double)per_warp * warps;
        printf("{\"mode\": \"lsu\", \"row_bytes\": %d, \"table_MiB\": %.1f, \"rows\": %.0f, \"ms\": %.3f, \"Grows_per_s\": %.1f, "
               "\"GB_per_s\": %.0f, \"rows_per_clk_per_sm\": %.3f}\n",
               ROW_BYTES, table_bytes / 1048576.0, rows, ms, rows / ms / 1e6, rows * ROW_BYTES / ms / 1e6,
               rows / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
    }
    // ---- TMA gather4 ----
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn) {
            printf("{\"mode\": \"tma_gather4\", \"error\": \"cuTensorMapEncodeTiled not available\"}\n");
            return 0;
        }
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)(ROW_BYTES / 2), (cuuint64_t)K};
        const cuuint64_t gstride[1] = {(cuuint64_t)ROW_BYTES};
        const cuuint32_t box[2] = {(cuuint32_t)(ROW_BYTES / 2), (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeTiled)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, table, gdim, gstride, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("{\"mode\": \"tma_gather4\", \"error\": \"cuTensorMapEncodeTiled returned %d (box_rows %d)\"}\n", (int)r, box_rows);
            return 0;
        }
        constexpr int STAGES = 4, RPS = 256;
        const size_t sm = (size_t)STAGES * RPS * ROW_BYTES + 2 * STAGES * 8 + 64;
        auto kern = k_tma<ROW_BYTES, STAGES, RPS>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        const int ctas = sms;
        const long long per_cta = (total / ctas) / RPS * RPS;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            kern<<<ctas, 256, sm>>>(map, idx, per_cta, n_idx - 1, out);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) {
                printf("{\"mode\": \"tma_gather4\", \"error\": \"%s (box_rows %d)\"}\n", cudaGetErrorString(e), box_rows);
                return 0;
            }
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double rows = (double)per_cta * ctas;
        printf("{\"mode\": \"tma_gather4\", \"row_bytes\": %d, \"box_rows\": %d, \"stages\": %d, \"rows_per_stage\": %d, \"rows\": %.0f, "
               "\"ms\": %.3f, \"Grows_per_s\": %.1f, \"GB_per_s\": %.0f, \"rows_per_clk_per_sm\": %.3f}\n",
               ROW_BYTES, box_rows, STAGES, RPS, rows, ms, rows / ms / 1e6, rows * ROW_BYTES / ms / 1e6,
               rows / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
    }
    return 0;
}

int main(int argc, char **argv) {
    const int row_bytes = argc > 1 ? atoi(argv[1]) : 64;
    const int log2_rows = argc > 2 ? atoi(argv[2]) : 18;
    const long long mg = argc > 3 ? atoll(argv[3]) : 256;
    const int box_rows = argc > 4 ? atoi(argv[4]) : 1;
    if (row_bytes == 64) return run<64>(log2_rows, mg, box_rows);
    if (row_bytes == 32) return run<32>(log2_rows, mg, box_rows);
    if (row_bytes == 128) return run<128>(log2_rows, mg, box_rows);
    printf("{\"error\": \"row_bytes must be 32, 64 or 128\"}\n");
    return 1;
}
