// k_outliers.cuh -- update_centroids numeric kernel: find_outliers.
// Part of kernels.cuh (included from there, in order; not a standalone header).
// ==========================================================================================
// update path: find_outliers, update.rs:490-608 -- rows whose minimum squared L2 distance
// to any centroid exceeds threshold_sq.  The reference spells its own loops out (no third-party GEMM),
// so the arithmetic is reproduced operation for operation: squared_norm with four partial sums and
// plain mul + add (update.rs:427-449), the dot as a sequential mul + add over the dimension,
// dist = (|x|^2 + |c|^2) - 2*dot, f32::min, and the f64 re-check of rows within 1e-5 of the threshold
// (update.rs:456-473, :592-599).
// ==========================================================================================
PB_DEV float squared_norm_ref(const float *__restrict__ row, int dim) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = 0;
    for (; i + 4 <= dim; i += 4) {
        s0 = __fadd_rn(s0, __fmul_rn(row[i], row[i]));
        s1 = __fadd_rn(s1, __fmul_rn(row[i + 1], row[i + 1]));
        s2 = __fadd_rn(s2, __fmul_rn(row[i + 2], row[i + 2]));
        s3 = __fadd_rn(s3, __fmul_rn(row[i + 3], row[i + 3]));
    }
    float total = __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
    for (; i < dim; ++i) total = __fadd_rn(total, __fmul_rn(row[i], row[i]));
    return total;
}

__global__ void k_squared_norms_ref(const float *__restrict__ X, long long n, int dim, float *__restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
        out[r] = squared_norm_ref(X + (size_t)r * dim, dim);
}

// unfused twin of tile_dots: acc = acc + q*v with separate roundings, j ascending
template <int DIM>
PB_DEV void tile_dots_unfused(const float *__restrict__ Qs, const float *__restrict__ Vs, float (&acc)[8][4]) {
    constexpr int LD = DIM + 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0.0f;
#pragma unroll 2
    for (int j = 0; j < DIM; j += 4) {
        float4 q[8], v[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = *reinterpret_cast<const float4 *>(Qs + i * LD + j);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4 *>(Vs + (32 * k) * LD + j);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float a = acc[i][k];
                a = __fadd_rn(a, __fmul_rn(q[i].x, v[k].x));
                a = __fadd_rn(a, __fmul_rn(q[i].y, v[k].y));
                a = __fadd_rn(a, __fmul_rn(q[i].z, v[k].z));
                a = __fadd_rn(a, __fmul_rn(q[i].w, v[k].w));
                acc[i][k] = a;
            }
    }
}

// min over centroids of (|x|^2 + |c|^2) - 2*dot; same tiling as k_assign (64 rows resident, centroid
// tiles streamed).  grid = ceil(n/64), 256 threads.
template <int DIM>
__global__ void __launch_bounds__(256, 1)
k_min_dist(const float *__restrict__ X, long long n, const float *__restrict__ xnorm, const float *__restrict__ C,
           long long K, const float *__restrict__ cnorm, float *__restrict__ min_dist) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LD = DIM + 4;
    constexpr int NB = DIM <= 128 ? 2 : 1;
    float *Vs0 = smem;
    float *Xs = smem + NB * PB_TOK_TILE * LD;
    const long long x0 = (long long)blockIdx.x * 64;
    const int nx = (int)min(64ll, n - x0);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    load_rows_padded_async<DIM>(Xs, X + (size_t)x0 * DIM, nx, 64);
    const long long n_tiles = (K + PB_TOK_TILE - 1) / PB_TOK_TILE;
    load_rows_padded_async<DIM>(Vs0, C, (int)min((long long)PB_TOK_TILE, K), PB_TOK_TILE);
    float best[8], en[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        best[i] = INFINITY;
        const long long r = x0 + 8 * w + i;
        en[i] = r < n ? xnorm[r] : 0.0f;
    }
    for (long long t = 0; t < n_tiles; ++t) {
        const int buf = NB == 2 ? (int)(t & 1) : 0;
        if (NB == 1 && t > 0) {
            __syncthreads();
            const long long c1 = t * PB_TOK_TILE;
            load_rows_padded_async<DIM>(Vs0, C + (size_t)c1 * DIM, (int)min((long long)PB_TOK_TILE, K - c1), PB_TOK_TILE);
        }
        cp_async_wait_all();
        __syncthreads();
        if (NB == 2 && t + 1 < n_tiles) {
            const long long c1 = (t + 1) * PB_TOK_TILE;
            load_rows_padded_async<DIM>(Vs0 + (buf ^ 1) * PB_TOK_TILE * LD, C + (size_t)c1 * DIM,
                                        (int)min((long long)PB_TOK_TILE, K - c1), PB_TOK_TILE);
        }
        float acc[8][4];
        tile_dots_unfused<DIM>(Xs + 8 * w * LD, Vs0 + buf * PB_TOK_TILE * LD + lane * LD, acc);
        const long long c0 = t * PB_TOK_TILE;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long c = c0 + lane + 32 * k;
            if (c < K) {
                const float cn = cnorm[c];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float dist = __fsub_rn(__fadd_rn(en[i], cn), __fmul_rn(2.0f, acc[i][k]));
                    best[i] = fminf(best[i], dist);  // f32::min: a NaN operand is ignored
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float b = best[i];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) b = fminf(b, __shfl_xor_sync(PB_FULL, b, m));
        const long long r = x0 + 8 * w + i;
        if (lane == 0 && r < n) min_dist[r] = b;
    }
}

// decision per row; rows inside the re-check band get min_distance_sq_precise (f64, d ascending)
// from a whole warp (lanes split the centroids).  grid-stride, one warp per row.
__global__ void __launch_bounds__(256)
k_outlier_decide(const float *__restrict__ X, long long n, int dim, const float *__restrict__ C, long long K,
                 const float *__restrict__ min_dist, float threshold_sq, uint8_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const float band = __fmul_rn(fmaxf(fabsf(threshold_sq), 1.0f), 1e-5f);
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += nw) {
        float md = min_dist[r];
        if (fabsf(__fsub_rn(md, threshold_sq)) <= band) {
            const float *row = X + (size_t)r * dim;
            float m = INFINITY;
            for (long long c = lane; c < K; c += 32) {
                const float *cen = C + (size_t)c * dim;
                double d2 = 0.0;
                for (int d = 0; d < dim; ++d) {
                    const double diff = __dsub_rn((double)row[d], (double)cen[d]);
                    d2 = __dadd_rn(d2, __dmul_rn(diff, diff));
                }
                m = fminf(m, (float)d2);
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) m = fminf(m, __shfl_xor_sync(PB_FULL, m, o));
            md = m;
        }
        if (lane == 0) flags[r] = md > threshold_sq ? 1 : 0;
    }
}
