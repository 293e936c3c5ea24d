// libat_b200: the use_convolution branch of stages 2-3.
//
// With config.use_convolution the reference expands every 64-value mel frame through nn.Conv1d(1 -> num_kernels,
// kernel_size, padding = kernel_size // 2) along the mel axis and clusters / tokenizes the (n_mels * num_kernels)-value rows
// (processors/cluster_creator.py:28-34,68-81; processors/spec_tokenizer.py:92-104,115-121).  Two kernels:
//   k_conv_expand   rows (n, n_mels) -> (n, n_mels * num_kernels), out[i, m * Kc + c] = b[c] + sum_t w[c][t] x[i, m + t - pad]
//                   (the layout of conv_output.transpose(1, 2).reshape(-1, Kc * n_mels))
//   k_assign_gemm   exact fp32 nearest-centroid search for rows wider than the register-resident kernel covers (d > 128):
//                   a 128 x 128 tile of inner products per block (8 x 8 per thread in packed fp32, K chunks of 8 through
//                   double-buffered shared memory), |x|^2 + |c|^2 - 2 <x, c> clamped at 0, strict '<' in ascending column
//                   order (lowest index wins ties).
#include "at_index.cuh"

namespace at {

// Block = kc * (256 / kc) threads: thread t owns channel c = t % kc (its ks weights and bias live in registers for the
// whole launch) and walks the mel bins m = t / kc, + threads / kc, ... of one row after the other, so consecutive threads
// write consecutive floats of the output row (coalesced) and nothing is divided inside the loops.  CONV_ROWS rows are
// staged in shared memory per block step (zero padded at both ends: no bounds tests on the taps).
constexpr int CONV_ROWS = 8, CONV_MAX_KS = 15;
template <int KS>   // taps (compile time: the tap loop is exactly KS FMAs, no predicates)
__global__ void __launch_bounds__(256) k_conv_expand(const float *__restrict__ x, int64_t n, int n_mels,
                                                     const float *__restrict__ w, const float *__restrict__ b, int kc,
                                                     float *__restrict__ out) {
    constexpr int ks = KS;
    extern __shared__ float s_x[];   // CONV_ROWS rows of (pad + n_mels + pad) floats
    const int pad = ks / 2, stride = n_mels + 2 * pad;
    const int c = (int)threadIdx.x % kc, m0 = (int)threadIdx.x / kc, mstep = (int)blockDim.x / kc;
    float wr[KS];
#pragma unroll
    for (int t = 0; t < KS; t++) wr[t] = w[c * ks + t];
    const float bias = b[c];
    const int64_t d_out = (int64_t)n_mels * kc;
    for (int64_t r0 = (int64_t)blockIdx.x * CONV_ROWS; r0 < n; r0 += (int64_t)gridDim.x * CONV_ROWS) {
        const int rows = (int)min((int64_t)CONV_ROWS, n - r0);
        __syncthreads();   // the previous step's rows have been consumed
        for (int i = threadIdx.x; i < rows * stride; i += blockDim.x) {
            const int r = i / stride, p = i - r * stride - pad;
            s_x[i] = (p >= 0 && p < n_mels) ? x[(r0 + r) * n_mels + p] : 0.f;
        }
        __syncthreads();
        for (int r = 0; r < rows; r++) {
            const float *xr = s_x + r * stride;   // xr[m + t] = x[row, m + t - pad]
            float *orow = out + (r0 + r) * d_out + c;
            for (int m = m0; m < n_mels; m += mstep) {
                float acc = bias;
#pragma unroll
                for (int t = 0; t < KS; t++) acc = fmaf(wr[t], xr[m + t], acc);   // ascending taps (same rounding as before)
                orow[(int64_t)m * kc] = acc;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_assign_gemm: exact fp32 nearest-centroid search for wide rows.  A block owns 128 rows and sweeps the centroids in tiles
// of 128; thread (ty, tx) of a 16 x 16 grid accumulates the 8 x 8 inner products of rows {4 ty .. 4 ty + 3, 64 + 4 ty ..}
// and columns {4 tx .. 4 tx + 3, 64 + 4 tx ..}.  K chunks of 8 go through double-buffered shared memory (one float4 of each
// operand per thread and chunk, fetched one chunk ahead into registers).  The products are packed fp32 (fma.rn.f32x2: two
// adjacent columns per instruction): the row operand is stored DUPLICATED in shared memory ((a, a) pairs) so that a 128-bit
// load yields two ready-made register pairs, the centroid operand's adjacent columns are pairs as they are -- 32 FFMA2 + 6
// LDS.128 per k instead of 64 FFMA + 4 LDS.128.  Each inner product is still one sequential fp32 FMA chain over k.
// |x|^2 + |c|^2 - 2 <x, c> clamped at 0, strict '<' in ascending column order (the lowest index wins exact ties).
constexpr int GM = 128, GN = 128, GK = 8;
typedef unsigned long long u64g;
__device__ __forceinline__ u64g gfma2(u64g a, u64g b, u64g c) {
    u64g r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ void gunpack(u64g v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

// list != nullptr: the rows are list[0 .. *n_list) (the uncertified rows of the wide tensor search) instead of 0 .. n; the
// grid is then (row tiles, centroid tiles): a short list would otherwise leave most of the GPU idle behind a few blocks that
// each sweep all centroids.  Block (., y) handles centroid tile y only and folds its result into keys[list position] =
// (distance bits << 32 | column) with atomicMin (non-negative floats order like their patterns; equal distances: the lower
// column), k_gemm_keys turns the keys into labels.  The row tiles are walked with a grid stride.
__global__ void __launch_bounds__(256, 2) k_assign_gemm(const float *__restrict__ x, int64_t n, int d, const float *__restrict__ c,
                                                        const float *__restrict__ cn, int k, int32_t *__restrict__ labels32,
                                                        int64_t *__restrict__ labels64, float *__restrict__ dist,
                                                        const uint32_t *__restrict__ list,
                                                        const unsigned int *__restrict__ n_list,
                                                        unsigned long long *__restrict__ counters,
                                                        unsigned long long *__restrict__ keys) {
    __shared__ __align__(16) float2 sa[2][GK][GM];   // (a, a): rows duplicated            16 KB
    __shared__ __align__(16) float sb[2][GK][GN];    //                                      8 KB
    __shared__ float s_xn[GM];
    __shared__ int64_t s_row[GM];                    // global row of the block's row slot (-1: none)
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    if (list) n = (int64_t)*n_list;
    if (counters && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) atomicAdd(&counters[1], (unsigned long long)n);   // rows scanned exactly
    const int jbeg = keys ? (int)blockIdx.y * GN : 0, jend = keys ? min(k, jbeg + GN) : k;
    for (int64_t row0 = (int64_t)blockIdx.x * GM; row0 < n; row0 += (int64_t)gridDim.x * GM) {   // block-uniform
    __syncthreads();   // the previous row tile is done with s_row / s_xn
    if (tid < GM) s_row[tid] = row0 + tid < n ? (list ? (int64_t)list[row0 + tid] : row0 + tid) : -1;
    __syncthreads();
    // |x|^2 of the block's rows: one sequential FMA chain per row (thread t < 128 walks row t)
    if (tid < GM) {
        float q = 0.f;
        const int64_t r = s_row[tid];
        if (r >= 0) {
            const float *xr = x + r * d;
            for (int t = 0; t < d; t++) q = fmaf(xr[t], xr[t], q);
        }
        s_xn[tid] = q;
    }
    // loader role: float4 `tid` of a 128 x 8 chunk = row / column (tid >> 1), k offset 4 (tid & 1)
    const int lr = tid >> 1, lk = (tid & 1) * 4;
    const bool vec = (d & 3) == 0;
    auto fetch = [&](const float *base, int64_t r, int64_t rmax, int k0) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r >= 0 && r < rmax) {
            const float *p = base + r * d + k0 + lk;
            if (vec && k0 + lk + 3 < d) v = *reinterpret_cast<const float4 *>(p);
            else {
                if (k0 + lk + 0 < d) v.x = p[0];
                if (k0 + lk + 1 < d) v.y = p[1];
                if (k0 + lk + 2 < d) v.z = p[2];
                if (k0 + lk + 3 < d) v.w = p[3];
            }
        }
        return v;
    };
    auto stash = [&](int buf, const float4 av, const float4 bv) {
        sa[buf][lk + 0][lr] = make_float2(av.x, av.x), sa[buf][lk + 1][lr] = make_float2(av.y, av.y);
        sa[buf][lk + 2][lr] = make_float2(av.z, av.z), sa[buf][lk + 3][lr] = make_float2(av.w, av.w);
        sb[buf][lk + 0][lr] = bv.x, sb[buf][lk + 1][lr] = bv.y, sb[buf][lk + 2][lr] = bv.z, sb[buf][lk + 3][lr] = bv.w;
    };
    float best[8];
    int bidx[8];
#pragma unroll
    for (int r = 0; r < 8; r++) best[r] = INFINITY, bidx[r] = 0x7FFFFFFF;
    const int nchunk = (d + GK - 1) / GK;
    for (int j0 = jbeg; j0 < jend; j0 += GN) {
        u64g acc[8][4];   // acc[r][q]: row r of the thread, columns 2q, 2q+1 of its eight
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[r][q] = 0ull;
        const int64_t xrow = s_row[lr];   // this loader thread's row of x (-1: zero fill)
        float4 av = fetch(x, xrow, INT64_MAX, 0), bv = fetch(c, (int64_t)j0 + lr, k, 0);
        __syncthreads();   // the previous tile's last chunk (and s_xn) are done with
        stash(0, av, bv);
        __syncthreads();
        for (int ch = 0; ch < nchunk; ch++) {
            const int buf = ch & 1;
            if (ch + 1 < nchunk) av = fetch(x, xrow, INT64_MAX, (ch + 1) * GK), bv = fetch(c, (int64_t)j0 + lr, k, (ch + 1) * GK);
#pragma unroll
            for (int kk = 0; kk < GK; kk++) {
                // rows 4 ty .. 4 ty + 3 and 64 + 4 ty ..: four 128-bit loads of (a, a) pairs; columns: two loads
                const ulonglong2 a01 = *reinterpret_cast<const ulonglong2 *>(&sa[buf][kk][ty * 4]);
                const ulonglong2 a23 = *reinterpret_cast<const ulonglong2 *>(&sa[buf][kk][ty * 4 + 2]);
                const ulonglong2 a45 = *reinterpret_cast<const ulonglong2 *>(&sa[buf][kk][64 + ty * 4]);
                const ulonglong2 a67 = *reinterpret_cast<const ulonglong2 *>(&sa[buf][kk][64 + ty * 4 + 2]);
                const ulonglong2 b03 = *reinterpret_cast<const ulonglong2 *>(&sb[buf][kk][tx * 4]);
                const ulonglong2 b47 = *reinterpret_cast<const ulonglong2 *>(&sb[buf][kk][64 + tx * 4]);
                const u64g ar[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
                const u64g bq[4] = {b03.x, b03.y, b47.x, b47.y};
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[r][q] = gfma2(ar[r], bq[q], acc[r][q]);
            }
            if (ch + 1 < nchunk) {
                stash(buf ^ 1, av, bv);   // the other buffer was last read in iteration ch - 1, before the barrier below
                __syncthreads();
            }
        }
        // tile epilogue: distances, the thread's best of its 8 columns per row (ascending), then across the 16 threads of the row
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const float xn = s_xn[(r < 4 ? 0 : 64) + ty * 4 + (r & 3)];
            float tb = INFINITY;
            int ti = 0x7FFFFFFF;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                float p0, p1;
                gunpack(acc[r][q], p0, p1);
                const int j = j0 + (q < 2 ? 0 : 64) + tx * 4 + (q & 1) * 2;
                if (j < k) {
                    const float dj = l2_expanded(xn, cn[j], p0);
                    if (dj < tb) tb = dj, ti = j;
                }
                if (j + 1 < k) {
                    const float dj = l2_expanded(xn, cn[j + 1], p1);
                    if (dj < tb) tb = dj, ti = j + 1;
                }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, tb, o);
                const int oi = __shfl_xor_sync(0xffffffffu, ti, o);
                if (od < tb || (od == tb && oi < ti)) tb = od, ti = oi;
            }
            if (tb < best[r]) best[r] = tb, bidx[r] = ti;   // tiles ascend: strict '<' keeps the lowest index
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const int slot = (r < 4 ? 0 : 64) + ty * 4 + (r & 3);
            const int64_t row = s_row[slot];
            if (row >= 0) {
                if (keys) {
                    if (bidx[r] != 0x7FFFFFFF)
                        atomicMin(&keys[row0 + slot], ((unsigned long long)__float_as_uint(best[r]) << 32) | (unsigned int)bidx[r]);
                } else {
                    const int bj = bidx[r] == 0x7FFFFFFF ? 0 : bidx[r];
                    if (labels32) labels32[row] = bj;
                    if (labels64) labels64[row] = bj;
                    if (dist) dist[row] = best[r];
                }
            }
        }
    }
    }   // row tiles
}

// keys[0 .. *n_list) = all ones (nothing seen)
__global__ void k_gemm_keys_init(unsigned long long *__restrict__ keys, const unsigned int *__restrict__ n_list) {
    const unsigned int n = *n_list;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) keys[i] = ~0ull;
}
// keys -> labels (and distances) of the listed rows; a row whose every distance was NaN keeps label 0 like the exact kernels
__global__ void k_gemm_keys(const uint32_t *__restrict__ list, const unsigned int *__restrict__ n_list,
                            const unsigned long long *__restrict__ keys, int32_t *__restrict__ labels32,
                            int64_t *__restrict__ labels64, float *__restrict__ dist) {
    const unsigned int n = *n_list;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long key = keys[i];
        const int64_t row = (int64_t)list[i];
        const int bj = key == ~0ull ? 0 : (int)(unsigned int)key;
        if (labels32) labels32[row] = bj;
        if (labels64) labels64[row] = bj;
        if (dist) dist[row] = key == ~0ull ? INFINITY : __uint_as_float((unsigned int)(key >> 32));
    }
}

int launch_assign_gemm(const at_index *ix, const float *x, int64_t n, int32_t *l32, int64_t *l64, float *dist, cudaStream_t st) {
    k_assign_gemm<<<(unsigned)ceil_div(n, GM), 256, 0, st>>>(x, n, ix->d, ix->c, ix->cn, ix->k, l32, l64, dist, nullptr, nullptr,
                                                             nullptr, nullptr);
    AT_LAUNCH_OK();
    return AT_OK;
}

// the listed rows only (the length lives on the device): (row tiles, centroid tiles) grid + keys, see k_assign_gemm
int launch_assign_gemm_list(const at_index *ix, const float *x, const uint32_t *list, const unsigned int *n_list, int64_t n_max,
                            unsigned long long *keys, int32_t *l32, int64_t *l64, float *dist, cudaStream_t st) {
    const int sms = sm_count() > 0 ? sm_count() : 1;
    int64_t gx = ceil_div(n_max, GM);
    if (gx > 2 * sms) gx = 2 * sms;   // grid stride over longer lists
    k_gemm_keys_init<<<sms, 256, 0, st>>>(keys, n_list);
    AT_LAUNCH_OK();
    k_assign_gemm<<<dim3((unsigned)gx, (unsigned)ceil_div(ix->k, GN)), 256, 0, st>>>(x, n_max, ix->d, ix->c, ix->cn, ix->k, nullptr,
                                                                                    nullptr, nullptr, list, n_list,
                                                                                    ix->tc_counters, keys);
    AT_LAUNCH_OK();
    k_gemm_keys<<<sms, 256, 0, st>>>(list, n_list, keys, l32, l64, dist);
    AT_LAUNCH_OK();
    return AT_OK;
}

}  // namespace at

using namespace at;

extern "C" int at_conv_expand(const float *x, int64_t n, int n_mels, const float *weight, const float *bias, int num_kernels,
                              int kernel_size, float *out, void *stream) {
    AT_REQUIRE(x && weight && bias && out && n >= 0 && n_mels > 0 && num_kernels > 0 && kernel_size > 0 && (kernel_size & 1),
               "at_conv_expand: bad arguments (odd kernel sizes only: padding = kernel_size // 2 keeps the width)");
    AT_REQUIRE(num_kernels <= 256 && kernel_size <= CONV_MAX_KS && n_mels <= 4096,
               "at_conv_expand: at most 256 kernels of at most %d taps over at most 4096 mel bins", CONV_MAX_KS);
    if (n == 0) return AT_OK;
    const int threads = num_kernels * (256 / num_kernels);
    int64_t want = ceil_div(n, CONV_ROWS);
    int blocks = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
    if (blocks < 1) blocks = 1;
    const size_t smem = sizeof(float) * (size_t)CONV_ROWS * (size_t)(n_mels + 2 * (kernel_size / 2));
    if (smem > 48 * 1024) {
        set_error("at_conv_expand: rows of %d mel bins do not fit the staging buffer", n_mels);
        return AT_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (kernel_size) {
        case 1: k_conv_expand<1><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        case 3: k_conv_expand<3><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        case 5: k_conv_expand<5><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        case 7: k_conv_expand<7><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        case 9: k_conv_expand<9><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        case 11: k_conv_expand<11><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        case 13: k_conv_expand<13><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
        default: k_conv_expand<15><<<blocks, threads, smem, st>>>(x, n, n_mels, weight, bias, num_kernels, out); break;
    }
    AT_LAUNCH_OK();
    return AT_OK;
}
