// libat_b200: the use_convolution branch of stages 2-3.
//
// With config.use_convolution the reference expands every 64-value mel frame through nn.Conv1d(1 -> num_kernels,
// kernel_size, padding = kernel_size // 2) along the mel axis and clusters / tokenizes the (n_mels * num_kernels)-value rows
// (processors/cluster_creator.py:28-34,68-81; processors/spec_tokenizer.py:92-104,115-121).  Two kernels:
//   k_conv_expand   rows (n, n_mels) -> (n, n_mels * num_kernels), out[i, m * Kc + c] = b[c] + sum_t w[c][t] x[i, m + t - pad]
//                   (the layout of conv_output.transpose(1, 2).reshape(-1, Kc * n_mels))
//   k_assign_gemm   exact fp32 nearest-centroid search for rows wider than the register-resident kernel covers (d > 128):
//                   a 64 x 64 tile of inner products per block (4 x 4 per thread, K chunks of 16 through shared memory),
//                   |x|^2 + |c|^2 - 2 <x, c> clamped at 0, strict '<' in ascending column order (lowest index wins ties).
#include "at_index.cuh"

namespace at {

// Block = kc * (256 / kc) threads: thread t owns channel c = t % kc (its ks weights and bias live in registers for the
// whole launch) and walks the mel bins m = t / kc, + threads / kc, ... of one row after the other, so consecutive threads
// write consecutive floats of the output row (coalesced) and nothing is divided inside the loops.  CONV_ROWS rows are
// staged in shared memory per block step (zero padded at both ends: no bounds tests on the taps).
constexpr int CONV_ROWS = 8, CONV_MAX_KS = 15;
__global__ void __launch_bounds__(256) k_conv_expand(const float *__restrict__ x, int64_t n, int n_mels,
                                                     const float *__restrict__ w, const float *__restrict__ b, int kc, int ks,
                                                     float *__restrict__ out) {
    extern __shared__ float s_x[];   // CONV_ROWS rows of (pad + n_mels + pad) floats
    const int pad = ks / 2, stride = n_mels + 2 * pad;
    const int c = (int)threadIdx.x % kc, m0 = (int)threadIdx.x / kc, mstep = (int)blockDim.x / kc;
    float wr[CONV_MAX_KS];
#pragma unroll
    for (int t = 0; t < CONV_MAX_KS; t++) wr[t] = t < ks ? w[c * ks + t] : 0.f;
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
                for (int t = 0; t < CONV_MAX_KS; t++)
                    if (t < ks) acc = fmaf(wr[t], xr[m + t], acc);   // ascending taps, like the first version (same rounding)
                orow[(int64_t)m * kc] = acc;
            }
        }
    }
}

constexpr int GM = 64, GN = 64, GK = 16;
__global__ void __launch_bounds__(256) k_assign_gemm(const float *__restrict__ x, int64_t n, int d, const float *__restrict__ c,
                                                     const float *__restrict__ cn, int k, int32_t *__restrict__ labels32,
                                                     int64_t *__restrict__ labels64, float *__restrict__ dist) {
    __shared__ float sa[GK][GM + 4];   // sa[kk][row]
    __shared__ float sb[GK][GN + 4];   // sb[kk][col]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t row0 = (int64_t)blockIdx.x * GM;
    float best[4];
    int bidx[4];
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 4; r++) best[r] = INFINITY, bidx[r] = 0;
    // loader roles: 256 threads fetch a 64 x 16 tile of each operand, one float4 per thread (d is a multiple of 4 or
    // handled element-wise)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const bool vec = (d & 3) == 0;
    for (int j0 = 0; j0 < k; j0 += GN) {
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[r][q] = 0.f;
        float xs[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k0 = 0; k0 < d; k0 += GK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
            const int64_t ar = row0 + lr;
            const int bc = j0 + lr;
            if (ar < n) {
                if (vec && k0 + lk + 3 < d) {
                    const float4 v = *reinterpret_cast<const float4 *>(x + ar * d + k0 + lk);
                    av[0] = v.x, av[1] = v.y, av[2] = v.z, av[3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (k0 + lk + e < d) av[e] = x[ar * d + k0 + lk + e];
                }
            }
            if (bc < k) {
                if (vec && k0 + lk + 3 < d) {
                    const float4 v = *reinterpret_cast<const float4 *>(c + (int64_t)bc * d + k0 + lk);
                    bv[0] = v.x, bv[1] = v.y, bv[2] = v.z, bv[3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (k0 + lk + e < d) bv[e] = c[(int64_t)bc * d + k0 + lk + e];
                }
            }
            __syncthreads();   // the previous chunk has been consumed
#pragma unroll
            for (int e = 0; e < 4; e++) sa[lk + e][lr] = av[e], sb[lk + e][lr] = bv[e];
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < GK; kk++) {
                const float4 a4 = *reinterpret_cast<const float4 *>(&sa[kk][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4 *>(&sb[kk][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    if (j0 == 0) xs[r] = fmaf(a[r], a[r], xs[r]);
#pragma unroll
                    for (int q = 0; q < 4; q++) acc[r][q] = fmaf(a[r], bb[q], acc[r][q]);
                }
            }
        }
        if (j0 == 0) {
#pragma unroll
            for (int r = 0; r < 4; r++) xn[r] = xs[r];
        }
        // tile epilogue: distance, per-row best of this thread's 4 columns (ascending), then across the 16 threads of the row
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float tb = INFINITY;
            int ti = 0x7FFFFFFF;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = j0 + tx * 4 + q;
                if (j < k) {
                    const float dj = l2_expanded(xn[r], cn[j], acc[r][q]);
                    if (dj < tb) tb = dj, ti = j;
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
        for (int r = 0; r < 4; r++) {
            const int64_t row = row0 + ty * 4 + r;
            if (row < n) {
                const int bj = bidx[r] == 0x7FFFFFFF ? 0 : bidx[r];
                if (labels32) labels32[row] = bj;
                if (labels64) labels64[row] = bj;
                if (dist) dist[row] = best[r];
            }
        }
    }
}

int launch_assign_gemm(const at_index *ix, const float *x, int64_t n, int32_t *l32, int64_t *l64, float *dist, cudaStream_t st) {
    k_assign_gemm<<<(unsigned)ceil_div(n, GM), 256, 0, st>>>(x, n, ix->d, ix->c, ix->cn, ix->k, l32, l64, dist);
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
    k_conv_expand<<<blocks, threads, smem, (cudaStream_t)stream>>>(x, n, n_mels, weight, bias, num_kernels, kernel_size, out);
    AT_LAUNCH_OK();
    return AT_OK;
}
