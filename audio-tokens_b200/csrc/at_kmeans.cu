// libat_b200: flat L2 index (exact fp32 SIMT search), k-means accumulate / finalize.
//
// Reference semantics restated here (upstream faiss v1.8.0, reached from processors/cluster_creator.py:42-58
// and processors/spec_tokenizer.py:77):
//   search      exhaustive_L2sqr_blas: |x|^2 + |c|^2 - 2<x,c>, clamp at 0, strict '<' (lowest index wins)
//   accumulate  compute_centroids' per-cluster sums and counts (here: exact fixed-point, order independent)
//   finalize    c = sum * (1/count); split_clusters with std::mt19937(1234), EPS = 1/1024
#include "at_index.cuh"
#include <stdlib.h>

#include <math.h>
#include <new>

namespace at {

// ---------------------------------------------------------------------------------------------
// normalize_vectors: half-warp per row, canonical sum of squares.
// ---------------------------------------------------------------------------------------------
__global__ void k_row_l2norm(const float *__restrict__ x, int64_t n, int d, float *__restrict__ out) {
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    int g = threadIdx.x & 15;
    bool live = row < n;
    const float *xr = x + (live ? row : 0) * d;
    float q = 0.f;
    if (live) {
        for (int base = 4 * g; base < d; base += 64) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (base + e < d) {
                    float v = xr[base + e];
                    q = fmaf(v, v, q);
                }
            }
        }
    }
    float s = half16_sum(q);
    if (!live) return;
    float den = l2_denominator(s);
    for (int base = 4 * g; base < d; base += 64) {
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (base + e < d) out[row * d + base + e] = __fdiv_rn(xr[base + e], den);
    }
}

// canonical |c|^2, one thread per centroid (k is small).  maxes (may be null): {max |c_ij|, max |c_j|^2} as float bit
// patterns raised with atomicMax (non-negative floats order like their patterns; NaN patterns sort above every number,
// so a non-finite centroid is seen by the consumer), consumed and reset by k_tc_scale.
// maxes (tensor path, d == 64): {max |c_ij - m_i|, max |c_j - m|^2, max |c_j|^2} with m = shift, the centring vector of the
// operand image (at_index.cuh).
__global__ void k_centroid_norms(const float *__restrict__ c, int k, int d, float *__restrict__ cn,
                                 unsigned int *__restrict__ maxes, const float *__restrict__ shift) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f, n2 = 0.f, ns = 0.f;
    if (j < k) {
        float q[16];
#pragma unroll
        for (int l = 0; l < 16; l++) q[l] = 0.f;
        const float *cj = c + (int64_t)j * d;
        for (int base = 0; base < d; base += 64) {
#pragma unroll
            for (int t = 0; t < 64; t++) {
                if (base + t < d) {
                    float v = cj[base + t];
                    q[(t >> 2) & 15] = fmaf(v, v, q[(t >> 2) & 15]);
                    if (maxes) {
                        const float vs = v - shift[base + t];
                        ns = fmaf(vs, vs, ns);
                        m = fmaxf(m, fabsf(vs));
                        if (!(vs == vs)) m = vs;
                    }
                }
            }
        }
        n2 = tree16(q);
        cn[j] = n2;
    }
    if (maxes) {
        unsigned int um = __float_as_uint(fabsf(m)), un = __float_as_uint(fabsf(ns)), uo = __float_as_uint(fabsf(n2));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            um = max(um, __shfl_xor_sync(0xffffffffu, um, o));
            un = max(un, __shfl_xor_sync(0xffffffffu, un, o));
            uo = max(uo, __shfl_xor_sync(0xffffffffu, uo, o));
        }
        if ((threadIdx.x & 31) == 0) atomicMax(maxes, um), atomicMax(maxes + 1, un), atomicMax(maxes + 2, uo);
    }
}

// ---------------------------------------------------------------------------------------------
// Exact fp32 search, one thread per row, row held in registers (zero padded to DP).
// ---------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(128) k_assign_simt(const float *__restrict__ x, int64_t n, int d,
                                                     const float *__restrict__ c,
                                                     const float *__restrict__ cn, int k, int l2norm,
                                                     int32_t *__restrict__ labels32,
                                                     int64_t *__restrict__ labels64,
                                                     float *__restrict__ dist) {
    constexpr int KT = 32;
    __shared__ __align__(16) float ctile[KT][DP];
    __shared__ float cns[KT];
    const int tid = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * 128 + tid;
    const bool live = row < n;

    float xr[DP];
#pragma unroll
    for (int t = 0; t < DP; t++) xr[t] = 0.f;
    if (live) {
        const float *xp = x + row * d;
        if ((d & 3) == 0) {
#pragma unroll
            for (int t = 0; t < DP; t += 4) {
                if (t < d) {
                    float4 v = *reinterpret_cast<const float4 *>(xp + t);
                    xr[t] = v.x, xr[t + 1] = v.y, xr[t + 2] = v.z, xr[t + 3] = v.w;
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < DP; t++)
                if (t < d) xr[t] = xp[t];
        }
    }
    float q[16];
    if (l2norm) {
#pragma unroll
        for (int l = 0; l < 16; l++) q[l] = 0.f;
#pragma unroll
        for (int t = 0; t < DP; t++) q[(t >> 2) & 15] = fmaf(xr[t], xr[t], q[(t >> 2) & 15]);
        float den = l2_denominator(tree16(q));
#pragma unroll
        for (int t = 0; t < DP; t++) xr[t] = __fdiv_rn(xr[t], den);
    }
#pragma unroll
    for (int l = 0; l < 16; l++) q[l] = 0.f;
#pragma unroll
    for (int t = 0; t < DP; t++) q[(t >> 2) & 15] = fmaf(xr[t], xr[t], q[(t >> 2) & 15]);
    const float xn = tree16(q);

    float best = INFINITY;
    int bj = 0;
    for (int j0 = 0; j0 < k; j0 += KT) {
        const int kt = min(KT, k - j0);
        for (int i = tid; i < KT * DP; i += 128) {
            int jj = i / DP, t = i % DP;
            ctile[jj][t] = (jj < kt && t < d) ? c[(int64_t)(j0 + jj) * d + t] : 0.f;
        }
        if (tid < KT) cns[tid] = tid < kt ? cn[j0 + tid] : INFINITY;
        __syncthreads();
        for (int jj = 0; jj < kt; jj++) {
#pragma unroll
            for (int l = 0; l < 16; l++) q[l] = 0.f;
#pragma unroll
            for (int t = 0; t < DP; t += 4) {
                float4 cv = *reinterpret_cast<const float4 *>(&ctile[jj][t]);
                const int l = (t >> 2) & 15;
                q[l] = fmaf(xr[t], cv.x, q[l]);
                q[l] = fmaf(xr[t + 1], cv.y, q[l]);
                q[l] = fmaf(xr[t + 2], cv.z, q[l]);
                q[l] = fmaf(xr[t + 3], cv.w, q[l]);
            }
            float dis = l2_expanded(xn, cns[jj], tree16(q));
            if (dis < best) {
                best = dis;
                bj = j0 + jj;
            }
        }
        __syncthreads();
    }
    if (live) {
        if (labels32) labels32[row] = bj;
        if (labels64) labels64[row] = bj;
        if (dist) dist[row] = best;
    }
}

// Rows whose label changed since the previous accumulate: two delta items each (leave the old cluster, join the new
// one), prev updated.  List slots are reserved once per block step (a per-warp atomic on the single counter serialises
// in L2).
__global__ void __launch_bounds__(256) k_diff(const int32_t *__restrict__ labels, int32_t *__restrict__ prev, int64_t n,
                                              int32_t *__restrict__ d_row, int32_t *__restrict__ d_lab,
                                              unsigned int *__restrict__ d_count) {
    __shared__ unsigned int s_cnt[8];
    __shared__ unsigned int s_base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n; base += stride) {   // block-uniform trip count
        const int64_t i = base + threadIdx.x;
        bool changed = false;
        int l = 0, p = 0;
        if (i < n) {
            l = labels[i], p = prev[i];
            changed = l != p;
        }
        const unsigned m = __ballot_sync(0xffffffffu, changed);
        if (lane == 0) s_cnt[w] = (unsigned int)__popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int tot = 0;
            for (int q = 0; q < 8; q++) {
                const unsigned int c = s_cnt[q];
                s_cnt[q] = tot;
                tot += c;
            }
            s_base = tot ? atomicAdd(d_count, tot) : 0u;
        }
        __syncthreads();
        if (changed) {
            const unsigned int s = s_base + s_cnt[w] + __popc(m & ((1u << lane) - 1u));
            d_row[2 * s] = (int32_t)((uint32_t)i | 0x80000000u), d_lab[2 * s] = p;
            d_row[2 * s + 1] = (int32_t)i, d_lab[2 * s + 1] = l;
            prev[i] = l;
        }
        __syncthreads();
    }
}

// Delta form (items != nullptr): besides the histogram of the items, the persistent per-cluster counts move by -1 /
// +1 per item (bit 31 of the item = the row leaves the cluster).
__global__ void k_counts(const int32_t *__restrict__ labels, int64_t n, unsigned long long *__restrict__ counts,
                         const unsigned int *__restrict__ n_dev = nullptr, int n_mult = 1,
                         const int32_t *__restrict__ items = nullptr, unsigned long long *__restrict__ lcounts = nullptr) {
    if (n_dev) n = (int64_t)n_mult * (int64_t)*n_dev;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // warp-uniform trip count so the full-mask match below is always convergent
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
        const int64_t i = base + lane;
        const int l = i < n ? labels[i] : -1 - lane;
        // warp-aggregate identical labels before touching L2
        unsigned peers = __match_any_sync(0xffffffffu, l);
        if (l >= 0 && (__ffs(peers) - 1) == lane) atomicAdd(&counts[l], (unsigned long long)__popc(peers));
        if (items) {
            const int leaves = (i < n) ? (int)((uint32_t)items[i] >> 31) : 0;
            const unsigned peers2 = __match_any_sync(0xffffffffu, 2 * l + leaves);
            if (l >= 0 && (__ffs(peers2) - 1) == lane) {
                const unsigned long long c = (unsigned long long)__popc(peers2);
                atomicAdd(&lcounts[l], leaves ? (0ULL - c) : c);
            }
        }
    }
}

// exclusive scan of the k local counts -> off[0..k], cursor[j] = off[j].  One block of 1024 threads.
__global__ void k_scan_counts(const unsigned long long *__restrict__ counts, int k, int64_t *__restrict__ off,
                              unsigned long long *__restrict__ cursor) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int per = (k + 1023) / 1024;
    const int j0 = tid * per, j1 = min(k, j0 + per);
    long long s = 0;
    for (int j = j0; j < j1; j++) s += (long long)counts[j];
    part[tid] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int o = 1; o < 1024; o <<= 1) {
        long long v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = part[tid] - s;
    for (int j = j0; j < j1; j++) {
        off[j] = run;
        cursor[j] = (unsigned long long)run;
        run += (long long)counts[j];
    }
    if (tid == 1023) off[k] = part[1023];
}

// order[cursor[label]++] = row  (order within a cluster is arbitrary: the sums below are exact integers)
__global__ void k_place(const int32_t *__restrict__ labels, int64_t n, unsigned long long *__restrict__ cursor,
                        int32_t *__restrict__ order, const unsigned int *__restrict__ n_dev = nullptr, int n_mult = 1) {
    if (n_dev) n = (int64_t)n_mult * (int64_t)*n_dev;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < n; base += stride) {
        const int64_t i = base + lane;
        const int l = i < n ? labels[i] : -1 - lane;
        unsigned peers = __match_any_sync(0xffffffffu, l);
        int leader = __ffs(peers) - 1;
        unsigned long long pos = 0;
        if (l >= 0 && lane == leader) pos = atomicAdd(&cursor[l], (unsigned long long)__popc(peers));
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (l >= 0) order[pos + __popc(peers & ((1u << lane) - 1u))] = (int32_t)i;
    }
}

// Each warp owns an equal slice of the cluster-grouped order[] array and accumulates x rows as int64
// fixed point (x * 2^e_sum), flushing to sums[c] with 64-bit atomics whenever the cluster changes.
// Delta form (items != nullptr): order[] indexes items[], an item is a row with bit 31 set when the row LEAVES the
// cluster (its value is subtracted); the item count is 2 * *n_dev.
template <int DPL, int U = 8>  // dims per lane = ceil(d / 32); rows fetched ahead per warp (fewer for wide rows: registers)
__global__ void __launch_bounds__(256) k_gather_sum(const float *__restrict__ x, int d,
                                                    const int32_t *__restrict__ order,
                                                    const int64_t *__restrict__ off, int k, int64_t n,
                                                    float scale, unsigned long long *__restrict__ sums,
                                                    const int32_t *__restrict__ items = nullptr,
                                                    const unsigned int *__restrict__ n_dev = nullptr,
                                                    unsigned long long *__restrict__ sumsq = nullptr, float sq_scale = 0.f) {
    if (n_dev) n = 2 * (int64_t)*n_dev;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t per = (n + nwarps - 1) / nwarps;
    const int64_t p0 = warp * per;
    const int64_t p1 = min(n, p0 + per);
    if (p0 >= p1) return;
    // cluster of p0: largest c with off[c] <= p0
    int lo = 0, hi = k;  // invariant off[lo] <= p0 < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= p0) lo = mid; else hi = mid;
    }
    int c = lo;
    int64_t cend = off[c + 1];
    long long acc[DPL];
#pragma unroll
    for (int u = 0; u < DPL; u++) acc[u] = 0;
    long long sq = 0;   // sum of x^2 over this warp's rows (fixed point), the constant term of the objective

    for (int64_t p = p0; p < p1; p += U) {
        int32_t rows[U];
        float v[U][DPL];
        float sgn[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            rows[u] = (p + u < p1) ? order[p + u] : -1;
            sgn[u] = scale;
            if (items && rows[u] >= 0) {
                const uint32_t it = (uint32_t)items[rows[u]];
                rows[u] = (int32_t)(it & 0x7FFFFFFFu);
                if (it >> 31) sgn[u] = -scale;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
#pragma unroll
            for (int w = 0; w < DPL; w++) {
                int t = lane + 32 * w;
                v[u][w] = (rows[u] >= 0 && t < d) ? x[(int64_t)rows[u] * d + t] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (p + u >= p1) break;
            while (p + u >= cend) {  // warp-uniform
#pragma unroll
                for (int w = 0; w < DPL; w++) {
                    int t = lane + 32 * w;
                    if (t < d && acc[w] != 0) atomicAdd(&sums[(int64_t)c * d + t], (unsigned long long)acc[w]);
                    acc[w] = 0;
                }
                c++;
                cend = off[c + 1];
            }
#pragma unroll
            for (int w = 0; w < DPL; w++) acc[w] += __float2ll_rn(v[u][w] * sgn[u]);
            if (sumsq) {
#pragma unroll
                for (int w = 0; w < DPL; w++) sq += __float2ll_rn(v[u][w] * v[u][w] * sq_scale);
            }
        }
    }
    if (sumsq) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) atomicAdd(sumsq, (unsigned long long)sq);
    }
#pragma unroll
    for (int w = 0; w < DPL; w++) {
        int t = lane + 32 * w;
        if (t < d && acc[w] != 0) atomicAdd(&sums[(int64_t)c * d + t], (unsigned long long)acc[w]);
    }
}

// ---------------------------------------------------------------------------------------------
// finalize: division, statistics, split_clusters.  One block.
// ---------------------------------------------------------------------------------------------
struct Mt19937 {
    uint32_t mt[624];
    int idx;
    __device__ void seed(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    __device__ uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; i++) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                uint32_t v = mt[(i + 397) % 624] ^ (y >> 1);
                if (y & 1u) v ^= 0x9908b0dfu;
                mt[i] = v;
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    // faiss RandomGenerator::rand_float: mt() / float(mt.max())
    __device__ float rand_float() { return __fdiv_rn((float)next(), 4294967296.0f); }
};

// compute_centroids' division, one warp per cluster: c[j] = sum * (1 / count) (empty clusters stay zero), plus the
// cluster's share of the objective against the centroids the rows were assigned with,
//     sum_i |x_i - c_old|^2 = sum_i |x_i|^2 + count |c_old|^2 - 2 <sum_i x_i, c_old>      (exact sums, double arithmetic),
// of the imbalance factor and of the empty-cluster count, as one partial triple per block (summed in a fixed order by
// k_finalize_split, so the statistics are reproducible bit for bit).
constexpr int FIN_WARPS = 8;
__global__ void __launch_bounds__(FIN_WARPS * 32) k_finalize_div(const long long *__restrict__ accum, int k, int d,
                                                               double inv_sum_scale, const float *__restrict__ c_old,
                                                               float *__restrict__ centroids, float *__restrict__ hassign,
                                                               double *__restrict__ part) {
    __shared__ double s_part[FIN_WARPS][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * FIN_WARPS + w;
    const int64_t kd = (int64_t)k * d;
    double obj = 0.0, imb = 0.0, ne = 0.0;
    if (c < k) {
        const long long cnt = accum[kd + c];
        const float norm = cnt != 0 ? __fdiv_rn(1.0f, (float)cnt) : 0.f;
        double dot = 0.0, cc = 0.0;
        for (int t = lane; t < d; t += 32) {
            const double sd = __ll2double_rn(accum[(int64_t)c * d + t]) * inv_sum_scale;
            const double co = (double)c_old[(int64_t)c * d + t];
            dot = fma(sd, co, dot);
            cc = fma(co, co, cc);
            centroids[(int64_t)c * d + t] = cnt != 0 ? __fmul_rn(__double2float_rn(sd), norm) : 0.f;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            cc += __shfl_xor_sync(0xffffffffu, cc, o);
        }
        obj = (double)cnt * cc - 2.0 * dot;
        imb = (double)cnt * (double)cnt;
        ne = cnt == 0 ? 1.0 : 0.0;
        if (lane == 0) hassign[c] = (float)cnt;
    }
    if (lane == 0) s_part[w][0] = obj, s_part[w][1] = imb, s_part[w][2] = ne;
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int q = 0; q < FIN_WARPS; q++) v += s_part[q][threadIdx.x];
        part[(int64_t)blockIdx.x * 3 + threadIdx.x] = v;
    }
}

// statistics + faiss::split_clusters.  One block; the split itself is sequential by construction.
__global__ void __launch_bounds__(256) k_finalize_split(const long long *__restrict__ accum, int k, int d, int64_t n_total,
                                                        double inv_obj_scale, const double *__restrict__ part, int nblk,
                                                        float *__restrict__ centroids, float *__restrict__ hassign,
                                                        float *__restrict__ stats) {
    __shared__ Mt19937 rng;
    __shared__ double s_red[3][256];
    const int tid = threadIdx.x;
    double v[3] = {0.0, 0.0, 0.0};
    for (int b = tid; b < nblk; b += 256) {
#pragma unroll
        for (int q = 0; q < 3; q++) v[q] += part[(int64_t)b * 3 + q];
    }
#pragma unroll
    for (int q = 0; q < 3; q++) s_red[q][tid] = v[q];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
#pragma unroll
            for (int q = 0; q < 3; q++) s_red[q][tid] += s_red[q][tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int64_t kd = (int64_t)k * d;
        const double obj = (double)accum[kd + k] * inv_obj_scale + s_red[0][0];   // accum[kd + k]: sum of |x_i|^2, fixed point
        const double tot = s_red[1][0];
        const int ne = (int)(s_red[2][0] + 0.5);
        int nsplit = 0;
        if (ne > 0) {
            // faiss::split_clusters (Clustering.cpp)
            rng.seed(1234u);
            for (int ci = 0; ci < k; ci++) {
                if (hassign[ci] == 0.f) {
                    int cj = 0;
                    for (;; cj = (cj + 1) % k) {
                        float p = (float)(((double)hassign[cj] - 1.0) / (double)(float)(n_total - k));
                        float r = rng.rand_float();
                        if (r < p) break;
                    }
                    for (int j = 0; j < d; j++) {
                        float c0 = centroids[(int64_t)cj * d + j];
                        float up = __fmul_rn(c0, 1.0009765625f), dn = __fmul_rn(c0, 0.9990234375f);
                        centroids[(int64_t)ci * d + j] = (j % 2 == 0) ? up : dn;
                        centroids[(int64_t)cj * d + j] = (j % 2 == 0) ? dn : up;
                    }
                    hassign[ci] = hassign[cj] / 2;
                    hassign[cj] -= hassign[ci];
                    nsplit++;
                }
            }
        }
        if (stats) {
            stats[0] = (float)fmax(obj, 0.0);
            stats[1] = (float)nsplit;
            stats[2] = (float)(tot * (double)k / ((double)n_total * (double)n_total));
            stats[3] = (float)ne;
        }
    }
}

static int launch_simt(const at_index *ix, const float *x, int64_t n, int l2norm, int32_t *l32, int64_t *l64,
                       float *dist, cudaStream_t st) {
    unsigned blocks = (unsigned)ceil_div(n, 128);
    if (ix->d <= 16)
        k_assign_simt<16><<<blocks, 128, 0, st>>>(x, n, ix->d, ix->c, ix->cn, ix->k, l2norm, l32, l64, dist);
    else if (ix->d <= 32)
        k_assign_simt<32><<<blocks, 128, 0, st>>>(x, n, ix->d, ix->c, ix->cn, ix->k, l2norm, l32, l64, dist);
    else if (ix->d <= 64)
        k_assign_simt<64><<<blocks, 128, 0, st>>>(x, n, ix->d, ix->c, ix->cn, ix->k, l2norm, l32, l64, dist);
    else if (ix->d <= 128)
        k_assign_simt<128><<<blocks, 128, 0, st>>>(x, n, ix->d, ix->c, ix->cn, ix->k, l2norm, l32, l64, dist);
    else {
        // wide rows (the use_convolution branch: d = n_mels * num_kernels): tiled exact fp32 kernel, rows pre-normalised
        if (l2norm) {
            set_error("search: l2norm_rows is fused for d <= 128 only; normalise wider rows with at_row_l2norm first (d=%d)", ix->d);
            return AT_ERR_UNSUPPORTED;
        }
        return launch_assign_gemm(ix, x, n, l32, l64, dist, st);
    }
    AT_LAUNCH_OK();
    return AT_OK;
}

// exact_dist: the tensor path reads certified rows' distances off the accumulator; 1 = re-evaluate them with the
// canonical fp32 formula (what the public search returns), 0 = keep the read-out (k-means objective).
static int index_search(at_index *ix, const float *x, int64_t n, int l2norm, int algo, int32_t *l32,
                        int64_t *l64, float *dist, int exact_dist, at_tc_rows *rows, cudaStream_t st) {
    if (n == 0) return AT_OK;
    bool tc = false, tcw = false;
    // wide rows (d = 64 NS): labels from the slice-accumulating tensor kernel; pre-normalised rows, no exact distances
    const bool wide_ok = assign_tc_wide_supported(ix) && !l2norm && !(dist && exact_dist);
    if (algo == AT_ALGO_TENSOR) {
        if (assign_tc_supported(ix)) tc = true;
        else if (wide_ok) tcw = true;
        else {
            set_error("search: tensor path needs d == 64 (or a multiple of 64 up to 1024 with pre-normalised rows and no "
                      "distances) and k >= 16 (d=%d, k=%d)", ix->d, ix->k);
            return AT_ERR_UNSUPPORTED;
        }
    } else if (algo == AT_ALGO_AUTO) {
        tc = assign_tc_supported(ix) && ix->k >= 64;
        // (below ~8 M scores the exact tile kernel is as fast: the image build and a 148-CTA launch are fixed costs)
        tcw = !tc && wide_ok && ix->k >= 64 && !dist && n * (int64_t)ix->k >= (8LL << 20);
    }
    ProfScope prof(PROF_SEARCH, st);
    if (tc) return assign_tc_search(ix, x, n, l2norm, l32, l64, dist, exact_dist, rows, st);
    if (tcw) return assign_tc_wide_search(ix, x, n, l32, l64, dist, rows && rows->d == ix->d ? rows : nullptr, st);
    return launch_simt(ix, x, n, l2norm, l32, l64, dist, st);
}

}  // namespace at

using namespace at;

extern "C" {

int at_row_l2norm(const float *x, int64_t n, int d, float *out, void *stream) {
    AT_REQUIRE(x && out && n >= 0 && d > 0, "at_row_l2norm: bad arguments");
    if (n == 0) return AT_OK;
    k_row_l2norm<<<(unsigned)ceil_div(n * 16, 256), 256, 0, (cudaStream_t)stream>>>(x, n, d, out);
    AT_LAUNCH_OK();
    return AT_OK;
}

// ----------------------------------------------------------------------------------- index
int at_index_create(int d, at_index **index) {
    AT_REQUIRE(index && d > 0, "at_index_create: bad arguments");
    AT_REQUIRE(d <= 1024, "at_index_create: d=%d > 1024 is not covered by this build", d);
    int dev;
    AT_CUDA_OK(cudaGetDevice(&dev));
    at_index *ix = new (std::nothrow) at_index();
    if (!ix) return AT_ERR_NOMEM;
    ix->d = d;
    *index = ix;
    return AT_OK;
}

int at_index_destroy(at_index *ix) {
    if (!ix) return AT_OK;
    cudaFree(ix->c);
    cudaFree(ix->cn);
    cudaFree(ix->op);
    cudaFree(ix->tc_scale);
    cudaFree(ix->tc_max);
    cudaFree(ix->shift);
    cudaFree(ix->tc_counters);
    cudaFree(ix->part_lab);
    tc_rows_free(&ix->rows);
    if (ix->side) cudaStreamDestroy(ix->side);
    if (ix->ev_fork) cudaEventDestroy(ix->ev_fork);
    if (ix->ev_join) cudaEventDestroy(ix->ev_join);
    delete ix;
    return AT_OK;
}

// canonical norms + (d == 64) the tensor operands of the centroids currently in ix->c
static int index_refresh(at_index *ix, cudaStream_t st) {
    const bool tc = assign_tc_supported(ix) || assign_tc_wide_supported(ix);
    if (tc && !ix->ext_shift) {   // the index's own centring vector follows its centroids
        int rc = tc_mean(ix->c, ix->k, ix->d, ix->shift, st);
        if (rc != AT_OK) return rc;
    }
    k_centroid_norms<<<(ix->k + 127) / 128, 128, 0, st>>>(ix->c, ix->k, ix->d, ix->cn, tc ? ix->tc_max : nullptr,
                                                          ix->ext_shift ? ix->ext_shift : ix->shift);
    AT_LAUNCH_OK();
    if (tc) return assign_tc_prepare(ix, st);
    return AT_OK;
}

int at_index_set_centroids(at_index *ix, const float *centroids, int k, void *stream) {
    AT_REQUIRE(ix && centroids && k > 0, "at_index_set_centroids: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (k > ix->kcap) {
        AT_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(ix->c), cudaFree(ix->cn), cudaFree(ix->op);
        ix->c = ix->cn = nullptr;
        ix->op = nullptr;
        ix->kcap = 0;
        int ktiles = (k + 127) / 128;
        AT_CUDA_OK(cudaMalloc(&ix->c, sizeof(float) * (size_t)k * ix->d));
        AT_CUDA_OK(cudaMalloc(&ix->cn, sizeof(float) * (size_t)k));
        if (ix->d % 64 == 0 && ix->d <= 1024) {
            AT_CUDA_OK(cudaMalloc(&ix->op, tc_operand_bytes(ix->d, ktiles)));
            if (!ix->tc_max) {
                AT_CUDA_OK(cudaMalloc(&ix->tc_max, 4 * sizeof(unsigned int)));
                AT_CUDA_OK(cudaMemsetAsync(ix->tc_max, 0, 4 * sizeof(unsigned int), st));
            }
            if (!ix->shift) {
                AT_CUDA_OK(cudaMalloc(&ix->shift, ix->d * sizeof(float)));
                AT_CUDA_OK(cudaMemsetAsync(ix->shift, 0, ix->d * sizeof(float), st));
            }
            if (!ix->tc_counters) {
                AT_CUDA_OK(cudaMalloc(&ix->tc_counters, 8 * sizeof(unsigned long long)));
                AT_CUDA_OK(cudaMemsetAsync(ix->tc_counters, 0, 8 * sizeof(unsigned long long), st));
            }
        }
        ix->kcap = k;
    }
    ix->k = k;
    ix->ktiles = (k + 127) / 128;
    if (centroids != ix->c)
        AT_CUDA_OK(cudaMemcpyAsync(ix->c, centroids, sizeof(float) * (size_t)k * ix->d, cudaMemcpyDeviceToDevice, st));
    return index_refresh(ix, st);
}

int at_index_ntotal(const at_index *ix) { return ix ? ix->k : 0; }

int at_index_set_tc_mode(at_index *ix, int mode) {
    AT_REQUIRE(ix && mode >= 0 && mode <= 2, "at_index_set_tc_mode: bad arguments");
    ix->tc_mode = mode;
    return AT_OK;
}
const float *at_index_centroids(const at_index *ix) { return ix ? ix->c : nullptr; }

int at_index_tc_stats(at_index *ix, uint64_t out[2]) {
    AT_REQUIRE(ix && out, "at_index_tc_stats: bad arguments");
    out[0] = out[1] = 0;
    if (!ix->tc_counters) return AT_OK;
    AT_CUDA_OK(cudaDeviceSynchronize());
    unsigned long long h[8];
    AT_CUDA_OK(cudaMemcpy(h, ix->tc_counters, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h[0], out[1] = h[1];
    if (getenv("AT_TC_DEBUG")) fprintf(stderr, "tc counters: recheck %llu full %llu !groups %llu !sib %llu fallback %llu !third %llu\n", h[0], h[1], h[2], h[3], h[4], h[5]);
    return AT_OK;
}

int at_index_search(at_index *ix, const float *x, int64_t n, int l2norm_rows, int algo, int32_t *labels32,
                    int64_t *labels64, float *dist, void *stream) {
    AT_REQUIRE(ix && (x || n == 0) && n >= 0, "at_index_search: bad arguments");
    AT_REQUIRE(ix->k > 0, "at_index_search: index is empty");
    return index_search(ix, x, n, l2norm_rows, algo, labels32, labels64, dist, 1, nullptr, (cudaStream_t)stream);
}

// ----------------------------------------------------------------------------------- k-means
int at_kmeans_create(int d, int k, at_kmeans **out) {
    AT_REQUIRE(out && d > 0 && k > 0, "at_kmeans_create: bad arguments");
    at_kmeans *km = new (std::nothrow) at_kmeans();
    if (!km) return AT_ERR_NOMEM;
    km->d = d, km->k = k;
    int rc = at_index_create(d, &km->index);
    if (rc != AT_OK) {
        delete km;
        return rc;
    }
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&km->off, sizeof(int64_t) * (size_t)(k + 1));
    if (e == cudaSuccess) e = cudaMalloc(&km->cursor, sizeof(unsigned long long) * (size_t)k);
    if (e == cudaSuccess) e = cudaMalloc(&km->hassign, sizeof(float) * (size_t)k);
    if (e == cudaSuccess) e = cudaMalloc(&km->newc, sizeof(float) * (size_t)k * d);
    if (e == cudaSuccess) e = cudaMalloc(&km->fin, sizeof(double) * 3 * (size_t)((k + FIN_WARPS - 1) / FIN_WARPS));
    if (e != cudaSuccess) {
        set_error("at_kmeans_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        at_kmeans_destroy(km);
        return AT_ERR_CUDA;
    }
    *out = km;
    return AT_OK;
}

int at_kmeans_destroy(at_kmeans *km) {
    if (!km) return AT_OK;
    at_index_destroy(km->index);
    cudaFree(km->labels), cudaFree(km->order);
    cudaFree(km->off), cudaFree(km->cursor), cudaFree(km->hassign), cudaFree(km->newc), cudaFree(km->fin);
    cudaFree(km->rows_sx), cudaFree(km->rows_shift);
    tc_rows_free(&km->rows);
    cudaFree(km->lacc), cudaFree(km->prev), cudaFree(km->d_row), cudaFree(km->d_lab), cudaFree(km->d_order);
    cudaFree(km->d_count), cudaFree(km->d_hist);
    delete km;
    return AT_OK;
}

int at_kmeans_set_incremental(at_kmeans *km, int on) {
    AT_REQUIRE(km, "at_kmeans_set_incremental: bad arguments");
    km->incremental_on = on != 0;
    km->prev_valid = false;
    return AT_OK;
}

int at_kmeans_invalidate(at_kmeans *km) {
    AT_REQUIRE(km, "at_kmeans_invalidate: bad arguments");
    km->rows_valid = false;   // the next accumulate rebuilds the fp16 row image ...
    km->prev_valid = false;   // ... and regroups every row
    return AT_OK;
}

int at_kmeans_set_centroids(at_kmeans *km, const float *centroids, void *stream) {
    AT_REQUIRE(km && centroids, "at_kmeans_set_centroids: bad arguments");
    km->prev_valid = false;   // new centroids from outside: the next accumulate regroups every row
    return at_index_set_centroids(km->index, centroids, km->k, stream);
}

const float *at_kmeans_centroids(const at_kmeans *km) { return km ? km->index->c : nullptr; }

int at_kmeans_get_centroids(const at_kmeans *km, float *out, void *stream) {
    AT_REQUIRE(km && out && km->index->k == km->k, "at_kmeans_get_centroids: bad arguments or centroids not set");
    AT_CUDA_OK(cudaMemcpyAsync(out, km->index->c, sizeof(float) * (size_t)km->k * km->d, cudaMemcpyDeviceToDevice,
                               (cudaStream_t)stream));
    return AT_OK;
}

static int ceil_log2_d(double v) {
    int e = 0;
    if (!(v > 0)) return 0;
    frexp(v, &e);  // v = m * 2^e, m in [0.5, 1)
    return e;      // v < 2^e
}

__global__ void k_set_float(float *p, float v) { *p = v; }

// Stream-ordered form: nothing here waits for the device, so a caller can keep copies / kernels of the NEXT training set in
// flight on other streams while this one starts (HotPath.run_host_stream).
int at_kmeans_begin_on(at_kmeans *km, float max_abs, int64_t n_total, void *stream) {
    AT_REQUIRE(km && n_total > 0 && max_abs >= 0 && isfinite(max_abs), "at_kmeans_begin: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = ceil_log2_d((double)n_total + 1.0);
    int mb = ceil_log2_d((double)max_abs);
    km->e_sum = 61 - nb - mb;  // |x * 2^e| < 2^(61-nb); n_total rows sum below 2^61
    if (km->e_sum > 100) km->e_sum = 100;
    double maxd = 4.0 * km->d * (double)max_abs * (double)max_abs;  // |x - c|^2 <= d (2 max_abs)^2
    km->e_obj = 61 - nb - ceil_log2_d(maxd);
    if (km->e_obj > 100) km->e_obj = 100;
    km->n_total = n_total;
    km->begun = true;
    // tensor path: fix the scale of the rows' fp16 operand image (Sx max|x| in [64, 128)) and re-derive the centroid
    // operands for it; the image itself is built by the first accumulate
    km->rows_valid = false;
    km->prev_valid = false;
    if (km->d % 64 == 0 && km->d <= 1024) {
        float sx = 1.0f;
        if (max_abs > 0.f) {
            int ex;
            frexpf(max_abs, &ex);       // max_abs = f * 2^ex, f in [0.5, 1)
            int e = 7 - ex;             // max_abs * 2^e in [64, 128)
            if (e > 60) e = 60;
            if (e < -60) e = -60;
            sx = ldexpf(1.0f, e);
        }
        if (!km->rows_sx) AT_CUDA_OK(cudaMalloc(&km->rows_sx, sizeof(float)));
        k_set_float<<<1, 1, 0, st>>>(km->rows_sx, sx);
        AT_LAUNCH_OK();
        km->index->ext_sx = km->rows_sx;
        km->index->ext_shift = nullptr;   // until the first accumulate builds the image and fixes its centre
        if (km->index->k > 0 && (assign_tc_supported(km->index) || assign_tc_wide_supported(km->index))) {
            int rc = index_refresh(km->index, st);
            if (rc != AT_OK) return rc;
        }
    }
    return AT_OK;
}

int at_kmeans_begin(at_kmeans *km, float max_abs, int64_t n_total) {
    // legacy form: ordered against everything on the device (callers on the legacy default stream)
    AT_CUDA_OK(cudaDeviceSynchronize());
    int rc = at_kmeans_begin_on(km, max_abs, n_total, nullptr);
    if (rc != AT_OK) return rc;
    AT_CUDA_OK(cudaDeviceSynchronize());
    return AT_OK;
}

int64_t at_kmeans_accum_words(const at_kmeans *km) { return km ? (int64_t)km->k * km->d + km->k + 1 : 0; }

static int km_reserve(at_kmeans *km, int64_t n, cudaStream_t st) {
    if (n <= km->ncap) return AT_OK;
    AT_CUDA_OK(cudaStreamSynchronize(st));
    cudaFree(km->labels), cudaFree(km->order);
    cudaFree(km->prev), cudaFree(km->d_row), cudaFree(km->d_lab), cudaFree(km->d_order);
    km->labels = km->order = nullptr, km->ncap = 0;
    km->prev = km->d_row = km->d_lab = km->d_order = nullptr, km->prev_valid = false;
    AT_CUDA_OK(cudaMalloc(&km->labels, sizeof(int32_t) * (size_t)n));
    AT_CUDA_OK(cudaMalloc(&km->order, sizeof(int32_t) * (size_t)n));
    AT_CUDA_OK(cudaMalloc(&km->prev, sizeof(int32_t) * (size_t)n));
    AT_CUDA_OK(cudaMalloc(&km->d_row, sizeof(int32_t) * (size_t)n * 2));
    AT_CUDA_OK(cudaMalloc(&km->d_lab, sizeof(int32_t) * (size_t)n * 2));
    AT_CUDA_OK(cudaMalloc(&km->d_order, sizeof(int32_t) * (size_t)n * 2));
    if (!km->d_count) AT_CUDA_OK(cudaMalloc(&km->d_count, sizeof(unsigned int)));
    if (!km->d_hist) AT_CUDA_OK(cudaMalloc(&km->d_hist, sizeof(unsigned long long) * (size_t)km->k));
    if (!km->lacc) AT_CUDA_OK(cudaMalloc(&km->lacc, sizeof(unsigned long long) * ((size_t)km->k * km->d + km->k + 1)));
    km->ncap = n;
    return AT_OK;
}

int at_kmeans_accumulate(at_kmeans *km, const float *x, int64_t n_local, int l2norm_rows, int algo,
                         int64_t *accum, int32_t *labels32, void *stream) {
    AT_REQUIRE(km && accum && n_local >= 0 && (x || n_local == 0), "at_kmeans_accumulate: bad arguments");
    AT_REQUIRE(km->begun, "at_kmeans_accumulate: call at_kmeans_begin first");
    AT_REQUIRE(km->index->k == km->k, "at_kmeans_accumulate: centroids not set");
    AT_REQUIRE(!l2norm_rows, "at_kmeans_accumulate: normalise once with at_row_l2norm (rows are re-read every iteration)");
    AT_REQUIRE(n_local < (1LL << 31), "at_kmeans_accumulate: more than 2^31 local rows");
    cudaStream_t st = (cudaStream_t)stream;
    const int k = km->k, d = km->d;
    const int64_t kd = (int64_t)k * d;
    if (n_local == 0) {
        AT_CUDA_OK(cudaMemsetAsync(accum, 0, sizeof(int64_t) * (size_t)(kd + k + 1), st));
        return AT_OK;
    }
    int rc = km_reserve(km, n_local, st);
    if (rc != AT_OK) return rc;
    int32_t *labels = labels32 ? labels32 : km->labels;
    // tensor path: the operand image of the rows is built once and re-used while (x, n_local) stay the same
    at_tc_rows *rows = nullptr;
    const bool tc_any = assign_tc_supported(km->index) || assign_tc_wide_supported(km->index);
    const bool tc = algo == AT_ALGO_TENSOR || (algo == AT_ALGO_AUTO && tc_any && km->index->k >= 64);
    if (tc && tc_any && km->rows_sx) {
        if (!km->rows_valid || km->rows.x != x || km->rows.n != n_local) {
            // centre of the image = mean of the centroids it is first searched with, fixed until the image is rebuilt;
            // the index's operands are re-derived for it
            if (!km->rows_shift) AT_CUDA_OK(cudaMalloc(&km->rows_shift, km->d * sizeof(float)));
            rc = tc_mean(km->index->c, km->k, km->d, km->rows_shift, st);
            if (rc != AT_OK) return rc;
            km->index->ext_shift = km->rows_shift;
            rc = index_refresh(km->index, st);
            if (rc != AT_OK) return rc;
            rc = tc_rows_build(&km->rows, x, n_local, 0, km->rows_sx, km->rows_shift, st, km->d);
            if (rc != AT_OK) return rc;
            km->rows_valid = true;
        }
        rows = &km->rows;
    }
    rc = index_search(km->index, x, n_local, 0, algo, labels, nullptr, nullptr, 0, rows, st);
    if (rc != AT_OK) return rc;
    unsigned long long *acc = (unsigned long long *)accum;
    unsigned long long *lacc = km->lacc;
    ProfScope prof(PROF_UPDATE, st);
    const int blocks = sm_count() * 8;
    const int gblocks = sm_count() * 8;  // 8 warps per block
    const float scale = ldexpf(1.0f, km->e_sum);
    const float obj_scale = ldexpf(1.0f, km->e_obj);
    const int dpl = (d + 31) / 32;
    const bool incremental = km->incremental_on && km->prev_valid && km->prev_x == x && km->prev_n == n_local;
    if (!incremental) {
        // every row: counts, cluster-grouped order, exact gather-sum into the persistent local accumulator
        AT_CUDA_OK(cudaMemsetAsync(lacc, 0, sizeof(int64_t) * (size_t)(kd + k + 1), st));
        k_counts<<<blocks, 256, 0, st>>>(labels, n_local, lacc + kd);
        AT_LAUNCH_OK();
        k_scan_counts<<<1, 1024, 0, st>>>(lacc + kd, k, km->off, km->cursor);
        AT_LAUNCH_OK();
        k_place<<<blocks, 256, 0, st>>>(labels, n_local, km->cursor, km->order);
        AT_LAUNCH_OK();
        if (dpl <= 1)
            k_gather_sum<1><<<gblocks, 256, 0, st>>>(x, d, km->order, km->off, k, n_local, scale, lacc, nullptr, nullptr,
                                                         lacc + kd + k, obj_scale);
        else if (dpl == 2)
            k_gather_sum<2><<<gblocks, 256, 0, st>>>(x, d, km->order, km->off, k, n_local, scale, lacc, nullptr, nullptr,
                                                         lacc + kd + k, obj_scale);
        else if (dpl <= 4)
            k_gather_sum<4><<<gblocks, 256, 0, st>>>(x, d, km->order, km->off, k, n_local, scale, lacc, nullptr, nullptr,
                                                         lacc + kd + k, obj_scale);
        else if (dpl <= 8)
            k_gather_sum<8, 2><<<gblocks, 256, 0, st>>>(x, d, km->order, km->off, k, n_local, scale, lacc, nullptr, nullptr,
                                                            lacc + kd + k, obj_scale);
        else if (dpl <= 20)
            k_gather_sum<20, 1><<<gblocks, 256, 0, st>>>(x, d, km->order, km->off, k, n_local, scale, lacc, nullptr, nullptr,
                                                             lacc + kd + k, obj_scale);
        else
            k_gather_sum<32, 1><<<gblocks, 256, 0, st>>>(x, d, km->order, km->off, k, n_local, scale, lacc, nullptr, nullptr,
                                                             lacc + kd + k, obj_scale);
        AT_LAUNCH_OK();
        AT_CUDA_OK(cudaMemcpyAsync(km->prev, labels, sizeof(int32_t) * (size_t)n_local, cudaMemcpyDeviceToDevice, st));
        km->prev_valid = true, km->prev_x = x, km->prev_n = n_local;
    } else {
        // only the rows whose label changed: two delta items each, grouped by cluster and summed exactly like the rest
        AT_CUDA_OK(cudaMemsetAsync(km->d_count, 0, sizeof(unsigned int), st));
        AT_CUDA_OK(cudaMemsetAsync(km->d_hist, 0, sizeof(unsigned long long) * (size_t)k, st));
        k_diff<<<blocks, 256, 0, st>>>(labels, km->prev, n_local, km->d_row, km->d_lab, km->d_count);
        AT_LAUNCH_OK();
        k_counts<<<blocks, 256, 0, st>>>(km->d_lab, 0, km->d_hist, km->d_count, 2, km->d_row, lacc + kd);
        AT_LAUNCH_OK();
        k_scan_counts<<<1, 1024, 0, st>>>(km->d_hist, k, km->off, km->cursor);
        AT_LAUNCH_OK();
        k_place<<<blocks, 256, 0, st>>>(km->d_lab, 0, km->cursor, km->d_order, km->d_count, 2);
        AT_LAUNCH_OK();
        if (dpl <= 1)
            k_gather_sum<1><<<gblocks, 256, 0, st>>>(x, d, km->d_order, km->off, k, 0, scale, lacc, km->d_row, km->d_count);
        else if (dpl == 2)
            k_gather_sum<2><<<gblocks, 256, 0, st>>>(x, d, km->d_order, km->off, k, 0, scale, lacc, km->d_row, km->d_count);
        else if (dpl <= 4)
            k_gather_sum<4><<<gblocks, 256, 0, st>>>(x, d, km->d_order, km->off, k, 0, scale, lacc, km->d_row, km->d_count);
        else if (dpl <= 8)
            k_gather_sum<8, 2><<<gblocks, 256, 0, st>>>(x, d, km->d_order, km->off, k, 0, scale, lacc, km->d_row, km->d_count);
        else if (dpl <= 20)
            k_gather_sum<20, 1><<<gblocks, 256, 0, st>>>(x, d, km->d_order, km->off, k, 0, scale, lacc, km->d_row, km->d_count);
        else
            k_gather_sum<32, 1><<<gblocks, 256, 0, st>>>(x, d, km->d_order, km->off, k, 0, scale, lacc, km->d_row, km->d_count);
        AT_LAUNCH_OK();
    }
    AT_CUDA_OK(cudaMemcpyAsync(acc, lacc, sizeof(int64_t) * (size_t)(kd + k + 1), cudaMemcpyDeviceToDevice, st));
    return AT_OK;
}

int at_index_search_trained_rows(at_index *ix, at_kmeans *km, const float *x, int64_t n, int l2norm_rows, int32_t *labels32,
                                 int64_t *labels64, float *dist, void *stream) {
    AT_REQUIRE(ix && km && x && n > 0, "at_index_search_trained_rows: bad arguments");
    AT_REQUIRE(ix->k > 0 && assign_tc_supported(ix), "at_index_search_trained_rows: needs d == 64 and 16 <= k <= 65536");
    AT_REQUIRE(km->rows_valid && km->rows.n == n && km->rows_sx && km->d == ix->d,
               "at_index_search_trained_rows: the k-means object holds no operand image of %lld rows", (long long)n);
    cudaStream_t st = (cudaStream_t)stream;
    // the image carries its own scale: re-derive this index's operands for it, search, and put the index back
    ix->ext_sx = km->rows_sx;
    ix->ext_shift = km->rows_shift;
    int rc = index_refresh(ix, st);
    if (rc == AT_OK) {
        ProfScope prof(PROF_SEARCH, st);
        rc = assign_tc_search(ix, x, n, l2norm_rows, labels32, labels64, dist, 1, &km->rows, st);
    }
    ix->ext_sx = nullptr;
    ix->ext_shift = nullptr;
    const int rc2 = index_refresh(ix, st);
    return rc != AT_OK ? rc : rc2;
}

int at_kmeans_finalize(at_kmeans *km, const int64_t *accum, int64_t n_total, float *stats, void *stream) {
    AT_REQUIRE(km && accum && n_total > 0, "at_kmeans_finalize: bad arguments");
    AT_REQUIRE(km->begun, "at_kmeans_finalize: call at_kmeans_begin first");
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(PROF_FINALIZE, st);
    const int nblk = (km->k + FIN_WARPS - 1) / FIN_WARPS;
    k_finalize_div<<<nblk, FIN_WARPS * 32, 0, st>>>((const long long *)accum, km->k, km->d, ldexp(1.0, -km->e_sum),
                                                  km->index->c, km->newc, km->hassign, km->fin);
    AT_LAUNCH_OK();
    k_finalize_split<<<1, 256, 0, st>>>((const long long *)accum, km->k, km->d, n_total, ldexp(1.0, -km->e_obj), km->fin,
                                        nblk, km->newc, km->hassign, stats);
    AT_LAUNCH_OK();
    return at_index_set_centroids(km->index, km->newc, km->k, st);
}

}  // extern "C"
