// Internal layout of the opaque at_index / at_kmeans objects and the canonical fp32 arithmetic every
// search kernel shares.
#pragma once
#include "at_common.cuh"
#include <cuda_fp16.h>

// Operand image of a set of rows for the tcgen05 search (at_assign_tc.cu: k_tc_rows) + the tail list of one search.
struct at_tc_rows {
    void *img = nullptr;        // (n_pad / 128) tiles of 20,480 bytes
    float *erow = nullptr;      // (n_pad) |Sx delta| per row
    float *xns = nullptr;       // (n_pad) Sx^2 |x|^2 per row
    uint4 *tail = nullptr;      // (n_pad) uncertified rows of the last search with a candidate list, one queue per scanning
                                // warp of k_assign_tc (no shared counter): {row, candidate columns 0|1, 2|3, -}
    uint32_t *full = nullptr;   // (n_pad) uncertified rows that need an exact scan (dense list)
    unsigned long long *keys = nullptr;   // (n_pad, wide rows only) (distance, column) keys of the listed rows' exact scans
    unsigned int *tail_count = nullptr;   // [1] rows in `full`, [2 + q] entries of candidate queue q
    int tail_queues = 0;        // queues allocated in tail_count
    int64_t cap = 0;            // rows allocated (multiple of 256)
    const float *x = nullptr;   // what the image was built from
    int64_t n = 0;
    int l2norm = 0;
    int d = 64;                 // row width the image was allocated for (64, or 64 NS for the wide-row kernel)
};

struct at_index {
    int d = 0;
    int k = 0;      // centroids currently held
    int kcap = 0;   // allocation, in centroids
    float *c = nullptr;    // (kcap, d) fp32 centroids
    float *cn = nullptr;   // (kcap) canonical |c|^2
    // tcgen05 operands (d == 64 only): per 128-centroid tile one 20,480-byte image = fp16(-2*Sc*c) as a 128x64 K-major
    // SWIZZLE_128B tile + a 128x16 no-swizzle tile carrying |c|^2 (at_assign_tc.cu)
    __half *op = nullptr;
    float *tc_scale = nullptr;  // device: {S, max centroid rounding error, 1 / S^2, tau, S max|c|, Sx, S / Sx}
    unsigned int *tc_max = nullptr;  // device: {max |c_ij|, max |c_j|^2} bit patterns (k_centroid_norms -> k_tc_scale)
    const float *ext_sx = nullptr;  // device float: scale of an attached row image (k-means); nullptr = the index's own S
    // Centering: |x - c|^2 = |(x - m) - (c - m)|^2 for any m, and the fp16 rounding errors of the operands scale with
    // |x - m| |c - m| instead of |x| |c| -- m = mean of the centroids (the index's own `shift`, refreshed with the
    // centroids), or the vector an attached row image was built with (`ext_shift`, k-means: fixed for the training run)
    float *shift = nullptr;         // (d) device
    const float *ext_shift = nullptr;
    at_tc_rows rows;            // workspace of one-off searches
    unsigned long long *tc_counters = nullptr;  // device: rows re-checked on their candidate columns, rows scanned exactly (cumulative)
    int tc_mode = 0;            // 0 auto, 1 stream operand tiles, 2 keep them resident when they fit
    int32_t *part_lab = nullptr;  // scratch labels for a distance-only request
    int64_t part_cap = 0;
    int ktiles = 0;
    // side stream + fork / join events of the tensor search's two tail kernels (created on first use)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_failed = false;
    bool side_ok() {
        if (side) return true;
        if (side_failed) return false;
        if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            side = nullptr, side_failed = true;
            return false;
        }
        return true;
    }
};

struct at_kmeans {
    int d = 0, k = 0;
    at_index *index = nullptr;
    // fixed-point accumulation (at_kmeans_begin)
    int e_sum = 0, e_obj = 0;
    int64_t n_total = 0;
    bool begun = false;
    // workspaces sized for the largest n_local seen
    int64_t ncap = 0;
    int32_t *labels = nullptr;
    int32_t *order = nullptr;
    int64_t *off = nullptr;     // k+1
    unsigned long long *cursor = nullptr;  // k
    float *hassign = nullptr;   // k
    float *newc = nullptr;      // (k, d) scratch for finalize
    double *fin = nullptr;      // per-block partials of finalize's statistics
    // tensor path: operand image of the training rows, built at the first accumulate and re-used while the caller
    // passes the same (x, n_local) -- the rows must not change between at_kmeans_begin and the last accumulate
    at_tc_rows rows;
    float *rows_sx = nullptr;   // device float: image scale, fixed by at_kmeans_begin from max |x|
    float *rows_shift = nullptr;   // device (64): centering vector of the image = mean of the centroids when it was built
    bool rows_valid = false;
    // incremental update: the local sums / counts persist between accumulate calls (exact integers), so an iteration
    // only moves the rows whose label changed: -x from the old cluster, +x to the new one (same invariant on x as above)
    unsigned long long *lacc = nullptr;    // k*d sums + k counts
    int32_t *prev = nullptr;               // labels of the previous accumulate (ncap)
    const float *prev_x = nullptr;
    int64_t prev_n = 0;
    bool prev_valid = false;
    bool incremental_on = true;
    int32_t *d_row = nullptr, *d_lab = nullptr, *d_order = nullptr;   // 2 items per changed row (2 * ncap)
    unsigned int *d_count = nullptr;       // device: changed rows of the last accumulate
    unsigned long long *d_hist = nullptr;  // k
};

namespace at {

// ---- canonical sums --------------------------------------------------------------------------
// A length-d sum (of squares, or of products) is always associated the same way:
//   16 partials q_l, l = (t/4) % 16, each a sequential FMA chain over its elements in ascending t,
//   combined by the xor butterfly  l^8, l^4, l^2, l^1.
// This shape is what a 16-lane shuffle reduction over float4 chunks produces, is invariant under rotating
// the chunk order, and can be evaluated by one thread with static register indices.
__device__ __forceinline__ float tree16(const float (&q)[16]) {
    float a[8], b[4], c2[2];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __fadd_rn(q[i], q[i + 8]);
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = __fadd_rn(a[i], a[i + 4]);
#pragma unroll
    for (int i = 0; i < 2; i++) c2[i] = __fadd_rn(b[i], b[i + 2]);
    return __fadd_rn(c2[0], c2[1]);
}

__device__ __forceinline__ float half16_sum(float v) {  // xor butterfly over a 16-lane group
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return v;
}

// normalize_vectors' scalar: 1 / (sqrt(S) + 1e-10) is NOT formed; numpy divides by (norm + 1e-10).
__device__ __forceinline__ float l2_denominator(float sumsq) { return __fadd_rn(__fsqrt_rn(sumsq), 1e-10f); }

// Implemented in at_assign_tc.cu.  Returns AT_ERR_UNSUPPORTED when the shape is outside the tensor path.
int assign_tc_prepare(at_index *ix, cudaStream_t st);
int assign_tc_search(at_index *ix, const float *x, int64_t n, int l2norm_rows, int32_t *labels32,
                     int64_t *labels64, float *dist, int exact_dist, at_tc_rows *rows, cudaStream_t st);
int tc_rows_build(at_tc_rows *r, const float *x, int64_t n, int l2norm, const float *sx, const float *shift, cudaStream_t st,
                  int d = 64);
int tc_mean(const float *c, int k, int d, float *shift, cudaStream_t st);   // shift = mean of the k centroids (d = 64 NS)
bool assign_tc_wide_supported(const at_index *ix);   // d = 128 .. 1024, a multiple of 64: the slice-accumulating kernel
size_t tc_operand_bytes(int d, int ktiles);          // bytes of the centroid operand images
int assign_tc_wide_search(at_index *ix, const float *x, int64_t n, int32_t *labels32, int64_t *labels64, float *dist,
                          at_tc_rows *rows, cudaStream_t st);
// at_conv.cu: the exact wide-row kernel over a list of rows (list, *n_list on the device; at most n_max)
int launch_assign_gemm_list(const at_index *ix, const float *x, const uint32_t *list, const unsigned int *n_list, int64_t n_max,
                            unsigned long long *keys, int32_t *l32, int64_t *l64, float *dist, cudaStream_t st);
void tc_rows_free(at_tc_rows *r);
bool assign_tc_supported(const at_index *ix);
// at_conv.cu: exact fp32 search for rows wider than 128 values (pre-normalised rows)
int launch_assign_gemm(const at_index *ix, const float *x, int64_t n, int32_t *l32, int64_t *l64, float *dist, cudaStream_t st);

}  // namespace at
