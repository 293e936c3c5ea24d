// Shared helpers for the libat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/audio_tokens_b200.h"

namespace at {

void set_error(const char *fmt, ...);

#define AT_CUDA_OK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            at::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                          __LINE__);                                                           \
            return AT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define AT_REQUIRE(cond, ...)                                                                  \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            at::set_error(__VA_ARGS__);                                                        \
            return AT_ERR_INVALID;                                                             \
        }                                                                                      \
    } while (0)

#define AT_LAUNCH_OK()                                                                         \
    do {                                                                                       \
        at::g_launches++;                                                                      \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            at::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),          \
                          __FILE__, __LINE__);                                                 \
            return AT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

constexpr int MAX_DEVICES = 64;
int current_device();  // cudaGetDevice clamped to [0, MAX_DEVICES)
int sm_count();  // of the current device (cached per device); 0 when no device
extern int64_t g_launches;

// Event-pair profiling (at_profile_*): ProfScope brackets the launches issued while it is alive.
enum { PROF_SEARCH = 0, PROF_MEL = 1, PROF_UPDATE = 2, PROF_FINALIZE = 3, PROF_TAGS = 4 };
extern bool g_prof_on;
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);
struct ProfScope {
    int tag;
    cudaStream_t st;
    ProfScope(int t, cudaStream_t s) : tag(t), st(s) { if (g_prof_on) prof_begin(tag, st); }
    ~ProfScope() { if (g_prof_on) prof_end(tag, st); }
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// FAISS's expanded distance, evaluated the same way in every kernel of this library so that the exact
// SIMT path, the tensor path's re-check and the objective all agree bit for bit:
//   ip  = sum_t x[t]*c[t]   (sequential fp32 FMA chain over t, four interleaved partial chains)
//   dis = (|x|^2 + |c|^2) - 2*ip, clamped at 0       (faiss/utils/distances.cpp exhaustive_L2sqr_blas)
__device__ __forceinline__ float l2_expanded(float xn, float cn, float ip) {
    float dis = __fsub_rn(__fadd_rn(xn, cn), 2.0f * ip);
    return dis < 0.f ? 0.f : dis;
}

// Sum of squares / dot product over d elements with a fixed association: four interleaved FMA chains
// (elements t, t+1, t+2, t+3 of each group of four), combined as (s0+s1)+(s2+s3).
__device__ __forceinline__ float combine4(float s0, float s1, float s2, float s3) {
    return __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
}

}  // namespace at
