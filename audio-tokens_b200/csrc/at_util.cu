// libat_b200: error plumbing, device info, bincount, abs-max, synthetic clips.
#include "at_common.cuh"

#include <string.h>
#include <random>
#include <vector>

namespace at {

static thread_local char g_err[512] = "";
int64_t g_launches = 0;
bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof[PROF_TAGS];
static cudaEvent_t g_prof_open[PROF_TAGS];

void prof_begin(int tag, cudaStream_t st) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof_open[tag] = e;
}
void prof_end(int tag, cudaStream_t st) {
    cudaEvent_t e;
    if (!g_prof_open[tag] || cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_prof[tag].push_back({g_prof_open[tag], e});
    g_prof_open[tag] = nullptr;
}

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return dev < 0 ? 0 : (dev >= MAX_DEVICES ? MAX_DEVICES - 1 : dev);
}

int sm_count() {
    static int cached[MAX_DEVICES] = {};   // 0 = not asked yet
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (dev >= 0 && dev < MAX_DEVICES && cached[dev] > 0) return cached[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (dev >= 0 && dev < MAX_DEVICES) cached[dev] = n;
    return n;
}

__global__ void k_absmax(const float *__restrict__ x, int64_t n, float *__restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax((int *)out, __float_as_int(m));
}

__global__ void k_bincount(const int32_t *__restrict__ labels, int64_t n, int k,
                           unsigned long long *__restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int l = labels[i];
        if (l >= 0 && l < k) atomicAdd(&counts[l], 1ULL);
    }
}

// ---------------------------------------------------------------------------------------------
// Synthetic clips: integer-only twin of oracle/synth_ref.py.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

struct SynthPartial {
    uint32_t hp, inc, phi0;
    int amp;
};

__global__ void k_synth(uint32_t seed, int64_t first, int64_t n_samples, const int16_t *__restrict__ table,
                        float *__restrict__ out) {
    __shared__ SynthPartial part[8];
    __shared__ int s_P, s_noise, s_gain;
    __shared__ uint32_t s_h0;
    __shared__ int16_t s_tab[4096];
    int64_t clip = blockIdx.y;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_tab[i] = table[i];
    if (threadIdx.x == 0) {
        uint32_t h0 = hash32(seed * 0x9E3779B1u + (uint32_t)(first + clip));
        h0 = hash32(h0 ^ 0x85EBCA6Bu);
        s_h0 = h0;
        s_P = 3 + (int)(hash32(h0 + 1u) % 6u);
        s_noise = 16 + (int)(hash32(h0 + 2u) % 240u);
        s_gain = 96 + (int)(hash32(h0 + 3u) % 160u);
        for (int p = 0; p < s_P; p++) {
            uint32_t hp = hash32(h0 + 16u + (uint32_t)p);
            uint32_t hq = hash32(hp);
            part[p].hp = hp;
            part[p].inc = (15600000u + ((hq >> 8) & 0xFFFFFFu)) << (hp % 7u);
            part[p].phi0 = hash32(hp + 0x1234567u);
            part[p].amp = 512 + (int)(hash32(hq + 7u) % 3584u);
        }
    }
    __syncthreads();
    const int P = s_P, na = s_noise, gain = s_gain;
    const uint32_t h0 = s_h0;
    float *dst = out + clip * n_samples;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_samples; n += (int64_t)gridDim.x * blockDim.x) {
        uint32_t un = (uint32_t)n;
        uint32_t seg = un >> 13;
        int r = (int)(un & 8191u);
        int total = 0;
        for (int p = 0; p < P; p++) {
            uint32_t phase = part[p].phi0 + un * part[p].inc;
            int s = s_tab[phase >> 20];
            uint32_t hp = part[p].hp;
            int e0 = (int)(hash32((hp ^ (seg * 0x9E3779B1u)) + 0x55u) % 384u) - 128;
            int e1 = (int)(hash32((hp ^ ((seg + 1u) * 0x9E3779B1u)) + 0x55u) % 384u) - 128;
            e0 = e0 < 0 ? 0 : e0;
            e1 = e1 < 0 ? 0 : e1;
            int env = (e0 * (8192 - r) + e1 * r) >> 13;
            total += (((s * part[p].amp) >> 12) * env) >> 8;
        }
        total = (total * gain) >> 9;
        uint32_t hn = hash32(h0 ^ hash32(un + 0x68E31DA4u));
        int nz = (int)(hn % (uint32_t)(2 * na + 1)) - na;
        int v = total + nz;
        v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
        dst[n] = (float)v * (1.0f / 32768.0f);
    }
}

}  // namespace at

using namespace at;

// 16-bit PCM -> fp32 in [-1, 1): sample / 32768, exactly what torchaudio.load(normalize=True) returns for a 16-bit file
// (processors/spectrogram_generator.py:99).  Eight samples per thread step when both pointers allow 16-byte accesses.
__global__ void __launch_bounds__(256) k_pcm16_to_f32(const int16_t *__restrict__ pcm, int64_t n, float *__restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(pcm) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t n8 = vec ? n >> 3 : 0;
    for (int64_t i = tid; i < n8; i += nth) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(pcm) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            f[2 * j] = (float)(int16_t)(w[j] & 0xFFFFu) * (1.0f / 32768.0f);
            f[2 * j + 1] = (float)(int16_t)(w[j] >> 16) * (1.0f / 32768.0f);
        }
        float4 *o = reinterpret_cast<float4 *>(out) + 2 * i;
        o[0] = make_float4(f[0], f[1], f[2], f[3]);
        o[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
    for (int64_t i = (n8 << 3) + tid; i < n; i += nth) out[i] = (float)pcm[i] * (1.0f / 32768.0f);
}

// AmplitudeToDB.forward (torchaudio functional.amplitude_to_DB): multiplier * log10(max(x, amin)) - multiplier * db_multiplier,
// through MUFU.LG2 like the fused mel kernel; optional top_db clamp against a device-resident maximum.
__global__ void __launch_bounds__(256) k_amp_to_db(const float *__restrict__ x, int64_t n, float mul_log2, float amin,
                                                   float sub, float *__restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t n4 = vec ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(x) + i);
        v.x = mul_log2 * __log2f(fmaxf(v.x, amin)) - sub, v.y = mul_log2 * __log2f(fmaxf(v.y, amin)) - sub;
        v.z = mul_log2 * __log2f(fmaxf(v.z, amin)) - sub, v.w = mul_log2 * __log2f(fmaxf(v.w, amin)) - sub;
        reinterpret_cast<float4 *>(out)[i] = v;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += nth) out[i] = mul_log2 * __log2f(fmaxf(x[i], amin)) - sub;
}

extern "C" {

int at_version(void) { return AT_B200_VERSION; }

int at_amplitude_to_db(const float *x, int64_t n, float multiplier, float amin, float db_multiplier, float *out, void *stream) {
    AT_REQUIRE(x && out && n >= 0 && amin > 0.f, "at_amplitude_to_db: bad arguments");
    if (n == 0) return AT_OK;
    int64_t want = ceil_div(n, 256 * 8);
    int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
    if (blocks < 1) blocks = 1;
    // multiplier * log10(v) = multiplier * log10(2) * log2(v)
    k_amp_to_db<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, multiplier * 0.301029995663981195f, amin,
                                                         multiplier * db_multiplier, out);
    AT_LAUNCH_OK();
    return AT_OK;
}

const char *at_last_error(void) { return g_err; }

int64_t at_kernel_launches(void) { return g_launches; }

int at_profile_enable(int on) {
    g_prof_on = on != 0;
    return AT_OK;
}

int at_profile_summary(int tag, int64_t *launches, double *total_ms) {
    AT_REQUIRE(tag >= 0 && tag < PROF_TAGS && launches && total_ms, "at_profile_summary: bad arguments");
    double tot = 0;
    for (auto &p : g_prof[tag]) {
        AT_CUDA_OK(cudaEventSynchronize(p.second));
        float ms = 0;
        AT_CUDA_OK(cudaEventElapsedTime(&ms, p.first, p.second));
        tot += ms;
        cudaEventDestroy(p.first), cudaEventDestroy(p.second);
    }
    *launches = (int64_t)g_prof[tag].size();
    *total_ms = tot;
    g_prof[tag].clear();
    return AT_OK;
}

int at_device_info(int *sm, int *major, int *minor) {
    int dev = 0;
    AT_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    AT_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    return AT_OK;
}

int at_absmax(const float *x, int64_t n_elems, float *out_dev, void *stream) {
    AT_REQUIRE(x && out_dev && n_elems >= 0, "at_absmax: bad arguments");
    AT_CUDA_OK(cudaMemsetAsync(out_dev, 0, sizeof(float), (cudaStream_t)stream));
    if (n_elems == 0) return AT_OK;
    int blocks = (int)(ceil_div(n_elems, 256 * 8) < (int64_t)sm_count() * 8 ? ceil_div(n_elems, 256 * 8) : (int64_t)sm_count() * 8);
    if (blocks < 1) blocks = 1;
    k_absmax<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n_elems, out_dev);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_pcm16_to_f32(const int16_t *pcm, int64_t n, float *out, void *stream) {
    AT_REQUIRE(pcm && out && n >= 0, "at_pcm16_to_f32: bad arguments");
    if (n == 0) return AT_OK;
    int blocks = sm_count() * 16;
    if (blocks < 1) blocks = 1;
    k_pcm16_to_f32<<<blocks, 256, 0, (cudaStream_t)stream>>>(pcm, n, out);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_bincount(const int32_t *labels, int64_t n, int k, int64_t *counts, void *stream) {
    AT_REQUIRE(labels && counts && n >= 0 && k > 0, "at_bincount: bad arguments");
    AT_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)k, (cudaStream_t)stream));
    if (n == 0) return AT_OK;
    int blocks = sm_count() * 4;
    if (blocks < 1) blocks = 1;
    k_bincount<<<blocks, 256, 0, (cudaStream_t)stream>>>(labels, n, k, (unsigned long long *)counts);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_rand_perm_host(int32_t *perm, int64_t n, int64_t seed) {
    AT_REQUIRE(perm && n >= 0 && n < (1LL << 31), "at_rand_perm_host: bad arguments");
    std::mt19937 mt((unsigned)seed);
    for (int64_t i = 0; i < n; i++) perm[i] = (int32_t)i;
    for (int64_t i = 0; i + 1 < n; i++) {
        int64_t i2 = i + (int64_t)((uint64_t)mt() % (uint64_t)(n - i));
        int32_t t = perm[i];
        perm[i] = perm[i2];
        perm[i2] = t;
    }
    return AT_OK;
}

int at_synth_clips(uint32_t seed, int64_t first_index, int64_t count, int64_t n_samples,
                   const int16_t *sine_table4096, float *out, void *stream) {
    AT_REQUIRE(sine_table4096 && out && count >= 0 && n_samples > 0, "at_synth_clips: bad arguments");
    AT_REQUIRE(n_samples < (1LL << 31), "at_synth_clips: n_samples too large");
    int64_t done = 0;
    while (done < count) {  // gridDim.y limit
        int64_t chunk = count - done < 32768 ? count - done : 32768;
        int bx = (int)(ceil_div(n_samples, 256 * 16) < 64 ? ceil_div(n_samples, 256 * 16) : 64);
        dim3 grid(bx, (unsigned)chunk);
        k_synth<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, first_index + done, n_samples, sine_table4096,
                                                          out + done * n_samples);
        AT_LAUNCH_OK();
        done += chunk;
    }
    return AT_OK;
}

}  // extern "C"
