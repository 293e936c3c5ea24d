// libat_b200, the step before the path: channel mean + sample-rate conversion of a decoded clip.
//
// Replaces SpectrogramGenerator.convert_to_mono + SpectrogramGenerator.resample
// (processors/spectrogram_generator.py:109-121): torch.mean over channels, then torchaudio.transforms.Resample(sr,
// common_sr) with its defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).  The reference builds a new
// Resample module -- i.e. the whole filter bank -- for every clip; here a plan holds the filter bank of one (orig, new)
// pair and a clip is one launch.
//
// Arithmetic of torchaudio (functional.py, _get_sinc_resample_kernel / _apply_sinc_resample_kernel), o = orig / gcd,
// n = new / gcd, width = ceil(6 o / (0.99 min(o, n))), K = 2 width + o taps per phase:
//     y[j n + p] = sum_k kern[p][k] * xpad[j o + k],   xpad = x padded with `width` zeros in front
// The bank is evaluated on the host exactly as torch does (float64 index grid, the phase -p / n in float32, rounded to
// float32 at the end); a phase's taps outside the clamped +-6 zero-crossing support are exact zeros and are skipped.
#include "at_common.cuh"

#include <math.h>
#include <string.h>
#include <new>
#include <numeric>
#include <vector>

struct at_resample_plan {
    int orig = 0, neu = 0;   // after division by the gcd
    int width = 0, taps = 0;
    float *kern = nullptr;   // [neu][taps]
    int2 *range = nullptr;   // [neu] first / one-past-last non-zero tap
    // compact form for the tiled kernel: the non-zero taps of every phase back to back
    float *ckern = nullptr;  // [ctotal]
    int *coff = nullptr;     // [neu] offset of phase p's taps in ckern
    int ctotal = 0, kmin = 0, kmax = 0;   // kmin / kmax: smallest first / largest one-past-last non-zero tap over the phases
};

namespace at {

__global__ void __launch_bounds__(256) k_resample_mono(const float *__restrict__ x, int C, int64_t L, int o, int n, int width,
                                                       int taps, const float *__restrict__ kern,
                                                       const int2 *__restrict__ range, int64_t out_len,
                                                       float *__restrict__ out) {
    x += (int64_t)blockIdx.y * C * L;          // uniform batch: clip blockIdx.y
    out += (int64_t)blockIdx.y * out_len;
    const float inv_note = (float)C;   // torch.mean divides the channel sum by C
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < out_len; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i / n;
        const int p = (int)(i - j * n);
        const int64_t base = j * o - width;
        const int2 r = range[p];
        const float *kp = kern + (size_t)p * taps;
        float acc = 0.f;
        for (int k = r.x; k < r.y; k++) {
            const int64_t s = base + k;
            if (s < 0 || s >= L) continue;
            float m = __ldg(x + s);
            if (C > 1) {
                for (int c = 1; c < C; c++) m += __ldg(x + (int64_t)c * L + s);
                m = __fdiv_rn(m, inv_note);
            }
            acc = fmaf(__ldg(kp + k), m, acc);
        }
        out[i] = acc;
    }
}

// Tiled form: a block takes RS_TILE consecutive output samples of one clip, stages the channel mean of the input span
// they touch in shared memory (every input sample is read from HBM and averaged once, not once per tap), keeps the
// compact filter bank in shared memory, and each thread accumulates its outputs' taps in ascending order.
constexpr int RS_TILE = 2048, RS_THREADS = 256, RS_SPAN_MAX = 8192, RS_CK_MAX = 12288;
__global__ void __launch_bounds__(RS_THREADS) k_resample_mono_tiled(const float *__restrict__ x, int C, int64_t L, int o, int n,
                                                                   int width, const float *__restrict__ ckern,
                                                                   const int *__restrict__ coff,
                                                                   const int2 *__restrict__ range, int ctotal, int kmin,
                                                                   int kmax, int span_cap, int64_t out_len,
                                                                   float *__restrict__ out) {
    extern __shared__ float rs_smem[];
    float *xs = rs_smem;              // span_cap
    float *ks = rs_smem + span_cap;   // ctotal
    x += (int64_t)blockIdx.y * C * L;
    out += (int64_t)blockIdx.y * out_len;
    for (int i = threadIdx.x; i < ctotal; i += RS_THREADS) ks[i] = ckern[i];
    const float cf = (float)C;
    const int64_t ntiles = (out_len + RS_TILE - 1) / RS_TILE;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t i0 = t * RS_TILE, i1 = min(out_len, i0 + RS_TILE);
        const int64_t s0 = (i0 / n) * o - width + kmin;           // first input sample any output of the tile may touch
        const int64_t s1 = ((i1 - 1) / n) * o - width + kmax;     // one past the last
        __syncthreads();   // the previous tile's span has been consumed (and ks is in place)
        for (int64_t s = s0 + threadIdx.x; s < s1; s += RS_THREADS) {
            float m = 0.f;
            if (s >= 0 && s < L) {
                m = __ldg(x + s);
                if (C > 1) {
                    for (int c = 1; c < C; c++) m += __ldg(x + (int64_t)c * L + s);
                    m = __fdiv_rn(m, cf);   // torch.mean divides the channel sum by C
                }
            }
            xs[s - s0] = m;
        }
        __syncthreads();
        const int64_t j0 = i0 / n;
        const int r0 = (int)(i0 - j0 * n), cnt = (int)(i1 - i0);   // 32-bit arithmetic inside the tile
        for (int li = threadIdx.x; li < cnt; li += RS_THREADS) {
            const int jl = (li + r0) / n;
            const int p = li + r0 - jl * n;
            const int2 r = range[p];
            const float *kp = ks + coff[p] - r.x;
            const float *xp = xs + (jl * o - kmin);   // (j0 + jl) o - width - s0
            float acc = 0.f;
            for (int k = r.x; k < r.y; k++) acc = fmaf(kp[k], xp[k], acc);
            out[i0 + li] = acc;
        }
    }
}

}  // namespace at

using namespace at;

extern "C" {

// torchaudio's filter bank for (orig_freq, new_freq), evaluated as _get_sinc_resample_kernel does (host, double).
static void build_bank(int orig_freq, int new_freq, std::vector<float> &kern, std::vector<int2> &range, int &o, int &n,
                       int &width, int &taps) {
    const int g = std::gcd(orig_freq, new_freq);
    o = orig_freq / g, n = new_freq / g;
    const double base = (double)(o < n ? o : n) * 0.99;
    const bool identity = orig_freq == new_freq;   // Resample.forward returns its input unchanged: channel mean only
    width = identity ? 0 : (int)ceil(6.0 * (double)o / base);
    taps = 2 * width + o;
    kern.assign((size_t)n * taps, 0.f);
    range.assign(n, make_int2(0, 0));
    const double scale = base / (double)o;
    for (int ph = 0; ph < n; ph++) {
        const double phase = (double)((float)(-ph) / (float)n);   // int64 tensor / int -> float32 in torch
        int first = taps, last = -1;
        for (int k = 0; k < taps; k++) {
            const double idx = (double)(k - width) / (double)o;
            double t = (phase + idx) * base;
            if (t < -6.0) t = -6.0;
            if (t > 6.0) t = 6.0;
            const double c = cos(t * M_PI / 6.0 / 2.0);
            const double window = c * c;
            t *= M_PI;
            const double s = t == 0.0 ? 1.0 : sin(t) / t;
            const float v = identity ? 1.0f : (float)(s * (window * scale));
            kern[(size_t)ph * taps + k] = v;
            if (v != 0.f) {
                if (k < first) first = k;
                last = k;
            }
        }
        range[ph] = last < 0 ? make_int2(0, 0) : make_int2(first, last + 1);
    }
}

// HOST function, HOST pointers (no device needed): the filter bank a plan for (orig_freq, new_freq) uses.  bank may be
// NULL to query the sizes: [*phases][*taps] floats.
int at_resample_bank_host(int orig_freq, int new_freq, float *bank, int *phases, int *taps_out, int *width_out) {
    AT_REQUIRE(orig_freq > 0 && new_freq > 0, "at_resample_bank_host: bad arguments");
    std::vector<float> kern;
    std::vector<int2> range;
    int o, n, width, taps;
    build_bank(orig_freq, new_freq, kern, range, o, n, width, taps);
    if (phases) *phases = n;
    if (taps_out) *taps_out = taps;
    if (width_out) *width_out = width;
    if (bank) memcpy(bank, kern.data(), sizeof(float) * kern.size());
    return AT_OK;
}

int at_resample_plan_create(int orig_freq, int new_freq, at_resample_plan **plan) {
    AT_REQUIRE(plan && orig_freq > 0 && new_freq > 0, "at_resample_plan_create: bad arguments");
    int dev;
    AT_CUDA_OK(cudaGetDevice(&dev));
    at_resample_plan *p = new (std::nothrow) at_resample_plan();
    if (!p) return AT_ERR_NOMEM;
    std::vector<float> kern;
    std::vector<int2> range;
    int o, n, width, taps;
    build_bank(orig_freq, new_freq, kern, range, o, n, width, taps);
    p->orig = o, p->neu = n, p->width = width, p->taps = taps;
    std::vector<float> ck;
    std::vector<int> coff(n);
    p->kmin = taps, p->kmax = 0;
    for (int ph = 0; ph < n; ph++) {
        coff[ph] = (int)ck.size();
        for (int k = range[ph].x; k < range[ph].y; k++) ck.push_back(kern[(size_t)ph * taps + k]);
        if (range[ph].y > range[ph].x) {
            if (range[ph].x < p->kmin) p->kmin = range[ph].x;
            if (range[ph].y > p->kmax) p->kmax = range[ph].y;
        }
    }
    if (p->kmax == 0) p->kmin = 0;
    if (ck.empty()) ck.push_back(0.f);
    p->ctotal = (int)ck.size();
    cudaError_t e = cudaMalloc(&p->kern, sizeof(float) * kern.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->ckern, sizeof(float) * ck.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->coff, sizeof(int) * coff.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->ckern, ck.data(), sizeof(float) * ck.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->coff, coff.data(), sizeof(int) * coff.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&p->range, sizeof(int2) * range.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->kern, kern.data(), sizeof(float) * kern.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->range, range.data(), sizeof(int2) * range.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        set_error("at_resample_plan_create: %s", cudaGetErrorString(e));
        cudaFree(p->kern), cudaFree(p->range), cudaFree(p->ckern), cudaFree(p->coff);
        delete p;
        return AT_ERR_CUDA;
    }
    *plan = p;
    return AT_OK;
}

int at_resample_plan_destroy(at_resample_plan *p) {
    if (!p) return AT_OK;
    cudaFree(p->kern), cudaFree(p->range), cudaFree(p->ckern), cudaFree(p->coff);
    delete p;
    return AT_OK;
}

int64_t at_resample_out_len(const at_resample_plan *p, int64_t n_in) {
    if (!p || n_in < 0) return -1;
    // torch.ceil(torch.as_tensor(new * length / orig)): Python true division in double, then ceil
    return (int64_t)ceil((double)p->neu * (double)n_in / (double)p->orig);
}

int at_resample_mono_batch(at_resample_plan *p, const float *wave, int channels, int64_t n_in, int B, float *out,
                           void *stream) {
    AT_REQUIRE(p && wave && out && channels > 0 && n_in >= 0 && B >= 0 && B <= 65535, "at_resample_mono_batch: bad arguments");
    const int64_t out_len = at_resample_out_len(p, n_in);
    if (out_len == 0 || B == 0) return AT_OK;
    // input span of one tile: ((RS_TILE - 1) / n + 1) * o + (kmax - kmin) samples
    const int64_t span = ((int64_t)(RS_TILE - 1) / p->neu + 1) * p->orig + (p->kmax - p->kmin);
    if (span <= RS_SPAN_MAX && p->ctotal <= RS_CK_MAX) {
        static bool configured[MAX_DEVICES] = {};   // the attribute is per device
        const int dev = current_device();
        const int span_cap = (int)((span + 3) & ~(int64_t)3);
        const size_t smem = sizeof(float) * (size_t)(span_cap + p->ctotal);
        if (!configured[dev]) {
            AT_CUDA_OK(cudaFuncSetAttribute(k_resample_mono_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(sizeof(float) * (RS_SPAN_MAX + RS_CK_MAX))));
            configured[dev] = true;
        }
        int64_t tiles = ceil_div(out_len, RS_TILE);
        const int64_t capt = ceil_div((int64_t)(sm_count() > 0 ? sm_count() : 1) * 8, B);
        if (tiles > capt) tiles = capt;
        if (tiles < 1) tiles = 1;
        k_resample_mono_tiled<<<dim3((unsigned)tiles, (unsigned)B), RS_THREADS, smem, (cudaStream_t)stream>>>(
            wave, channels, n_in, p->orig, p->neu, p->width, p->ckern, p->coff, p->range, p->ctotal, p->kmin, p->kmax,
            span_cap, out_len, out);
        AT_LAUNCH_OK();
        return AT_OK;
    }
    int64_t blocks = ceil_div(out_len, 256);
    const int64_t cap = ceil_div((int64_t)(sm_count() > 0 ? sm_count() : 1) * 16, B);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_resample_mono<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
        wave, channels, n_in, p->orig, p->neu, p->width, p->taps, p->kern, p->range, out_len, out);
    AT_LAUNCH_OK();
    return AT_OK;
}

int at_resample_mono(at_resample_plan *p, const float *wave, int channels, int64_t n_in, float *out, void *stream) {
    return at_resample_mono_batch(p, wave, channels, n_in, 1, out, stream);
}

}  // extern "C"
