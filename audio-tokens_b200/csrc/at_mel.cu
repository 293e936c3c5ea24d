// libat_b200 stage 1: waveform -> STFT power -> HTK mel -> dB -> per-clip min-max (-> row L2 norm).
//
// Replaces, for a whole batch of clips in one launch, what the reference does one clip at a time through
// torchaudio (processors/spectrogram_generator.py:123-131):
//   MelSpectrogram.forward  -> torch.stft (reflect pad n_fft/2, periodic Hann, onesided) -> |X|^2 -> fb matmul
//   AmplitudeToDB.forward   -> 10*log10(clamp(x, 1e-10))
//   normalize_spectrogram   -> (s - min s) / (max s - min s) over the clip
//   check_for_nan_inf       -> bad flag
//
// Kernel shape (one persistent CTA per SM = four independent groups of 4 warps, clips dealt round-robin to the groups; a
// group works through its clip in batches of 4 jobs and synchronises on its own named barrier, so one group's FFT phase
// overlaps the others' mel / write-out phases):
//   * the sample window of the next batch of frames (7 hops + n_fft samples at n_fft 1024, 18.4 KB) is fetched into
//     shared memory by one cp.async.bulk (TMA engine) while the current batch is in its mel / write-out phases, so
//     overlapping frames are read from HBM once and no warp ever waits on DRAM (reflected edge frames read global memory);
//   * a warp owns one "job": 1024 complex points = G complex FFTs of n_fft points, each packing TWO real
//     frames (frame a -> real part, frame b -> imaginary part), so no arithmetic is spent on the redundant
//     half of a real-input transform; at hop = n_fft / 2 the two frames share half their staged samples and every window
//     value is fetched once for both;
//   * n_fft = 32 * N2 is split Cooley-Tukey style: an N2-point FFT inside each lane's registers, a twiddle
//     multiply, one 32x32 transpose through the warp's private shared-memory tile, a 32-point FFT in registers;
//   * the two frames are separated with one shuffle per bin (Z[k], conj Z[n_fft-k]) and their power written
//     to a shared power tile (rows 16-byte aligned); 8 jobs fill the tile;
//   * the mel projection runs thread-per-(frame, filter) over the tile using the filterbank's sparsity
//     (each triangular filter touches only its own bin range, weights and powers read 128 bits at a time), then
//     10*log10 through MUFU.LG2, a transposed shared tile and coalesced 128-bit stores; min/max are tracked on the way out;
//   * after the clip's last frame the group normalises the clip's tile in place (it is still in L2) and writes the
//     L2-normalised copy the k-means stage consumes.
#include "at_common.cuh"
#include "at_index.cuh"
#include "at_ptx.cuh"

#include <math.h>
#include <new>
#include <vector>

struct at_mel_plan {
    int sample_rate = 0, n_fft = 0, hop = 0, n_mels = 0, normalize = 0;
    int power_out = 0;        // 1: write the mel POWER (MelSpectrogram alone), 0: dB (MelSpectrogram + AmplitudeToDB)
    float *absmax_out = nullptr;   // device float raised (atomicMax) to the largest |element| of the L2-normalised copy
    int log2nf = 0;
    // device constants
    float *win = nullptr;     // n_fft
    float2 *tw = nullptr;     // N2 * 32 inter-pass twiddles [k2][n1]
    int *fstart = nullptr;    // n_mels: first bin of each filter's support
    int *fcnt = nullptr;      // n_mels: bins in the support
    int *woff = nullptr;      // n_mels: offset into wt
    float *wt = nullptr;      // concatenated weights of each filter's support, pre-scaled by 1/4, each filter padded
                              // with zeros to a multiple of 4 entries (16-byte aligned segments)
    int wt_count = 0;         // floats in wt
    // host copies
    std::vector<float> h_win, h_fb;
    // staging for the _host entry point
    float *stage_in[2] = {nullptr, nullptr};
    float *stage_out[2] = {nullptr, nullptr};
    int32_t *stage_bad[2] = {nullptr, nullptr};
    int64_t stage_clips = 0, stage_samples = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
};

namespace at {

// ---------------------------------------------------------------------------------------------
// compile-time twiddles: W_32^m = (c32(m), -s32(m)), m in [0, 16)
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr float c32(int m) {
    switch (m & 31) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323044913f;
        case 2: return 0.92387953251128675613f;
        case 3: return 0.83146961230254523708f;
        case 4: return 0.70710678118654752440f;
        case 5: return 0.55557023301960222474f;
        case 6: return 0.38268343236508977173f;
        case 7: return 0.19509032201612826785f;
        case 8: return 0.0f;
        case 9: return -0.19509032201612826785f;
        case 10: return -0.38268343236508977173f;
        case 11: return -0.55557023301960222474f;
        case 12: return -0.70710678118654752440f;
        case 13: return -0.83146961230254523708f;
        case 14: return -0.92387953251128675613f;
        case 15: return -0.98078528040323044913f;
        default: return 0.0f;
    }
}
// sin(2 pi m / 32) for m in [0, 16): cos(pi/2 - x) below the quarter turn, cos(x - pi/2) above it
__host__ __device__ constexpr float s32(int m) { return m <= 8 ? c32(8 - m) : c32(m - 8); }

__host__ __device__ constexpr int brev(int p, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r |= ((p >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// In-place radix-2 decimation-in-frequency FFT of N points held in registers (all indices static after
// unrolling).  Output position p holds frequency brev(p).
template <int N>
__device__ __forceinline__ void fft_dif(float *re, float *im) {
    constexpr float R = 0.70710678118654752440f;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int b = 0; b < N; b += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; j++) {
                const int m = j * (16 / h);  // W_{2h}^j = W_32^m
                const float ar = re[b + j], ai = im[b + j];
                const float br = re[b + j + h], bi = im[b + j + h];
                re[b + j] = ar + br;
                im[b + j] = ai + bi;
                const float tr = ar - br, ti = ai - bi;
                if (m == 0) {
                    re[b + j + h] = tr;
                    im[b + j + h] = ti;
                } else if (m == 8) {  // -i
                    re[b + j + h] = ti;
                    im[b + j + h] = -tr;
                } else if (m == 4) {  // (1 - i)/sqrt2
                    re[b + j + h] = (tr + ti) * R;
                    im[b + j + h] = (ti - tr) * R;
                } else if (m == 12) {  // (-1 - i)/sqrt2
                    re[b + j + h] = (ti - tr) * R;
                    im[b + j + h] = -(tr + ti) * R;
                } else {  // (tr + i ti)(c - i s)
                    const float c = c32(m), s = s32(m);
                    re[b + j + h] = fmaf(tr, c, ti * s);
                    im[b + j + h] = fmaf(ti, c, -(tr * s));
                }
            }
        }
    }
}

// Packed form of the same transform: a register pair holds (real, imaginary) of one point, so the butterfly's two complex
// additions are two FADD2 (fp32x2, sm_100) instead of four FADD, and the window multiply of a frame pair is one FMUL2; the
// twiddle products stay scalar on the halves of the pairs.  Same arithmetic, same rounding, fewer issue slots -- the kernel
// is bound by instruction issue, not by the FMA lanes.
typedef float2 f2;
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

template <int N>
__device__ __forceinline__ void fft_dif_ri(f2 *z) {
    constexpr float R = 0.70710678118654752440f;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int b = 0; b < N; b += 2 * h) {
#pragma unroll
            for (int j = 0; j < h; j++) {
                const int m = j * (16 / h);  // W_{2h}^j = W_32^m
                const f2 a = z[b + j], c2 = z[b + j + h];
                z[b + j] = add2(a, c2);
                const f2 t = sub2(a, c2);
                const float tr = t.x, ti = t.y;
                if (m == 0) {
                    z[b + j + h] = t;
                } else if (m == 8) {  // -i
                    z[b + j + h] = make_float2(ti, -tr);
                } else if (m == 4) {  // (1 - i)/sqrt2
                    z[b + j + h] = make_float2((tr + ti) * R, (ti - tr) * R);
                } else if (m == 12) {  // (-1 - i)/sqrt2
                    z[b + j + h] = make_float2((ti - tr) * R, -(tr + ti) * R);
                } else {  // (tr + i ti)(c - i s)
                    const float c = c32(m), s = s32(m);
                    z[b + j + h] = make_float2(fmaf(tr, c, ti * s), fmaf(ti, c, -(tr * s)));
                }
            }
        }
    }
}

__device__ __forceinline__ int64_t reflect_index(int64_t s, int64_t L) {
    if (s < 0) s = -s;
    if (s >= L) s = 2 * (L - 1) - s;
    return s;
}

// A CTA holds MEL_GROUPS independent groups of MEL_WARPS warps.  Each group runs its own clips through its own
// exchange / power / staging buffers and synchronises on its own named barrier, so one group's FFT phase overlaps the
// other's mel projection, write-out and clip epilogue (one 16-warp group spends a third of its cycles at barriers).
#ifndef AT_MEL_GROUPS
// measured on B200 (tools/bench_mel.py, 1024 / 512 / 64): 2 groups x 8 warps 13.19 ms per 20,000 clips, 4 x 4 12.65 ms --
// four phase streams per SM overlap the FMA-heavy FFT phase with the shared-memory-heavy mel / write-out phases better
#define AT_MEL_GROUPS 4
#define AT_MEL_WARPS 4
#define AT_MEL_STAGE 4608   // staged sample window of a batch: (BF - 1) hops + n_fft samples at 1024 / 512 (8 frames)
#define AT_MEL_WT 1536      // filterbank weights kept in shared memory when they fit (1024 / 64 needs ~1,300)
#define AT_MEL_PAR 128
#endif
constexpr int MEL_GROUPS = AT_MEL_GROUPS;
constexpr int MEL_WARPS = AT_MEL_WARPS;                 // per group
constexpr int MEL_THREADS = MEL_WARPS * 32;             // per group
constexpr int CTA_THREADS = MEL_GROUPS * MEL_THREADS;
constexpr int XCH_STRIDE = 33;                          // floats per exchange row (32 + 1 pad)
constexpr int XCH_WARP_F = 32 * XCH_STRIDE;             // floats per warp (real and imaginary parts go through in turn)
constexpr int XCH_BYTES = MEL_WARPS * XCH_WARP_F * 4;   // 33,792 per group
constexpr int STAGE_FLOATS = AT_MEL_STAGE;                      // staged sample window: 15 * 512 + 1024 samples (34,816 B) per group
constexpr int WT_SMEM_FLOATS = AT_MEL_WT;                    // filterbank weights kept in shared memory when they fit
constexpr int PAR_SMEM_MELS = AT_MEL_PAR;                      // filter parameters (first bin, groups of four, weight offset)

__device__ __forceinline__ void group_sync(int grp) {   // named barrier 1 + grp over the group's threads
#ifdef AT_MEL_NOSYNC   // timing experiment only (races, wrong results): what the group barriers cost
    __syncwarp();
#else
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(MEL_THREADS) : "memory");
#endif
}

template <int LOG2NF>
struct MelCfg {
    static constexpr int NF = 1 << LOG2NF;
    static constexpr int N2 = NF / 32;       // in-lane FFT size of pass 1
    static constexpr int LOG2N2 = LOG2NF - 5;
    static constexpr int G = 32 / N2;        // complex FFTs per job
    static constexpr int FR = 2 * G;         // real frames per job
    static constexpr int BF = MEL_WARPS * FR;  // frames per batch
    static constexpr int NB = NF / 2 + 1;    // bins
    static constexpr int PS = (NB + 3) & ~3; // power-tile row stride: rows start 16-byte aligned (vector reads in the mel phase)
    static constexpr int P_FLOATS = BF * PS;
    // per group: [exchange | power tile | staged samples] (the dB tile aliases exchange + power-tile space);
    // then, shared by the groups: [twiddles | window | filter weights | filter parameters]
    static constexpr size_t OFF_P = XCH_BYTES;
    static constexpr size_t OFF_STAGE = (OFF_P + (size_t)P_FLOATS * 4 + 127) & ~(size_t)127;
    static constexpr size_t GROUP_BYTES = OFF_STAGE + (size_t)STAGE_FLOATS * 4;
    static constexpr size_t OFF_TW = MEL_GROUPS * GROUP_BYTES;
    static constexpr size_t OFF_WIN = OFF_TW + (size_t)N2 * 32 * 8;
    static constexpr size_t OFF_WT = OFF_WIN + (size_t)NF * 4;
    static constexpr size_t OFF_PAR = OFF_WT + (size_t)WT_SMEM_FLOATS * 4;
    static constexpr size_t SMEM = OFF_PAR + (size_t)PAR_SMEM_MELS * 3 * 4;
    static_assert(GROUP_BYTES % 128 == 0 && SMEM <= 232448, "shared memory layout");
};

// One unit of CTA work: a batch of BF consecutive frames of one clip.
struct MelItem {
    int clip;
    int64_t t0;     // first frame of the batch
    int64_t s0, L, T, f0;
    bool staged;    // the sample window of this batch comes through the bulk-copy stage
    int64_t lo;     // first staged sample (clip-relative)
    uint32_t bytes; // staged bytes
};

template <int LOG2NF>
__global__ void __launch_bounds__(CTA_THREADS, 1)
k_mel(const float *__restrict__ wave, const int64_t *__restrict__ sample_offsets,
      const int64_t *__restrict__ frame_offsets, int64_t uniform_samples, int B, int hop, int n_mels,
      int normalize, int power_out, const float *__restrict__ g_win, const float2 *__restrict__ g_tw,
      const int *__restrict__ fstart, const int *__restrict__ fcnt, const int *__restrict__ woff,
      const float *__restrict__ wt, int wt_count, float *__restrict__ out, float *__restrict__ out_l2,
      int32_t *__restrict__ bad_flags, unsigned int *__restrict__ absmax_bits) {
    using C = MelCfg<LOG2NF>;
    constexpr int NF = C::NF, N2 = C::N2, G = C::G, FR = C::FR, BF = C::BF, PS = C::PS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int grp = threadIdx.x / MEL_THREADS;            // group of this thread
    const int tid = threadIdx.x % MEL_THREADS, lane = tid & 31, warp = tid >> 5;   // position inside the group
    unsigned char *gsm = smem_raw + (size_t)grp * C::GROUP_BYTES;
    float *xch_all = reinterpret_cast<float *>(gsm);
    float *dtile = reinterpret_cast<float *>(gsm);  // aliases the exchange region (mel stage only)
    float *ptile = reinterpret_cast<float *>(gsm + C::OFF_P);
    float *s_stage = reinterpret_cast<float *>(gsm + C::OFF_STAGE);
    float2 *s_tw = reinterpret_cast<float2 *>(smem_raw + C::OFF_TW);
    float *s_win = reinterpret_cast<float *>(smem_raw + C::OFF_WIN);
    float *s_wt = reinterpret_cast<float *>(smem_raw + C::OFF_WT);
    int *s_par = reinterpret_cast<int *>(smem_raw + C::OFF_PAR);   // [m]{first bin, groups of four, weight offset}
    __shared__ float s_red_all[MEL_GROUPS][2][MEL_WARPS];
    __shared__ int s_flag_all[MEL_GROUPS];
    __shared__ __align__(8) unsigned long long s_bar_all[MEL_GROUPS];
    float (*s_red)[MEL_WARPS] = s_red_all[grp];
    int &s_flag = s_flag_all[grp];

    for (int i = threadIdx.x; i < N2 * 32; i += CTA_THREADS) s_tw[i] = g_tw[i];
    for (int i = threadIdx.x; i < NF; i += CTA_THREADS) s_win[i] = g_win[i];
    const bool wt_in_smem = wt_count <= WT_SMEM_FLOATS && n_mels <= PAR_SMEM_MELS;
    if (wt_in_smem) {
        for (int i = threadIdx.x; i < wt_count; i += CTA_THREADS) s_wt[i] = wt[i];
        for (int i = threadIdx.x; i < n_mels; i += CTA_THREADS)
            s_par[3 * i] = fstart[i], s_par[3 * i + 1] = fcnt[i], s_par[3 * i + 2] = woff[i];
    }
    // the padding columns of the power tile are read (against zero weights) by the vector loads of the mel phase
    for (int i = tid; i < BF * (PS - C::NB); i += MEL_THREADS) ptile[(i / (PS - C::NB)) * PS + C::NB + i % (PS - C::NB)] = 0.f;
    const uint32_t bar = smem_u32(&s_bar_all[grp]);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_groups = (int)gridDim.x * MEL_GROUPS;     // clips are dealt round-robin to the groups of the grid

    float *xch = xch_all + warp * XCH_WARP_F;
    const int dstride = n_mels + 4;  // rows stay 16-byte aligned for the vector write-out
    // pass-2 role of this lane: complex FFT g2, residue k2
    const int g2 = lane / N2, k2 = lane % N2;
    const int partner = g2 * N2 + ((N2 - k2) & (N2 - 1));
    // staging needs 16-byte aligned windows: base pointer, clip offsets, lengths and hop all multiples of 4 samples
    const int64_t win_samples = (int64_t)(BF - 1) * hop + NF;
    const bool can_stage = win_samples <= STAGE_FLOATS && (hop & 3) == 0 && ((uintptr_t)wave & 15) == 0;

    // ---- work-item iterator (every thread evaluates it identically) -------------------------------------------
    auto load_clip = [&](MelItem &it) -> bool {  // fills the clip fields; false when the grid-stride range is exhausted
        while (it.clip < B) {
            it.s0 = sample_offsets ? sample_offsets[it.clip] : (int64_t)it.clip * uniform_samples;
            it.L = sample_offsets ? sample_offsets[it.clip + 1] - it.s0 : uniform_samples;
            it.T = 1 + it.L / hop;
            it.f0 = frame_offsets ? frame_offsets[it.clip] : (int64_t)it.clip * (1 + uniform_samples / hop);
            if (it.L > NF / 2) return true;
            // torch's reflect pad raises when the pad is not smaller than the input: flag 2, no output
            if (tid == 0 && bad_flags) bad_flags[it.clip] = 2;
            it.clip += n_groups;
        }
        return false;
    };
    auto set_window = [&](MelItem &it) {
        const int64_t start = it.t0 * hop - NF / 2;
        it.lo = start < 0 ? 0 : start;
        int64_t hi = start + win_samples;
        if (hi > it.L) hi = it.L;
        it.staged = can_stage && ((it.s0 | it.L) & 3) == 0 && hi > it.lo;
        it.bytes = it.staged ? (uint32_t)((hi - it.lo) * 4) : 0u;
    };
    auto next_item = [&](const MelItem &cur, MelItem &nx) -> bool {
        nx = cur;
        if (cur.t0 + BF < cur.T) {
            nx.t0 = cur.t0 + BF;
        } else {
            nx.clip = cur.clip + n_groups;
            nx.t0 = 0;
            if (!load_clip(nx)) return false;
        }
        set_window(nx);
        return true;
    };
    auto issue_stage = [&](const MelItem &it) {  // one thread: arm the barrier and start the bulk copy
        if (it.staged) {
            mbar_expect_tx(bar, it.bytes);
            bulk_g2s(smem_u32(s_stage), wave + it.s0 + it.lo, it.bytes, bar);
        }
    };

    MelItem cur;
    cur.clip = (int)blockIdx.x * MEL_GROUPS + grp;
    cur.t0 = 0;
    bool have = load_clip(cur);
    if (have) {
        set_window(cur);
        if (tid == 0) issue_stage(cur);
    }
    uint32_t stage_phase = 0;
    float vmin = INFINITY, vmax = -INFINITY;
    int nonfinite = 0;
    float l2max = 0.f;   // largest |element| this thread wrote to the L2-normalised copy (k-means' fixed-point scale)

    while (have) {
        const float *x = wave + cur.s0;
        float *dst = out + cur.f0 * n_mels;
        const int64_t L = cur.L, T = cur.T, t0 = cur.t0;
        MelItem nxt;
        const bool have_next = next_item(cur, nxt);
        if (cur.staged) {
            mbar_wait(bar, stage_phase);
            stage_phase ^= 1;
        }
        // ------------------------------------------------------------ phase A: FFT -> power tile
        // z[p] = (real, imaginary) = (frame a, frame b) samples of point p: packed fp32x2 additions (fft_dif_ri)
        {
            f2 z[32];
            const int64_t tj = t0 + (int64_t)warp * FR;
            // fast path (one frame pair per job, hop = n_fft / 2, both frames interior and staged): the second frame's samples
            // are the first frame's shifted by half a window, so the pair needs 1.5 n_fft staged samples, not 2, and each
            // window value is fetched once for both frames (one FMUL2 per point)
            bool fast_pair = false;
            if (G == 1 && hop == NF / 2 && cur.staged) {
                const int64_t start = tj * hop - NF / 2;
                fast_pair = tj + 1 < T && start >= 0 && start + hop + NF <= L;
                if (fast_pair) {
                    const float *sp = s_stage + (start - cur.lo) + lane;
#pragma unroll
                    for (int n2 = 0; n2 < N2; n2++) {
                        const float w = s_win[lane + 32 * n2];
                        z[n2] = __fmul2_rn(make_float2(sp[32 * n2], sp[32 * (n2 + N2 / 2)]), make_float2(w, w));
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (fast_pair) break;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int64_t f = tj + 2 * g + half;
                    const int64_t start = f * hop - NF / 2;
                    float v[N2];
                    if (f < T && start >= 0 && start + NF <= L) {
                        if (cur.staged) {
                            const float *sp = s_stage + (start - cur.lo) + lane;
#pragma unroll
                            for (int n2 = 0; n2 < N2; n2++) v[n2] = sp[32 * n2] * s_win[lane + 32 * n2];
                        } else {
#pragma unroll
                            for (int n2 = 0; n2 < N2; n2++) v[n2] = __ldg(x + start + lane + 32 * n2) * s_win[lane + 32 * n2];
                        }
                    } else if (f < T) {
#pragma unroll
                        for (int n2 = 0; n2 < N2; n2++)
                            v[n2] = __ldg(x + reflect_index(start + lane + 32 * n2, L)) * s_win[lane + 32 * n2];
                    } else {
#pragma unroll
                        for (int n2 = 0; n2 < N2; n2++) v[n2] = 0.f;
                    }
#pragma unroll
                    for (int n2 = 0; n2 < N2; n2++) {
                        if (half == 0) z[g * N2 + n2].x = v[n2];
                        else z[g * N2 + n2].y = v[n2];
                    }
                }
            }
            // pass 1: N2-point FFTs over n2, then twiddle W_NF^(n1*k2)
#pragma unroll
            for (int g = 0; g < G; g++) fft_dif_ri<N2>(z + g * N2);
#pragma unroll
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int p = 0; p < N2; p++) {
                    const int kk = brev(p, C::LOG2N2);
                    const float2 w = s_tw[kk * 32 + lane];
                    const float a = z[g * N2 + p].x, b = z[g * N2 + p].y;
                    z[g * N2 + p] = make_float2(fmaf(a, w.x, -(b * w.y)), fmaf(a, w.y, b * w.x));
                }
            }
            // 32x32 transpose through the warp's exchange tile, real parts then imaginary parts
#pragma unroll
            for (int part = 0; part < 2; part++) {
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int p = 0; p < N2; p++)
                        xch[(g * N2 + brev(p, C::LOG2N2)) * XCH_STRIDE + lane] = part ? z[g * N2 + p].y : z[g * N2 + p].x;
                __syncwarp();
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    if (part) z[n1].y = xch[lane * XCH_STRIDE + n1];
                    else z[n1].x = xch[lane * XCH_STRIDE + n1];
                }
                __syncwarp();
            }
            // pass 2: 32-point FFT over n1; position p holds k1 = brev5(p); Z[N2*k1 + k2]
            fft_dif_ri<32>(z);
            // separate the two real frames and store |.|^2 (the 1/4 lives in the mel weights)
            float *pa = ptile + (size_t)(warp * FR + 2 * g2) * PS;
            float *pb = pa + PS;
#pragma unroll
            for (int k1 = 0; k1 < 16; k1++) {
                const float Zr = z[brev(k1, 5)].x, Zi = z[brev(k1, 5)].y;
                const float qr = k2 == 0 ? z[brev((32 - k1) & 31, 5)].x : z[brev(31 - k1, 5)].x;
                const float qi = k2 == 0 ? z[brev((32 - k1) & 31, 5)].y : z[brev(31 - k1, 5)].y;
                const float pr = __shfl_sync(0xffffffffu, qr, partner);
                const float pi = __shfl_sync(0xffffffffu, qi, partner);
                const float ar = Zr + pr, ai = Zi - pi, br = Zi + pi, bi = pr - Zr;
                pa[N2 * k1 + k2] = fmaf(ar, ar, ai * ai);
                pb[N2 * k1 + k2] = fmaf(br, br, bi * bi);
            }
            if (k2 == 0) {  // Nyquist bin: Z[NF/2] is its own partner
                const float Zr = z[brev(16, 5)].x, Zi = z[brev(16, 5)].y;
                pa[NF / 2] = 4.f * Zr * Zr;
                pb[NF / 2] = 4.f * Zi * Zi;
            }
        }
        group_sync(grp);
        // the staged window has been consumed by every warp: start fetching the next batch's window now, it lands
        // while this batch goes through the mel projection and the write-out
        if (tid == 0 && have_next) issue_stage(nxt);
        // ------------------------------------------------------------ phase C: sparse mel + dB
        {
            // a warp item = FPW frames x MPW filters (lanes <-> frames first: power-tile rows are 513 floats apart,
            // conflict-free; with 16-frame batches the two half-warps take two adjacent filters)
            constexpr int FPW = BF < 32 ? BF : 32, MPW = 32 / FPW, FG = BF / FPW;
            const int mgroups = (n_mels + MPW - 1) / MPW;
            for (int item = warp; item < FG * mgroups; item += MEL_WARPS) {
                const int fg = item % FG, m = (item / FG) * MPW + lane / FPW;
                const int f = fg * FPW + lane % FPW;
                if (m >= n_mels) continue;
                const float *prow = ptile + (size_t)f * PS + (wt_in_smem ? s_par[3 * m] : fstart[m]);   // first bin: a multiple of 4
                const int cnt4 = wt_in_smem ? s_par[3 * m + 1] : fcnt[m];
                float acc = 0.f;
                if (wt_in_smem) {
                    const float4 *w4 = reinterpret_cast<const float4 *>(s_wt + s_par[3 * m + 2]);
                    // two independent chains (even / odd segments) and four segments in flight: one chain of dependent FMAs
                    // behind 128-bit shared loads left this phase waiting on the loads (24 % of the warp time, ncu)
                    float acc1 = 0.f;
                    int i = 0;
#pragma unroll 2
                    for (; i + 1 < cnt4; i += 2) {
                        const float4 w = w4[i], w2 = w4[i + 1];  // broadcast: the lanes of a filter read the same segment
                        const float4 pw = *reinterpret_cast<const float4 *>(prow + 4 * i);   // conflict-free: rows are 516 floats apart
                        const float4 pw2 = *reinterpret_cast<const float4 *>(prow + 4 * i + 4);
                        acc = fmaf(w.x, pw.x, acc), acc1 = fmaf(w2.x, pw2.x, acc1);
                        acc = fmaf(w.y, pw.y, acc), acc1 = fmaf(w2.y, pw2.y, acc1);
                        acc = fmaf(w.z, pw.z, acc), acc1 = fmaf(w2.z, pw2.z, acc1);
                        acc = fmaf(w.w, pw.w, acc), acc1 = fmaf(w2.w, pw2.w, acc1);
                    }
                    if (i < cnt4) {
                        const float4 w = w4[i];
                        const float4 pw = *reinterpret_cast<const float4 *>(prow + 4 * i);
                        acc = fmaf(w.x, pw.x, acc), acc = fmaf(w.y, pw.y, acc), acc = fmaf(w.z, pw.z, acc), acc = fmaf(w.w, pw.w, acc);
                    }
                    acc += acc1;
                } else {
                    const float *w = wt + woff[m];
                    for (int i = 0; i < 4 * cnt4; i++) acc = fmaf(__ldg(w + i), prow[i], acc);
                }
                // 10 log10(x) = (10 log10 2) log2(x): MUFU.LG2 (relative error 2^-22) instead of the ~25-instruction log10f;
                // the result differs from torch's by < 3e-5 dB over the whole range (gate: 1e-4 of the clip's range)
                dtile[f * dstride + m] = power_out ? acc : 3.01029995663981195f * __log2f(fmaxf(acc, 1e-10f));
            }
        }
        group_sync(grp);
        // ------------------------------------------------------------ phase D: coalesced write-out
        {
            const int nf = (int)min((int64_t)BF, T - t0);
            float *o = dst + t0 * n_mels;
            if ((n_mels & 3) == 0) {
                const int q4 = n_mels >> 2;  // float4 per frame; a half-warp takes whole frames (no integer division)
                for (int f = warp * 2 + (lane >> 4); f < nf; f += MEL_WARPS * 2) {
                    for (int j = lane & 15; j < q4; j += 16) {
                        const float4 v = *reinterpret_cast<const float4 *>(dtile + f * dstride + 4 * j);
                        vmin = fminf(fminf(vmin, v.x), fminf(fminf(v.y, v.z), v.w));
                        vmax = fmaxf(fmaxf(vmax, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
                        nonfinite |= !(isfinite(v.x) && isfinite(v.y) && isfinite(v.z) && isfinite(v.w));
                        *reinterpret_cast<float4 *>(o + (size_t)f * n_mels + 4 * j) = v;
                    }
                }
            } else {
                for (int f = warp; f < nf; f += MEL_WARPS) {
                    for (int m = lane; m < n_mels; m += 32) {
                        const float v = dtile[f * dstride + m];
                        vmin = fminf(vmin, v);
                        vmax = fmaxf(vmax, v);
                        nonfinite |= !isfinite(v);
                        o[(size_t)f * n_mels + m] = v;
                    }
                }
            }
        }
        group_sync(grp);

        if (t0 + BF >= T) {
            // ---------------------------------------------------------------- clip epilogue
            vmin = warp_min(vmin);
            vmax = warp_max(vmax);
            if (lane == 0) s_red[0][warp] = vmin, s_red[1][warp] = vmax;
            if (tid == 0) s_flag = 0;
            group_sync(grp);
            float mn = s_red[0][0], mx = s_red[1][0];
#pragma unroll
            for (int w = 1; w < MEL_WARPS; w++) mn = fminf(mn, s_red[0][w]), mx = fmaxf(mx, s_red[1][w]);
            if (normalize || out_l2) {
                // half-warp per frame row; chunk class g handles elements 4g..4g+3 (+64, +128, ...)
                const float range = __fsub_rn(mx, mn);
                const float inv_range = __fdiv_rn(1.0f, range);  // range == 0 -> inf -> 0 * inf = NaN like the reference's 0/0
                const int g = tid & 15, hw = (tid >> 4) & 1;
                const bool vec = (n_mels & 3) == 0 && n_mels <= 256;
                constexpr int RSTEP = MEL_THREADS / 16;   // rows per pass of the group
                if (vec && n_mels <= 64) {
                    // one float4 per lane (the 64-mel configuration of the reference); four rows in flight per half-warp: the tile
                    // comes back from L2 and a single dependent load per pass left this loop waiting (13 % of the warp time, ncu)
                    constexpr int EU = 4;
                    const bool col_on = 4 * g < n_mels;
                    // warp-uniform trip count (both half-warps iterate together: the shuffles below need all 32 lanes)
                    for (int64_t r0 = (int64_t)(tid >> 5) * 2; r0 < T; r0 += RSTEP * EU) {
                        float4 v[EU];
                        bool on[EU];
#pragma unroll
                        for (int u = 0; u < EU; u++) {
                            const int64_t r = r0 + u * RSTEP + hw;
                            on[u] = r < T && col_on;
                            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (on[u]) v[u] = __ldcg(reinterpret_cast<const float4 *>(dst + r * n_mels + 4 * g));
                        }
#pragma unroll
                        for (int u = 0; u < EU; u++) {
                            const int64_t r = r0 + u * RSTEP + hw;
                            float4 w = v[u];
                            if (normalize) {
                                w.x = __fmul_rn(__fsub_rn(w.x, mn), inv_range), w.y = __fmul_rn(__fsub_rn(w.y, mn), inv_range);
                                w.z = __fmul_rn(__fsub_rn(w.z, mn), inv_range), w.w = __fmul_rn(__fsub_rn(w.w, mn), inv_range);
                                if (on[u]) {
                                    nonfinite |= !(isfinite(w.x) && isfinite(w.y) && isfinite(w.z) && isfinite(w.w));
                                    *reinterpret_cast<float4 *>(dst + r * n_mels + 4 * g) = w;
                                } else {
                                    w = make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
                            if (out_l2) {
                                float q = w.x * w.x;
                                q = fmaf(w.y, w.y, q), q = fmaf(w.z, w.z, q), q = fmaf(w.w, w.w, q);
                                const float den = l2_denominator(half16_sum(q));
                                if (on[u]) {
                                    const float4 o4 = make_float4(__fdiv_rn(w.x, den), __fdiv_rn(w.y, den), __fdiv_rn(w.z, den),
                                                                  __fdiv_rn(w.w, den));
                                    l2max = fmaxf(fmaxf(l2max, fmaxf(fabsf(o4.x), fabsf(o4.y))), fmaxf(fabsf(o4.z), fabsf(o4.w)));
                                    *reinterpret_cast<float4 *>(out_l2 + (cur.f0 + r) * n_mels + 4 * g) = o4;
                                }
                            }
                        }
                    }
                } else
                // warp-uniform trip count (both half-warps iterate together: the shuffles below need all 32 lanes)
                for (int64_t r0 = (int64_t)(tid >> 5) * 2; r0 < T; r0 += RSTEP) {
                    const int64_t r = r0 + hw;
                    const bool live = r < T;
                    float *row = dst + (live ? r : 0) * n_mels;
                    float q = 0.f;
                    if (vec) {
                        float4 v[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int base = 4 * g + 64 * j;
                            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (live && base < n_mels) v[j] = __ldcg(reinterpret_cast<const float4 *>(row + base));
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int base = 4 * g + 64 * j;
                            if (live && base < n_mels) {
                                if (normalize) {
                                    v[j].x = __fmul_rn(__fsub_rn(v[j].x, mn), inv_range);
                                    v[j].y = __fmul_rn(__fsub_rn(v[j].y, mn), inv_range);
                                    v[j].z = __fmul_rn(__fsub_rn(v[j].z, mn), inv_range);
                                    v[j].w = __fmul_rn(__fsub_rn(v[j].w, mn), inv_range);
                                    nonfinite |= !(isfinite(v[j].x) && isfinite(v[j].y) && isfinite(v[j].z) && isfinite(v[j].w));
                                    *reinterpret_cast<float4 *>(row + base) = v[j];
                                }
                                q = fmaf(v[j].x, v[j].x, q), q = fmaf(v[j].y, v[j].y, q);
                                q = fmaf(v[j].z, v[j].z, q), q = fmaf(v[j].w, v[j].w, q);
                            }
                        }
                        if (out_l2) {
                            const float den = l2_denominator(half16_sum(q));
                            float *orow = out_l2 + (cur.f0 + (live ? r : 0)) * n_mels;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int base = 4 * g + 64 * j;
                                if (live && base < n_mels) {
                                    const float4 o4 = make_float4(__fdiv_rn(v[j].x, den), __fdiv_rn(v[j].y, den),
                                                                  __fdiv_rn(v[j].z, den), __fdiv_rn(v[j].w, den));
                                    l2max = fmaxf(fmaxf(l2max, fmaxf(fabsf(o4.x), fabsf(o4.y))), fmaxf(fabsf(o4.z), fabsf(o4.w)));
                                    *reinterpret_cast<float4 *>(orow + base) = o4;
                                }
                            }
                        }
                        continue;
                    }
                    if (live) {
                        for (int base = 4 * g; base < n_mels; base += 64) {
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                if (base + e < n_mels) {
                                    float v = __ldcg(row + base + e);
                                    if (normalize) {
                                        v = __fmul_rn(__fsub_rn(v, mn), inv_range);
                                        nonfinite |= !isfinite(v);
                                        row[base + e] = v;
                                    }
                                    q = fmaf(v, v, q);
                                }
                            }
                        }
                    }
                    if (out_l2) {
                        const float den = l2_denominator(half16_sum(q));
                        if (live) {
                            float *orow = out_l2 + (cur.f0 + r) * n_mels;
                            for (int base = 4 * g; base < n_mels; base += 64) {
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    // final value of the row (this thread wrote it just above when normalising)
                                    if (base + e < n_mels) {
                                        const float o1 = __fdiv_rn(__ldcg(row + base + e), den);
                                        l2max = fmaxf(l2max, fabsf(o1));
                                        orow[base + e] = o1;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (nonfinite) s_flag = 1;
            if (absmax_bits && out_l2) {   // non-negative floats order like their bit patterns; NaNs never raise fmaxf
                l2max = warp_max(l2max);
                if (lane == 0 && l2max > 0.f) atomicMax(absmax_bits, __float_as_uint(l2max));
                l2max = 0.f;
            }
            group_sync(grp);
            if (tid == 0 && bad_flags) bad_flags[cur.clip] = s_flag;
            group_sync(grp);
            vmin = INFINITY, vmax = -INFINITY, nonfinite = 0;
        }
        cur = nxt;
        have = have_next;
    }
}

static void build_filterbank_htk(int sample_rate, int n_fft, int n_mels, std::vector<float> &fb) {
    // torchaudio.functional.melscale_fbanks(n_freqs, 0, sr/2, n_mels, sr, norm=None, mel_scale="htk")
    // (functional.py:518-587), evaluated in double and rounded once.
    const int nb = n_fft / 2 + 1;
    const double f_max = (double)(sample_rate / 2);  // MelScale default f_max = float(sample_rate // 2)
    const double m_min = 2595.0 * log10(1.0 + 0.0 / 700.0);
    const double m_max = 2595.0 * log10(1.0 + f_max / 700.0);
    std::vector<double> f_pts(n_mels + 2);
    for (int i = 0; i < n_mels + 2; i++) {
        double m = m_min + (m_max - m_min) * (double)i / (double)(n_mels + 1);
        f_pts[i] = 700.0 * (pow(10.0, m / 2595.0) - 1.0);
    }
    fb.assign((size_t)nb * n_mels, 0.f);
    for (int k = 0; k < nb; k++) {
        double freq = (double)(sample_rate / 2) * (double)k / (double)(nb - 1);
        for (int m = 0; m < n_mels; m++) {
            double down = (freq - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
            double up = (f_pts[m + 2] - freq) / (f_pts[m + 2] - f_pts[m + 1]);
            double v = down < up ? down : up;
            fb[(size_t)k * n_mels + m] = v > 0.0 ? (float)v : 0.f;
        }
    }
}

static int upload_constants(at_mel_plan *p) {
    const int nf = p->n_fft, nb = nf / 2 + 1, nm = p->n_mels, n2 = nf / 32;
    // CSR by filter of the dense (nb, nm) filterbank; support = [first non-zero, last non-zero]
    std::vector<int> fstart(nm), fcnt(nm), woff(nm);
    std::vector<float> wt;
    for (int m = 0; m < nm; m++) {
        int lo = nb, hi = -1;
        for (int k = 0; k < nb; k++)
            if (p->h_fb[(size_t)k * nm + m] != 0.f) {
                if (k < lo) lo = k;
                hi = k;
            }
        fstart[m] = hi < 0 ? 0 : (lo & ~3);   // segments start on a multiple of 4 bins (16-byte aligned power-tile reads)
        fcnt[m] = hi < 0 ? 0 : hi - fstart[m] + 1;
        woff[m] = (int)wt.size();
        for (int k = fstart[m]; k < fstart[m] + fcnt[m]; k++) wt.push_back(0.25f * p->h_fb[(size_t)k * nm + m]);
        while (wt.size() % 4) wt.push_back(0.f);
        fcnt[m] = (fcnt[m] + 3) / 4;  // in groups of four bins
    }
    if (wt.empty()) wt.assign(4, 0.f);
    p->wt_count = (int)wt.size();
    std::vector<float2> tw((size_t)n2 * 32);
    for (int kk = 0; kk < n2; kk++)
        for (int n1 = 0; n1 < 32; n1++) {
            double a = -2.0 * M_PI * (double)((n1 * kk) % nf) / (double)nf;
            tw[(size_t)kk * 32 + n1] = make_float2((float)cos(a), (float)sin(a));
        }
    cudaFree(p->wt);
    p->wt = nullptr;
    AT_CUDA_OK(cudaMalloc(&p->wt, sizeof(float) * wt.size()));
    if (!p->win) {
        AT_CUDA_OK(cudaMalloc(&p->win, sizeof(float) * nf));
        AT_CUDA_OK(cudaMalloc(&p->tw, sizeof(float2) * tw.size()));
        AT_CUDA_OK(cudaMalloc(&p->fstart, sizeof(int) * nm));
        AT_CUDA_OK(cudaMalloc(&p->fcnt, sizeof(int) * nm));
        AT_CUDA_OK(cudaMalloc(&p->woff, sizeof(int) * nm));
    }
    AT_CUDA_OK(cudaMemcpy(p->win, p->h_win.data(), sizeof(float) * nf, cudaMemcpyHostToDevice));
    AT_CUDA_OK(cudaMemcpy(p->tw, tw.data(), sizeof(float2) * tw.size(), cudaMemcpyHostToDevice));
    AT_CUDA_OK(cudaMemcpy(p->fstart, fstart.data(), sizeof(int) * nm, cudaMemcpyHostToDevice));
    AT_CUDA_OK(cudaMemcpy(p->fcnt, fcnt.data(), sizeof(int) * nm, cudaMemcpyHostToDevice));
    AT_CUDA_OK(cudaMemcpy(p->woff, woff.data(), sizeof(int) * nm, cudaMemcpyHostToDevice));
    AT_CUDA_OK(cudaMemcpy(p->wt, wt.data(), sizeof(float) * wt.size(), cudaMemcpyHostToDevice));
    return AT_OK;
}

template <int LOG2NF>
static int launch_mel(at_mel_plan *p, const float *wave, const int64_t *so, const int64_t *fo, int64_t us, int B,
                      float *out, float *out_l2, int32_t *bad, cudaStream_t st) {
    using C = MelCfg<LOG2NF>;
    static bool configured[MAX_DEVICES] = {};   // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        AT_CUDA_OK(cudaFuncSetAttribute(k_mel<LOG2NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured[dev] = true;
    }
    int grid = sm_count();
    if (grid > (B + MEL_GROUPS - 1) / MEL_GROUPS) grid = (B + MEL_GROUPS - 1) / MEL_GROUPS;
    if (grid < 1) grid = 1;
    ProfScope prof(PROF_MEL, st);
    k_mel<LOG2NF><<<grid, CTA_THREADS, C::SMEM, st>>>(wave, so, fo, us, B, p->hop, p->n_mels, p->normalize, p->power_out, p->win,
                                                     p->tw, p->fstart, p->fcnt, p->woff, p->wt, p->wt_count, out, out_l2,
                                                     bad, reinterpret_cast<unsigned int *>(p->absmax_out));
    AT_LAUNCH_OK();
    return AT_OK;
}

}  // namespace at

using namespace at;

extern "C" {

int at_mel_plan_create(int sample_rate, int n_fft, int hop_length, int n_mels, int normalize, at_mel_plan **plan) {
    AT_REQUIRE(plan, "at_mel_plan_create: null plan pointer");
    AT_REQUIRE(sample_rate > 0 && hop_length > 0 && n_mels > 0, "at_mel_plan_create: bad arguments");
    if (!(n_fft == 256 || n_fft == 512 || n_fft == 1024)) {
        set_error("at_mel_plan_create: n_fft=%d is not covered (256, 512, 1024 are)", n_fft);
        return AT_ERR_UNSUPPORTED;
    }
    // the dB tile of one batch (8 * 2048 / n_fft frames x (n_mels + 4) floats) must fit a group's 33,792-byte exchange area
    const int max_mels = 33792 / (4 * (8 * 2048 / n_fft)) - 4 < 256 ? 33792 / (4 * (8 * 2048 / n_fft)) - 4 : 256;
    if (n_mels > max_mels) {
        set_error("at_mel_plan_create: n_mels=%d > %d is not covered for n_fft=%d", n_mels, max_mels, n_fft);
        return AT_ERR_UNSUPPORTED;
    }
    int dev;
    AT_CUDA_OK(cudaGetDevice(&dev));
    at_mel_plan *p = new (std::nothrow) at_mel_plan();
    if (!p) return AT_ERR_NOMEM;
    p->sample_rate = sample_rate, p->n_fft = n_fft, p->hop = hop_length, p->n_mels = n_mels, p->normalize = normalize;
    p->log2nf = ilog2(n_fft);
    p->h_win.resize(n_fft);
    for (int n = 0; n < n_fft; n++) p->h_win[n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)n_fft));
    build_filterbank_htk(sample_rate, n_fft, n_mels, p->h_fb);
    int rc = upload_constants(p);
    if (rc != AT_OK) {
        at_mel_plan_destroy(p);
        return rc;
    }
    *plan = p;
    return AT_OK;
}

int at_mel_plan_set_constants_host(at_mel_plan *p, const float *window, const float *fb) {
    AT_REQUIRE(p, "at_mel_plan_set_constants_host: null plan");
    if (window) p->h_win.assign(window, window + p->n_fft);
    if (fb) p->h_fb.assign(fb, fb + (size_t)(p->n_fft / 2 + 1) * p->n_mels);
    AT_CUDA_OK(cudaDeviceSynchronize());
    return upload_constants(p);
}

int at_mel_plan_set_output(at_mel_plan *p, int kind) {
    AT_REQUIRE(p && (kind == AT_MEL_OUT_DB || kind == AT_MEL_OUT_POWER), "at_mel_plan_set_output: bad arguments");
    AT_REQUIRE(!(kind == AT_MEL_OUT_POWER && p->normalize), "at_mel_plan_set_output: the power output is not min-max normalised");
    p->power_out = kind == AT_MEL_OUT_POWER;
    return AT_OK;
}

int at_mel_plan_destroy(at_mel_plan *p) {
    if (!p) return AT_OK;
    cudaFree(p->win), cudaFree(p->tw), cudaFree(p->fstart), cudaFree(p->fcnt), cudaFree(p->woff), cudaFree(p->wt);
    for (int i = 0; i < 2; i++) {
        cudaFree(p->stage_in[i]), cudaFree(p->stage_out[i]), cudaFree(p->stage_bad[i]);
        if (p->streams[i]) cudaStreamDestroy(p->streams[i]);
    }
    delete p;
    return AT_OK;
}

int at_mel_plan_set_absmax_out(at_mel_plan *p, float *absmax_dev) {
    AT_REQUIRE(p, "at_mel_plan_set_absmax_out: null plan");
    p->absmax_out = absmax_dev;
    return AT_OK;
}

int at_mel_work_groups(const at_mel_plan *p) {
    (void)p;
    const int sms = sm_count();
    return (sms > 0 ? sms : 1) * MEL_GROUPS;
}

int64_t at_mel_num_frames(const at_mel_plan *p, int64_t n_samples) {
    if (!p || n_samples < 0) return -1;
    return 1 + n_samples / p->hop;
}

int at_mel_forward(at_mel_plan *p, const float *wave, const int64_t *sample_offsets, const int64_t *frame_offsets,
                   int64_t uniform_samples, int B, float *out, float *out_l2, int32_t *bad_flags, void *stream) {
    AT_REQUIRE(p && wave && out && B >= 0, "at_mel_forward: bad arguments");
    AT_REQUIRE((sample_offsets == nullptr) == (frame_offsets == nullptr),
               "at_mel_forward: pass both offset arrays or neither");
    AT_REQUIRE(sample_offsets || uniform_samples > 0, "at_mel_forward: uniform_samples must be > 0 without offsets");
    if (B == 0) return AT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (p->log2nf) {
        case 8: return launch_mel<8>(p, wave, sample_offsets, frame_offsets, uniform_samples, B, out, out_l2, bad_flags, st);
        case 9: return launch_mel<9>(p, wave, sample_offsets, frame_offsets, uniform_samples, B, out, out_l2, bad_flags, st);
        case 10: return launch_mel<10>(p, wave, sample_offsets, frame_offsets, uniform_samples, B, out, out_l2, bad_flags, st);
    }
    set_error("at_mel_forward: unsupported n_fft");
    return AT_ERR_UNSUPPORTED;
}

int at_mel_forward_host(at_mel_plan *p, const float *wave, int64_t uniform_samples, int B, float *out,
                        int32_t *bad_flags) {
    AT_REQUIRE(p && wave && out && B >= 0 && uniform_samples > 0, "at_mel_forward_host: bad arguments");
    if (B == 0) return AT_OK;
    const int64_t T = 1 + uniform_samples / p->hop;
    // chunk = a few waves of the persistent grid, bounded to ~256 MB of samples per buffer
    int64_t chunk = (int64_t)sm_count() * 2;
    const int64_t max_chunk = (256LL << 20) / (uniform_samples * 4) > 0 ? (256LL << 20) / (uniform_samples * 4) : 1;
    if (chunk > max_chunk) chunk = max_chunk;
    if (chunk > B) chunk = B;
    if (p->stage_clips < chunk || p->stage_samples != uniform_samples) {
        AT_CUDA_OK(cudaDeviceSynchronize());
        for (int i = 0; i < 2; i++) {
            cudaFree(p->stage_in[i]), cudaFree(p->stage_out[i]), cudaFree(p->stage_bad[i]);
            p->stage_in[i] = p->stage_out[i] = nullptr, p->stage_bad[i] = nullptr;
            AT_CUDA_OK(cudaMalloc(&p->stage_in[i], sizeof(float) * (size_t)(chunk * uniform_samples)));
            AT_CUDA_OK(cudaMalloc(&p->stage_out[i], sizeof(float) * (size_t)(chunk * T * p->n_mels)));
            AT_CUDA_OK(cudaMalloc(&p->stage_bad[i], sizeof(int32_t) * (size_t)chunk));
            if (!p->streams[i]) AT_CUDA_OK(cudaStreamCreateWithFlags(&p->streams[i], cudaStreamNonBlocking));
        }
        p->stage_clips = chunk, p->stage_samples = uniform_samples;
    }
    int buf = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, buf ^= 1) {
        const int64_t nb = B - b0 < chunk ? B - b0 : chunk;
        cudaStream_t st = p->streams[buf];  // stream order serialises reuse of this buffer
        AT_CUDA_OK(cudaMemcpyAsync(p->stage_in[buf], wave + b0 * uniform_samples,
                                   sizeof(float) * (size_t)(nb * uniform_samples), cudaMemcpyHostToDevice, st));
        AT_CUDA_OK(cudaMemsetAsync(p->stage_bad[buf], 0, sizeof(int32_t) * (size_t)nb, st));
        int rc = at_mel_forward(p, p->stage_in[buf], nullptr, nullptr, uniform_samples, (int)nb, p->stage_out[buf],
                                nullptr, p->stage_bad[buf], st);
        if (rc != AT_OK) return rc;
        AT_CUDA_OK(cudaMemcpyAsync(out + b0 * T * p->n_mels, p->stage_out[buf],
                                   sizeof(float) * (size_t)(nb * T * p->n_mels), cudaMemcpyDeviceToHost, st));
        if (bad_flags)
            AT_CUDA_OK(cudaMemcpyAsync(bad_flags + b0, p->stage_bad[buf], sizeof(int32_t) * (size_t)nb,
                                       cudaMemcpyDeviceToHost, st));
    }
    AT_CUDA_OK(cudaStreamSynchronize(p->streams[0]));
    AT_CUDA_OK(cudaStreamSynchronize(p->streams[1]));
    return AT_OK;
}

}  // extern "C"
