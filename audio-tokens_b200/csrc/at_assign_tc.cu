// libat_b200: nearest-centroid search on the 5th-generation tensor cores (tcgen05 + TMEM), d == 64.
//
// Replaces the blocked sgemm + top-1 scan of faiss::exhaustive_L2sqr_blas (reached from
// processors/spec_tokenizer.py:77 and, inside faiss.Kmeans.train, processors/cluster_creator.py:54-56).
//
// Arithmetic.  The tensor core evaluates, for a tile of 128 rows x 128 centroids,
//     acc = S^2 * ( |x|^2 + |c|^2 - 2 <x, c> )
// in ONE chain of 13 tcgen05.mma (kind::f16, fp32 accumulate in TMEM):
//     x_hi * c_hi  +  x_hi * c_lo  +  x_lo * c_hi          (fp16 hi/lo split of S*x and -2*S*c: ~22 mantissa bits)
//   + [xn pieces | 2^12 2^12 2^12] * [2^12 2^12 2^12 | cn pieces]   (one extra K=16 step carrying both norms)
// so the epilogue only scans packed (distance | column) keys.  The ALU pipe (FMNMX, LOP3: one warp instruction per two
// cycles per scheduler) bounds that scan, so it keeps the top-2 over minima of groups of four adjacent columns
// (1.8 ALU ops per score) instead of an exact per-column top-2 (3.5, measured ALU-bound at 2.1x the MMA time).  The
// two candidates are re-evaluated by separate warps with the library's canonical fp32 formula (at_index.cuh) -- the
// one the exact SIMT kernel uses -- and the smaller wins (lowest index on exact ties).
// A row whose runner-up is safely behind the winner (gap above a bound on the split-fp16 error) skips the re-check: its
// label is final and its distance is the accumulator value / S^2 (agrees with the canonical fp32 value to ~2e-5
// relative, 1e-6 absolute for unit-norm data).  All other rows get the exact canonical distance.
// Tie note: token ids agree with the fp32 path except when (a) three or more centroids lie within ~1e-7 (absolute,
// unit-norm data) of the minimum, or (b) the runner-up shares the winner's group of four AND lies within ~1e-7 of it
// (the group hides it from the re-check); both are ties below what the fp32 formula itself resolves.
//
// Roofline note: 2*N*K*64 algorithmic flops are executed as 3.25x that many fp16 MMA flops.
//
// Two modes share one kernel:
//   RESIDENT  (K <= 512 per CTA): the CTA's centroid operand tiles (36 KB each, pre-swizzled by k_tc_prep) are loaded
//             into shared memory once and stay there; for K <= 2048 the centroids are cut into S = ceil(tiles/4)
//             slices, CTA b serves slice b % S, and a small merge kernel takes the minimum over slices.  No operand
//             re-streaming from L2 (at K = 1024 the streaming mode moved 22 GB per launch through L2).
//   STREAM    (any K): operand tiles are streamed through a 4-slot ring with cp.async.bulk.
//
// Structure: persistent CTAs (one per SM), 16 warps:
//   warp 1  lane 0   issues tcgen05.mma, commits to mbarriers
//   warp 2           TMEM allocation / deallocation
//   warp 3  lane 0   bulk-copies (cp.async.bulk, TMA engine) centroid operand tiles into shared memory
//   warps 4-7        read the fp32 row tile (coalesced 128-bit loads), optional row L2 normalisation, |x|^2, fp16 hi/lo
//                    split written straight into the SWIZZLE_128B K-major layout the MMA descriptors expect
//   warps 8-11       epilogue: tcgen05.ld the accumulator, packed keys, group minima, top-2 -> two candidate columns
//   warps 12-15      fp32 re-check of the candidates (off the MMA critical path), labels / distances out
#include "at_index.cuh"
#include "at_ptx.cuh"

namespace at {

constexpr int TC_THREADS = 512;
constexpr int TM = 128;          // rows per tile (UMMA M)
constexpr int TN = 128;          // centroids per tile (UMMA N)
constexpr int B_SLOTS = 4;       // operand tiles resident per CTA / ring depth
constexpr uint32_t A_MAIN_BYTES = TM * 128;              // 128 rows x 64 fp16
constexpr uint32_t AUG_BYTES = TM * 32;                  // 128 rows x 16 fp16, no-swizzle core matrices
constexpr uint32_t A_BUF_BYTES = 2 * A_MAIN_BYTES + AUG_BYTES;   // hi | lo | aug = 36,864
constexpr uint32_t B_TILE_BYTES = 2 * TN * 128 + TN * 32;        // hi | lo | aug = 36,864

// shared memory map (dynamic, 1024-B aligned base)
constexpr uint32_t OFF_A = 0;                                    // 2 buffers
constexpr uint32_t OFF_B = OFF_A + 2 * A_BUF_BYTES;              // 4 slots
constexpr uint32_t OFF_BAR = OFF_B + B_SLOTS * B_TILE_BYTES;     // mbarriers
constexpr uint32_t OFF_FLAGS = OFF_BAR + 256;                    // row fallback flags 2 x 128 bytes
constexpr uint32_t OFF_CAND = OFF_FLAGS + 256;                   // candidate pairs 2 x 128 x int2
constexpr uint32_t TC_SMEM = OFF_CAND + 2048 + 1024;             // + slack for manual 1024-B alignment
static_assert(OFF_A % 1024 == 0 && OFF_B % 1024 == 0 && A_BUF_BYTES % 1024 == 0 && B_TILE_BYTES % 1024 == 0, "align");
static_assert(TC_SMEM <= 232448, "shared memory budget");

enum {
    BAR_A_FULL = 0,    // +2
    BAR_A_EMPTY = 2,   // +2
    BAR_B_FULL = 4,    // +4
    BAR_B_EMPTY = 8,   // +4
    BAR_ACC_FULL = 12,   // +2
    BAR_ACC_EMPTY = 14,  // +2
    BAR_CAND_FULL = 16,  // +2
    BAR_CAND_EMPTY = 18, // +2
    BAR_COUNT = 20
};

constexpr float AUG_ONE = 4096.0f;            // 2^12, exact in fp16
constexpr float AUG_INV = 1.0f / 4096.0f;
constexpr float PAD_NORM = 30000.0f;          // aug entry of padding columns: acc ~ 1.2e8, never selected
constexpr float ROW_LIMIT = 1024.0f;          // |S*x| above this -> exact fallback for the row

// (mbarrier / bulk-copy wrappers: at_ptx.cuh)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major operand descriptors (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 |
// version 1 <<46 | layout <<61
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {  // rows of 128 B, 8-row swizzle atoms 1024 B apart
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr) {  // 8x16B core matrices: K-adjacent 128 B apart, row groups 256 B
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) |
           ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: D fp32, A/B fp16, both K-major, N at [17,23) >> 3, M at [24,29) >> 4
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// value split into three fp16 pieces (p1 + p2 + p3 ~ v to ~33 bits)
__device__ __forceinline__ void split3(float v, __half &p1, __half &p2, __half &p3) {
    p1 = __float2half_rn(v);
    float r = v - __half2float(p1);
    p2 = __float2half_rn(r);
    r -= __half2float(p2);
    p3 = __float2half_rn(r);
}

// byte offset of 16-byte chunk `chunk` (0..7) of row r in a SWIZZLE_128B K-major tile
__device__ __host__ __forceinline__ uint32_t sw128_off(int r, int chunk) { return (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4); }
// byte offset of 16-byte K-chunk kc (0..1) of row r in the no-swizzle aug tile
__device__ __host__ __forceinline__ uint32_t aug_off(int r, int kc) { return (uint32_t)(r >> 3) * 256u + (uint32_t)kc * 128u + (uint32_t)(r & 7) * 16u; }

// ------------------------------------------------------------------------------------------ operand prep
__global__ void k_tc_scale(const float *__restrict__ c, int n, float *__restrict__ scale) {
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(c[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, red[w]);
        int e = 0;
        if (m > 0.f && isfinite(m)) {
            int ex;
            frexpf(m, &ex);  // m < 2^ex
            e = 7 - ex;      // m * 2^e in [64, 128)
        }
        float S = ldexpf(1.0f, e);
        scale[0] = S;
        scale[1] = S * S * AUG_INV;
        scale[2] = 1.0f / (S * S);
        // absolute part of the "is the runner-up safely behind?" threshold, in accumulator units: 2^-19 * S^2 * 64 m^2
        // bounds 2^-19 S^2 |c|^2 (m = max |c_ij|); split-fp16 products carry ~2^-21 of |x||c| S^2, fp32 accumulation
        // a few 2^-24 of the same, so this leaves a factor ~4 of head-room
        scale[3] = ldexpf(S * S * 64.0f * m * m, -19);
    }
}

// one thread per (padded centroid, 16-byte chunk): chunks 0..7 main columns, chunk 8 = aug
__global__ void k_tc_prep(const float *__restrict__ c, const float *__restrict__ cn, int k, int ktiles,
                          const float *__restrict__ scale, unsigned char *__restrict__ op) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int j = idx / 9, chunk = idx % 9;
    if (j >= ktiles * TN) return;
    const float S = scale[0], S2A = scale[1];
    unsigned char *tile = op + (size_t)(j / TN) * B_TILE_BYTES;
    const int r = j % TN;
    if (chunk < 8) {
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            float v = j < k ? -2.0f * S * c[(size_t)j * 64 + chunk * 8 + e] : 0.f;
            hi[e] = __float2half_rn(v);
            lo[e] = __float2half_rn(v - __half2float(hi[e]));
        }
        *reinterpret_cast<uint4 *>(tile + sw128_off(r, chunk)) = *reinterpret_cast<uint4 *>(hi);
        *reinterpret_cast<uint4 *>(tile + TN * 128 + sw128_off(r, chunk)) = *reinterpret_cast<uint4 *>(lo);
    } else {
        __align__(16) __half a[8];
        const __half one = __float2half_rn(AUG_ONE), zero = __float2half_rn(0.f);
        a[0] = a[1] = a[2] = one;
        if (j < k) {
            split3(cn[j] * S2A, a[3], a[4], a[5]);
        } else {
            a[3] = __float2half_rn(PAD_NORM), a[4] = zero, a[5] = zero;
        }
        a[6] = a[7] = zero;
        unsigned char *aug = tile + 2 * TN * 128;
        *reinterpret_cast<uint4 *>(aug + aug_off(r, 0)) = *reinterpret_cast<uint4 *>(a);
        *reinterpret_cast<uint4 *>(aug + aug_off(r, 1)) = make_uint4(0, 0, 0, 0);
    }
}

// ------------------------------------------------------------------------------------------ main kernel
// canonical fp32 distance of row x (registers, already normalised) to centroid j
__device__ __forceinline__ float exact_dist(const float (&xr)[64], float xn, const float *__restrict__ c,
                                            const float *__restrict__ cn, int j) {
    const float4 *cj = reinterpret_cast<const float4 *>(c + (size_t)j * 64);
    float q[16];
#pragma unroll
    for (int l = 0; l < 16; l++) {
        float4 v = __ldg(cj + l);
        float s = xr[4 * l] * v.x;
        s = fmaf(xr[4 * l + 1], v.y, s);
        s = fmaf(xr[4 * l + 2], v.z, s);
        s = fmaf(xr[4 * l + 3], v.w, s);
        q[l] = s;
    }
    return l2_expanded(xn, __ldg(cn + j), tree16(q));
}

// 32 accumulator columns -> packed (distance | column) keys -> minima of groups of four adjacent columns (FMNMX3 +
// FMNMX: half an ALU op per score) -> running top-2 over GROUP minima, two independent chains.
// ALU-pipe budget: 1 LOP3 + 0.5 + 0.31 ops per score (the exact per-column top-2 needs 3.5 and is ALU-bound).
// The best column overall is always the minimum of the best group; the runner-up is the minimum of the second-best
// group unless it sits in the best group itself (3 of K-1 positions) -- see the tie note in the file header.
__device__ __forceinline__ void fold32(const uint32_t (&r)[32], int cb, float (&t1)[2], float (&t2)[2]) {
#pragma unroll
    for (int gp = 0; gp < 4; gp++) {
        const int e = 8 * gp, ch = gp & 1;
        float kx[8];
#pragma unroll
        for (int i = 0; i < 8; i++) kx[i] = __uint_as_float((r[e + i] & 0xFFFFFF80u) | (uint32_t)(cb + e + i));
        const float a = fminf(fmin3(kx[0], kx[1], kx[2]), kx[3]);
        const float b = fminf(fmin3(kx[4], kx[5], kx[6]), kx[7]);
        const float lo = fminf(a, b), hi = fmaxf(a, b);
        t2[ch] = fmin3(t2[ch], hi, fmaxf(t1[ch], lo));
        t1[ch] = fminf(t1[ch], lo);
    }
}
// merge chain b into chain a
__device__ __forceinline__ void merge_top2(float &a1, float &a2, float b1, float b2) {
    const float lo = fminf(a1, b1), hi = fmaxf(a1, b1);
    a2 = fmin3(hi, a2, b2);
    a1 = lo;
}

// RESIDENT: blockIdx.x % nslices selects the centroid slice [tile0, tile0 + ntl), kept in shared memory.
template <bool RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_assign_tc(const float *__restrict__ x, int64_t n, int l2norm, const unsigned char *__restrict__ op, int ktiles,
            int nslices, int k, const float *__restrict__ c, const float *__restrict__ cn,
            const float *__restrict__ scale, int32_t *__restrict__ labels32, int64_t *__restrict__ labels64,
            float *__restrict__ dist, float *__restrict__ part_dist, int32_t *__restrict__ part_lab) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint32_t s_tmem_base;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *sm = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t bar0 = base + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    unsigned char *flags = sm + OFF_FLAGS;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slice = RESIDENT ? (int)(blockIdx.x % nslices) : 0;
    const int worker = RESIDENT ? (int)(blockIdx.x / nslices) : (int)blockIdx.x;
    const int workers = RESIDENT ? (int)(gridDim.x / nslices) : (int)gridDim.x;
    const int tile0 = RESIDENT ? slice * B_SLOTS : 0;                       // first centroid tile of this CTA
    const int ntl = RESIDENT ? min(B_SLOTS, ktiles - tile0) : ktiles;       // centroid tiles this CTA visits per row tile
    const int64_t ntiles = (n + TM - 1) / TM;
    const int64_t my_tiles = worker < ntiles ? (ntiles - worker + workers - 1) / workers : 0;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(BAR(BAR_A_FULL + i), 4);
            mbar_init(BAR(BAR_A_EMPTY + i), 1);
            mbar_init(BAR(BAR_ACC_FULL + i), 1);
            mbar_init(BAR(BAR_ACC_EMPTY + i), 4);
            mbar_init(BAR(BAR_CAND_FULL + i), 4);
            mbar_init(BAR(BAR_CAND_EMPTY + i), 4);
        }
        for (int i = 0; i < B_SLOTS; i++) {
            mbar_init(BAR(BAR_B_FULL + i), 1);
            mbar_init(BAR(BAR_B_EMPTY + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    if (warp == 3) {
        // ================================================================== centroid-tile producer
        if (lane == 0 && my_tiles > 0) {
            if (RESIDENT) {
                for (int jt = 0; jt < ntl; jt++) {
                    mbar_expect_tx(BAR(BAR_B_FULL + jt), B_TILE_BYTES);
                    bulk_g2s(base + OFF_B + jt * B_TILE_BYTES, op + (size_t)(tile0 + jt) * B_TILE_BYTES, B_TILE_BYTES,
                             BAR(BAR_B_FULL + jt));
                }
            } else {
                uint32_t s = 0;
                for (int64_t i = 0; i < my_tiles; i++) {
                    for (int jt = 0; jt < ktiles; jt++, s++) {
                        const uint32_t st = s % B_SLOTS, ph = (s / B_SLOTS) & 1;
                        mbar_wait(BAR(BAR_B_EMPTY + st), ph ^ 1);
                        mbar_expect_tx(BAR(BAR_B_FULL + st), B_TILE_BYTES);
                        bulk_g2s(base + OFF_B + st * B_TILE_BYTES, op + (size_t)jt * B_TILE_BYTES, B_TILE_BYTES,
                                 BAR(BAR_B_FULL + st));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        if (lane == 0) {
            uint32_t s = 0, u = 0;
            for (int64_t i = 0; i < my_tiles; i++) {
                const uint32_t ab = (uint32_t)(i & 1);
                mbar_wait(BAR(BAR_A_FULL + ab), (uint32_t)((i >> 1) & 1));
                const uint32_t a_hi = base + OFF_A + ab * A_BUF_BYTES, a_lo = a_hi + A_MAIN_BYTES, a_aug = a_lo + A_MAIN_BYTES;
                for (int jt = 0; jt < ntl; jt++, s++, u++) {
                    const uint32_t st = RESIDENT ? (uint32_t)jt : s % B_SLOTS;
                    const uint32_t bph = RESIDENT ? 0u : (s / B_SLOTS) & 1;
                    const uint32_t buf = u & 1, aph = (u >> 1) & 1;
                    if (!RESIDENT || i == 0) mbar_wait(BAR(BAR_B_FULL + st), bph);
                    mbar_wait(BAR(BAR_ACC_EMPTY + buf), aph ^ 1);
                    tc_fence_after();
                    const uint32_t b_hi = base + OFF_B + st * B_TILE_BYTES, b_lo = b_hi + TN * 128, b_aug = b_lo + TN * 128;
                    const uint32_t d = tmem + buf * TN;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) umma_f16(d, desc_sw128(a_hi + kk * 32), desc_sw128(b_hi + kk * 32), IDESC, kk > 0);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) umma_f16(d, desc_sw128(a_hi + kk * 32), desc_sw128(b_lo + kk * 32), IDESC, 1);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) umma_f16(d, desc_sw128(a_lo + kk * 32), desc_sw128(b_hi + kk * 32), IDESC, 1);
                    umma_f16(d, desc_nosw(a_aug), desc_nosw(b_aug), IDESC, 1);
                    if (!RESIDENT) umma_commit(BAR(BAR_B_EMPTY + st));
                    umma_commit(BAR(BAR_ACC_FULL + buf));
                }
                umma_commit(BAR(BAR_A_EMPTY + ab));
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ================================================================== converters
        const int cw = warp - 4;
        const int half = lane >> 4, g = lane & 15;
        const float S = scale[0], S2A = scale[1];
        for (int64_t i = 0; i < my_tiles; i++) {
            const int64_t tile = worker + i * workers;
            const int rows = (int)min((int64_t)TM, n - tile * TM);
            const uint32_t ab = (uint32_t)(i & 1);
            const float4 *xg = reinterpret_cast<const float4 *>(x + tile * TM * 64);
            // issue the global loads before waiting for the buffer: they do not depend on it
            float4 vv[16];
#pragma unroll
            for (int it = 0; it < 16; it++) {
                const int r = cw * 32 + it * 2 + half;
                vv[it] = r < rows ? __ldg(xg + r * 16 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(BAR(BAR_A_EMPTY + ab), (uint32_t)(((i >> 1) & 1) ^ 1));
            unsigned char *a_hi = sm + OFF_A + ab * A_BUF_BYTES, *a_lo = a_hi + A_MAIN_BYTES, *a_aug = a_lo + A_MAIN_BYTES;
#pragma unroll
            for (int it = 0; it < 16; it++) {
                const int r = cw * 32 + it * 2 + half;
                float4 v = vv[it];
                if (l2norm) {
                    float q = v.x * v.x;
                    q = fmaf(v.y, v.y, q), q = fmaf(v.z, v.z, q), q = fmaf(v.w, v.w, q);
                    const float den = l2_denominator(half16_sum(q));
                    v.x = __fdiv_rn(v.x, den), v.y = __fdiv_rn(v.y, den), v.z = __fdiv_rn(v.z, den), v.w = __fdiv_rn(v.w, den);
                }
                float q = v.x * v.x;
                q = fmaf(v.y, v.y, q), q = fmaf(v.z, v.z, q), q = fmaf(v.w, v.w, q);
                const float xn = half16_sum(q);
                float sx = S * v.x, sy = S * v.y, sz = S * v.z, sw = S * v.w;
                float amax = fmaxf(fmaxf(fabsf(sx), fabsf(sy)), fmaxf(fabsf(sz), fabsf(sw)));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 8));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                const bool fallback = !(amax <= ROW_LIMIT);  // also catches NaN
                if (fallback) sx = sy = sz = sw = 0.f;
                __half2 h01 = __floats2half2_rn(sx, sy), h23 = __floats2half2_rn(sz, sw);
                float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                __half2 l01 = __floats2half2_rn(sx - f01.x, sy - f01.y), l23 = __floats2half2_rn(sz - f23.x, sw - f23.y);
                const uint32_t off = sw128_off(r, g >> 1) + (uint32_t)(g & 1) * 8u;
                *reinterpret_cast<uint2 *>(a_hi + off) = make_uint2(*reinterpret_cast<uint32_t *>(&h01), *reinterpret_cast<uint32_t *>(&h23));
                *reinterpret_cast<uint2 *>(a_lo + off) = make_uint2(*reinterpret_cast<uint32_t *>(&l01), *reinterpret_cast<uint32_t *>(&l23));
                if (g == 0) {
                    __align__(16) __half a[8];
                    const __half one = __float2half_rn(AUG_ONE), zero = __float2half_rn(0.f);
                    split3(fallback ? 0.f : xn * S2A, a[0], a[1], a[2]);
                    a[3] = a[4] = a[5] = one;
                    a[6] = a[7] = zero;
                    *reinterpret_cast<uint4 *>(a_aug + aug_off(r, 0)) = *reinterpret_cast<uint4 *>(a);
                    *reinterpret_cast<uint4 *>(a_aug + aug_off(r, 1)) = make_uint4(0, 0, 0, 0);
                    flags[ab * 128 + r] = fallback ? 1 : 0;
                }
            }
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_A_FULL + ab));
        }
    } else if (warp >= 8 && warp < 12) {
        // ================================================================== epilogue: accumulator scan
        const int ew = warp - 8;  // == warp % 4: the TMEM lane quadrant this warp may read
        const int row_in_tile = ew * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        constexpr float BIG = 3.0e38f;
        int2 *cand = reinterpret_cast<int2 *>(sm + OFF_CAND);
        const float inv_s2 = scale[2], tau_abs = scale[3];
        uint32_t u = 0;
        for (int64_t i = 0; i < my_tiles; i++) {
            float g1 = BIG, g2 = BIG;
            int j1 = 0, j2 = 0;
            int fallback = 0;
            for (int jt = 0; jt < ntl; jt++, u++) {
                const uint32_t buf = u & 1, ph = (u >> 1) & 1;
                mbar_wait(BAR(BAR_ACC_FULL + buf), ph);
                tc_fence_after();
                // the converters may rewrite this flag slot as soon as MMA(i) retires: read it now
                if (jt == 0) fallback = flags[(i & 1) * 128 + row_in_tile];
                float c1[2] = {BIG, BIG}, c2[2] = {BIG, BIG};
                const uint32_t ta = tmem + lane_addr + buf * TN;
                uint32_t ra[32], rb[32];
                tmem_ld32(ta, ra);
                tmem_ld32(ta + 32, rb);
                tmem_ld_wait();
                fold32(ra, 0, c1, c2);
                tmem_ld32(ta + 64, ra);
                fold32(rb, 32, c1, c2);
                tmem_ld32(ta + 96, rb);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(BAR_ACC_EMPTY + buf));  // accumulator is in registers: free it early
                fold32(ra, 64, c1, c2);
                fold32(rb, 96, c1, c2);
                merge_top2(c1[0], c2[0], c1[1], c2[1]);
                const float t1 = c1[0], t2 = c2[0];
                if (t1 < g1) {
                    if (t2 < g1) g2 = t2, j2 = jt; else g2 = g1, j2 = j1;
                    g1 = t1, j1 = jt;
                } else if (t1 < g2) {
                    g2 = t1, j2 = jt;
                }
            }
            // Runner-up safely behind the winner (beyond what the split-fp16 product can get wrong)?  Then the winner
            // is final and its distance is read off the accumulator; otherwise the re-check warps decide in fp32.
            const int64_t tile = worker + i * workers;
            const int64_t row = tile * TM + row_in_tile;
            const float v1 = __uint_as_float(__float_as_uint(g1) & 0xFFFFFF80u);
            const float v2 = __uint_as_float(__float_as_uint(g2) & 0xFFFFFF80u);
            const int ca = (tile0 + j1) * TN + (int)(__float_as_uint(g1) & 127u);
            const int cb = (tile0 + j2) * TN + (int)(__float_as_uint(g2) & 127u);
            const bool sliced = RESIDENT && nslices > 1;  // slices are merged on exact distances
            const bool safe = !fallback && !sliced && (v2 - v1 > fmaf(fabsf(v1), 1.52587890625e-5f, tau_abs)) && ca < k;
            if (safe && row < n) {
                if (labels32) labels32[row] = ca;
                if (labels64) labels64[row] = ca;
                if (dist) dist[row] = fmaxf(v1, 0.f) * inv_s2;
            }
            const uint32_t cbuf = (uint32_t)(i & 1);
            mbar_wait(BAR(BAR_CAND_EMPTY + cbuf), (uint32_t)(((i >> 1) & 1) ^ 1));
            cand[cbuf * 128 + row_in_tile] = make_int2(safe ? -1 : (ca | (fallback << 30)), cb);
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_CAND_FULL + cbuf));
        }
    } else if (warp >= 12) {
        // ================================================================== fp32 re-check + outputs
        const int row_in_tile = (warp - 12) * 32 + lane;
        const int2 *cand = reinterpret_cast<const int2 *>(sm + OFF_CAND);
        for (int64_t i = 0; i < my_tiles; i++) {
            const int64_t tile = worker + i * workers;
            const int64_t row = tile * TM + row_in_tile;
            const uint32_t cbuf = (uint32_t)(i & 1);
            mbar_wait(BAR(BAR_CAND_FULL + cbuf), (uint32_t)((i >> 1) & 1));
            int2 cc = cand[cbuf * 128 + row_in_tile];
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(BAR_CAND_EMPTY + cbuf));
            const bool live = row < n && cc.x >= 0;   // cc.x < 0: the scan already wrote this row
            if (!__any_sync(0xffffffffu, live)) continue;
            const int fallback = (cc.x >> 30) & 1;
            cc.x &= 0x3FFFFFFF;
            float xr[64];
            float xn = 0.f;
            if (live) {
                const float4 *xp = reinterpret_cast<const float4 *>(x + row * 64);
#pragma unroll
                for (int l = 0; l < 16; l++) {
                    float4 v = __ldg(xp + l);
                    xr[4 * l] = v.x, xr[4 * l + 1] = v.y, xr[4 * l + 2] = v.z, xr[4 * l + 3] = v.w;
                }
                float q[16];
                if (l2norm) {
#pragma unroll
                    for (int l = 0; l < 16; l++) {
                        float s = xr[4 * l] * xr[4 * l];
                        s = fmaf(xr[4 * l + 1], xr[4 * l + 1], s), s = fmaf(xr[4 * l + 2], xr[4 * l + 2], s), s = fmaf(xr[4 * l + 3], xr[4 * l + 3], s);
                        q[l] = s;
                    }
                    const float den = l2_denominator(tree16(q));
#pragma unroll
                    for (int t = 0; t < 64; t++) xr[t] = __fdiv_rn(xr[t], den);
                }
#pragma unroll
                for (int l = 0; l < 16; l++) {
                    float s = xr[4 * l] * xr[4 * l];
                    s = fmaf(xr[4 * l + 1], xr[4 * l + 1], s), s = fmaf(xr[4 * l + 2], xr[4 * l + 2], s), s = fmaf(xr[4 * l + 3], xr[4 * l + 3], s);
                    q[l] = s;
                }
                xn = tree16(q);
            }
            if (live) {
                int best = 0;
                float bd = INFINITY;
                if (!fallback) {
                    int ca = cc.x, cb = cc.y;
                    if (ca >= k) ca = tile0 * TN;  // cannot happen for finite data; keeps the loads in bounds
                    if (cb >= k) cb = ca;
                    const float da = exact_dist(xr, xn, c, cn, ca);
                    const float db = exact_dist(xr, xn, c, cn, cb);
                    const bool take_b = db < da || (db == da && cb < ca);
                    best = take_b ? cb : ca;
                    bd = take_b ? db : da;
                } else {  // out-of-range row: exact scan of this CTA's centroid range (rare)
                    const int jend = min(k, (tile0 + ntl) * TN);
                    for (int j = tile0 * TN; j < jend; j++) {
                        const float dj = exact_dist(xr, xn, c, cn, j);
                        if (dj < bd) bd = dj, best = j;
                    }
                }
                if (RESIDENT && nslices > 1) {
                    part_dist[(size_t)slice * n + row] = bd;
                    part_lab[(size_t)slice * n + row] = best;
                } else {
                    if (labels32) labels32[row] = best;
                    if (labels64) labels64[row] = best;
                    if (dist) dist[row] = bd;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
    }
}

// minimum over centroid slices (slices are in ascending label order: strict '<' keeps the lowest index on ties)
__global__ void k_tc_merge(const float *__restrict__ part_dist, const int32_t *__restrict__ part_lab, int64_t n,
                           int nslices, int32_t *__restrict__ labels32, int64_t *__restrict__ labels64,
                           float *__restrict__ dist) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float bd = part_dist[i];
    int best = part_lab[i];
    for (int s = 1; s < nslices; s++) {
        const float d = part_dist[(size_t)s * n + i];
        if (d < bd) bd = d, best = part_lab[(size_t)s * n + i];
    }
    if (labels32) labels32[i] = best;
    if (labels64) labels64[i] = best;
    if (dist) dist[i] = bd;
}

// ------------------------------------------------------------------------------------------ host side
bool assign_tc_supported(const at_index *ix) { return ix->d == 64 && ix->k >= 16 && ix->op != nullptr; }

int assign_tc_prepare(at_index *ix, cudaStream_t st) {
    if (!ix->tc_scale) AT_CUDA_OK(cudaMalloc(&ix->tc_scale, 4 * sizeof(float)));
    k_tc_scale<<<1, 1024, 0, st>>>(ix->c, ix->k * 64, ix->tc_scale);
    AT_LAUNCH_OK();
    const int total = ix->ktiles * TN * 9;
    k_tc_prep<<<(total + 255) / 256, 256, 0, st>>>(ix->c, ix->cn, ix->k, ix->ktiles, ix->tc_scale,
                                                  reinterpret_cast<unsigned char *>(ix->op));
    AT_LAUNCH_OK();
    return AT_OK;
}

int assign_tc_search(at_index *ix, const float *x, int64_t n, int l2norm_rows, int32_t *labels32, int64_t *labels64,
                     float *dist, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) {
        set_error("search: tensor path needs 16-byte aligned rows");
        return AT_ERR_UNSUPPORTED;
    }
    static bool configured = false;
    if (!configured) {
        AT_CUDA_OK(cudaFuncSetAttribute(k_assign_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        AT_CUDA_OK(cudaFuncSetAttribute(k_assign_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        configured = true;
    }
    const int64_t ntiles = (n + TM - 1) / TM;
    const int sms = sm_count() > 0 ? sm_count() : 1;
    const unsigned char *op = reinterpret_cast<const unsigned char *>(ix->op);
    const int nslices = (ix->ktiles + B_SLOTS - 1) / B_SLOTS;
    const int mode = ix->tc_mode;  // 0 auto, 1 force stream, 2 force resident
    const bool resident = mode == 2 ? nslices <= sms : (mode == 1 ? false : nslices <= 1);
    if (resident) {
        int workers = sms / nslices;
        if (workers > ntiles) workers = (int)ntiles;
        if (workers < 1) workers = 1;
        if (nslices > 1 && (int64_t)nslices * n > ix->part_cap) {
            AT_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(ix->part_dist), cudaFree(ix->part_lab);
            ix->part_dist = nullptr, ix->part_lab = nullptr, ix->part_cap = 0;
            AT_CUDA_OK(cudaMalloc(&ix->part_dist, sizeof(float) * (size_t)nslices * n));
            AT_CUDA_OK(cudaMalloc(&ix->part_lab, sizeof(int32_t) * (size_t)nslices * n));
            ix->part_cap = (int64_t)nslices * n;
        }
        k_assign_tc<true><<<workers * nslices, TC_THREADS, TC_SMEM, st>>>(x, n, l2norm_rows, op, ix->ktiles, nslices, ix->k,
                                                                         ix->c, ix->cn, ix->tc_scale, labels32, labels64,
                                                                         dist, ix->part_dist, ix->part_lab);
        AT_LAUNCH_OK();
        if (nslices > 1) {
            k_tc_merge<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ix->part_dist, ix->part_lab, n, nslices, labels32,
                                                                    labels64, dist);
            AT_LAUNCH_OK();
        }
        return AT_OK;
    }
    int grid = sms;
    if (grid > ntiles) grid = (int)ntiles;
    if (grid < 1) grid = 1;
    k_assign_tc<false><<<grid, TC_THREADS, TC_SMEM, st>>>(x, n, l2norm_rows, op, ix->ktiles, 1, ix->k, ix->c, ix->cn,
                                                         ix->tc_scale, labels32, labels64, dist, nullptr, nullptr);
    AT_LAUNCH_OK();
    return AT_OK;
}

}  // namespace at
