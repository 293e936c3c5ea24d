// libat_b200: nearest-centroid search on the 5th-generation tensor cores (tcgen05 + TMEM), d == 64.
//
// Replaces the blocked sgemm + top-1 scan of faiss::exhaustive_L2sqr_blas (reached from
// processors/spec_tokenizer.py:77 and, inside faiss.Kmeans.train, processors/cluster_creator.py:54-56).
//
// Four kernels:
//   k_tc_rows    rows (fp32) -> optional L2 normalisation -> minus the centring vector m -> the fp16 operand image of the
//                rows, laid out exactly as the tensor core reads it (128-row tiles: 128x64 K-major SWIZZLE_128B + a 128x16
//                tile carrying |x - m|^2), plus per row the norm of the fp16 rounding error |delta| and Sx^2 |x - m|^2.
//                K-means builds it ONCE per training set and re-uses it for every Lloyd iteration (the rows do not
//                change); a one-off search builds it per call.
//   k_assign_tc  distances + scan (below).  Certifies ~98.6 % of the rows from the accumulators (benchmark data).
//   k_tc_tail    uncertified rows with a candidate list (~1.4 %): canonical fp32 re-evaluation of <= 4 columns, a
//                16-lane group per row, one block per candidate queue (a queue per scanning warp of k_assign_tc).
//   k_tc_full    the remaining rows (~0.02 %): exact scan of every centroid (packed fp32).
//
// Centring.  |x - c|^2 = |(x - m) - (c - m)|^2 for any vector m; the operands are the SHIFTED rows and centroids
// (x' = x - m, c' = c - m, m = mean of the centroids: at_index.cuh), because every rounding error below scales with
// |x'| |c'| instead of |x| |c| (L2-normalised mel frames: |x| = 1, |x'| ~ 0.3).  In what follows x, c, d stand for the
// shifted quantities; the tail kernels evaluate the UNSHIFTED rows and centroids with the canonical formula, whose own
// resolution is part of the threshold.
//
// Arithmetic.  For a tile of 128 rows x 128 centroids the tensor core evaluates, in ONE chain of 5 tcgen05.mma
// (kind::f16, fp32 accumulate in TMEM),
//     acc = BIAS + P * ( |x|^2 + |c|^2 - 2 <x~, c~> )         P = S^2 a power of two chosen from max |c|
//   = x~ * c~                                        (x~ = fp16(Sx x), c~ = fp16(-2 (P/Sx) c): four K = 16 steps)
//   + [xn pieces | 2^12 2^12 2^12] * [w w w | cn pieces]      (one extra K = 16 step: both norms and the bias)
// Two first-order errors.  The rounding of the ROW, acc_j - P d_j = 2 P <x - x~/Sx, c_j>, is the same vector
// delta = x - x~/Sx for every centroid, so for two candidates j, j' the error of the DIFFERENCE is bounded by
// 2 |delta| |c_j - c_j'| <= 2 |delta| (sqrt d_j + sqrt d_j') -- small exactly when the two are close competitors.  The
// rounding of the CENTROIDS, <x~, e_j> with e_j = c~_j - (-2 (P/Sx) c_j), is bounded per column by |x~| e_max
// (Cauchy-Schwarz), e_max = the largest |e_j|, measured when the operands are built (k_tc_prep).  (The scan, not the tensor
// pipe, bounds the kernel; a second product with the fp16 residual of c would remove the second term at no scan cost, but
// 9 MMAs per tile push the board into its power cap -- SPLIT_C below, off.)
//
// BIAS = 2^14 puts every accumulator into the binades [2^13, 2^17): the low 25 bits of its fp32 pattern are then a
// monotone fixed-point image of the distance, and  key = pattern * 128 + column  (ONE IMAD, fma pipe) is an unsigned
// integer whose order is the order of the distances with the column riding in the low 7 bits.  Measured pipe rates
// (tools/ubench_pipes.cu): VIMNMX / VIMNMX3 / LOP3 one warp instruction per 2 cycles per scheduler (alu pipe), IMAD
// one per 2 cycles on the fma pipe, concurrently.  The scan keeps TWO orthogonal groupings of the columns:
//     A: per tile, 8 groups of 16 adjacent columns -> exact running top-3 over the group minima (with their tiles)
//     B: 16 classes = column mod 16, kept as running minima over the WHOLE sweep -> top-3 taken once per row
// = 1.0 fma-pipe + 1.25 alu-pipe instructions per score (fold32).  An A group and a B class share exactly one column, so
// the second smallest COLUMN of a row is exactly min(second A minimum, second B minimum): a runner-up can hide behind the
// winner in one grouping, never in both.
//
// Certification (per row, accumulator units):
//   tau = 1.0625 (tau_abs + 4 S|delta| sqrt(ub) + 2 |x~| e_max + 2^-19 S^2 (|x| + |c|)^2 [unshifted: the canonical formula])
//   second smallest column further than tau from the best -> the best column is the argmin: label final.
//   else, third A minimum and third B minimum both further than tau -> every column within reach lies where one of the
//   two best A groups meets one of the two best B classes: <= 4 columns, re-evaluated by k_tc_tail with the library's
//   canonical fp32 formula (at_index.cuh, the one the exact SIMT kernel uses);
//   else (or the row is outside the fp16 / accumulator range) -> k_tc_full scans all centroids exactly.
// Labels therefore equal the exact fp32 kernel's except ties below what the fp32 formula itself resolves.
// Distances: canonical fp32 for tail rows, accumulator read-out for certified rows; at_index_search re-evaluates them
// exactly in a second pass when the caller asks for distances (k-means does not use them: its objective comes from the
// exact cluster sums).
//
// Roofline note: 2*N*K*64 algorithmic flops are executed as 1.25x that many fp16 MMA flops; the alu pipe of the scan binds
// (ncu: alu 81 % of peak, tensor 64 %, SM clock 1.62 GHz under the power cap, profiles/r02_ncu_full_summary.txt).
//
// k_assign_tc: persistent CTAs (one per SM), 16 warps; the unit of work is a SUPER TILE of 384 rows (three 128-row MMA
// tiles) so every centroid operand tile fetched from L2 is used three times and every scheduler holds three scanning
// warps (the scan is alu-pipe work: a third warp per scheduler hides its dependent-issue latency):
//   warp 0   bulk-copies (cp.async.bulk, TMA engine) centroid operand tiles into a 4-slot ring (or once, when all
//            tiles fit: RESIDENT)
//   warp 1   issues tcgen05.mma, commits to mbarriers
//   warp 2   TMEM allocation / deallocation (all 512 columns = a ring of four 128-column accumulators)
//   warp 3   bulk-copies the row operand image, one 60 KB super tile at a time, double-buffered
//   warps 4-15  epilogue: warps 4-7 scan row tile 0, 8-11 row tile 1, 12-15 row tile 2 (tcgen05.ld -> keys -> fold32);
//            an accumulator's TMEM slot is released as soon as its last columns are in registers (half way through
//            the tile's scan)
// setmaxnreg: 56 registers for warps 0-3, 152 for the scanning warps.  Producer, MMA and row-image warps run their loops
// converged and issue from one elected lane (uniform operands).
#include "at_index.cuh"
#include "at_ptx.cuh"
#include <stdlib.h>

namespace at {

#ifndef AT_TC_RT
#define AT_TC_RT 3
#endif
constexpr int TM = 128;          // rows per MMA tile (UMMA M)
constexpr int TN = 128;          // centroids per tile (UMMA N)
constexpr int RT = AT_TC_RT;     // row tiles per super tile
constexpr int TC_THREADS = 128 + RT * 128;   // 4 service warps + 4 scanning warps per row tile
constexpr int SROWS = RT * TM;   // 384
// SPLIT_C (build-time experiment, off): the centroid operand as the sum of two fp16 tiles (hi + lo, ~22 significant bits)
// multiplied in two passes (9 MMAs per accumulator instead of 5).  The certification threshold then loses its dominant term
// (the centroids' fp16 rounding) and 4-5x fewer rows reach the tail kernels, and the scan still hides the MMAs -- but the
// board reaches its power cap (SM clock 1,965 -> 1,837 MHz measured) and the whole step is 15 % slower: 5 MMAs it is.
#ifndef AT_TC_SPLIT_C
#define AT_TC_SPLIT_C 0
#endif
constexpr bool SPLIT_C = AT_TC_SPLIT_C != 0;
constexpr int B_SLOTS = SPLIT_C ? 2 : 4;   // operand tiles resident per CTA / ring depth
constexpr int ACC_SLOTS = 4;     // 128-column accumulators in TMEM, used as a ring by consecutive (centroid tile, row tile) pairs
constexpr uint32_t A_MAIN_BYTES = TM * 128;              // 128 rows x 64 fp16
constexpr uint32_t AUG_BYTES = TM * 32;                  // 128 rows x 16 fp16, no-swizzle core matrices
constexpr uint32_t A_TILE_BYTES = A_MAIN_BYTES + AUG_BYTES;      // 20,480
constexpr uint32_t A_BUF_BYTES = RT * A_TILE_BYTES;              // 61,440
constexpr uint32_t B_MAIN_BYTES = (SPLIT_C ? 2 : 1) * TN * 128;   // hi [| lo]
constexpr uint32_t B_TILE_BYTES = B_MAIN_BYTES + TN * 32;        // hi [| lo] | aug = 20,480 or 36,864

// shared memory map (dynamic, 1024-B aligned base)
constexpr uint32_t OFF_A = 0;                                    // 2 buffers
constexpr uint32_t OFF_B = OFF_A + 2 * A_BUF_BYTES;              // 2 slots
constexpr uint32_t OFF_BAR = OFF_B + B_SLOTS * B_TILE_BYTES;     // mbarriers
constexpr uint32_t TC_SMEM = OFF_BAR + 256 + 1024;               // + slack for manual 1024-B alignment
static_assert(OFF_A % 1024 == 0 && OFF_B % 1024 == 0 && A_TILE_BYTES % 1024 == 0 && B_TILE_BYTES % 1024 == 0, "align");
static_assert(TC_SMEM <= 232448, "shared memory budget");

enum {
    BAR_A_FULL = 0,      // +2   (bulk copy tx)
    BAR_A_EMPTY = 2,     // +2   (MMA commit)
    BAR_B_FULL = 4,      // +4   (bulk copy tx)
    BAR_B_EMPTY = 8,     // +4   (MMA commit)
    BAR_ACC_FULL = 12,   // +4   [accumulator slot]  (MMA commit)
    BAR_ACC_EMPTY = 16,  // +4   (4 epilogue warps)
    BAR_COUNT = 20
};
static_assert(BAR_COUNT * 8 <= 256, "barrier area");

constexpr float AUG_ONE = 4096.0f;            // 2^12, exact in fp16
constexpr float AUG_INV = 1.0f / 4096.0f;
constexpr float BIAS = 16384.0f;              // 2^14: exponent field 141 == 1 (mod 4), see key_of / acc_of
constexpr uint32_t BIAS_HI = 0x46000000u;     // the 7 pattern bits shared by [2^13, 2^17)
constexpr float SC_LIMIT = 96.0f;             // S * max|c| <= 96
constexpr float X_LIMIT = 192.0f;             // S |x| above this -> exact scan in the tail kernel
constexpr float PAD_BUMP = 28672.0f;          // padding columns sit this far (accumulator units) behind centroid k-1:
                                              // (192 + 96)^2 + PAD_BUMP + BIAS = 128,000 < 2^17, the end of the key range
constexpr float TAU_SAFETY = 1.0625f;

// scale[] layout (device floats written by k_tc_scale).  Sx is the scale of the row image (x~ = fp16(Sx x)): the
// index's own S unless a prepared image with its own scale is attached (k-means), R = S / Sx.
// SC_CANON = S (|m| + max |c_j|): with S |x - m| it bounds S (|x| + |c|), the scale of the canonical fp32 formula's own
// rounding error (the tail kernels evaluate the UNSHIFTED rows and centroids).
enum { SC_S = 0, SC_EMAX = 1, SC_INV_S2 = 2, SC_TAU = 3, SC_CMAX = 4, SC_SX = 5, SC_RATIO = 6, SC_CANON = 7, SC_CCOEF = 8,
       SC_COUNT = 12 };

// (mbarrier / bulk-copy wrappers: at_ptx.cuh)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// The MMA / commit / bulk-copy wrappers are executed by a whole converged warp with warp-uniform operands and issue
// from one elected lane: the operands then live in uniform registers (a lane-0 branch instead makes the compiler move
// every descriptor through R2UR inside a per-instruction election loop, ~100 cycles per MMA).
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(bar) : "memory");
}
// expect_tx + bulk copy from one elected lane of a converged warp
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
        "}" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// arrive from one elected lane of a converged warp
__device__ __forceinline__ void mbar_arrive_elect(uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.shared::cta.b64 _, [%0];\n\t"
        "}" ::"r"(bar) : "memory");
}
// one non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // test_wait never suspends the warp
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

// blocking wait of the pipeline roles (build-time AT_TC_WAIT_NS: 0 = the default polling form)
#ifndef AT_TC_WAIT_NS
#define AT_TC_WAIT_NS 0
#endif
__device__ __forceinline__ void mbar_waitx(uint32_t bar, uint32_t parity) {
    if (AT_TC_WAIT_NS > 0) mbar_wait_hint(bar, parity, (uint32_t)AT_TC_WAIT_NS);
    else mbar_wait(bar, parity);
}

__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }  // VIMNMX3.U32
// accumulator value of an A key (pattern * 128 + id) / a B key (pattern * 16 + class), exact: the pattern bits pushed out
// by the multiplication are the same for every binade in range (BIAS_HI); an all-ones key (nothing seen) maps above the range
__device__ __forceinline__ float acc_a(uint32_t key) { return __uint_as_float(BIAS_HI | (key >> 7)); }
__device__ __forceinline__ float acc_b(uint32_t key) { return __uint_as_float(BIAS_HI | (key >> 4)); }

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
// S = 2^s, the largest power of two with S * max|c_j| <= SC_LIMIT (so P d + BIAS stays inside [2^13, 2^17) for every
// row with S |x| <= X_LIMIT); tau = absolute part of the certification threshold in accumulator units.
// maxes: bit patterns of {max |c_ij - m_i|, max |c_j - m|^2, max |c_j|^2} (k_centroid_norms).
__global__ void k_tc_scale(unsigned int *__restrict__ maxes, const float *__restrict__ ext_sx, const float *__restrict__ shift,
                           int d, float *__restrict__ scale) {
    if (threadIdx.x == 0) {
        const float m = __uint_as_float(maxes[0]), n2 = __uint_as_float(maxes[1]), n2_orig = __uint_as_float(maxes[2]);
        maxes[0] = 0u, maxes[1] = 0u, maxes[2] = 0u;   // ready for the next set of centroids
        float m2 = 0.f;
        for (int t = 0; t < d; t++) m2 = fmaf(shift[t], shift[t], m2);
        const float cmax = sqrtf(n2) * 1.0009765625f;
        int e = 0;
        if (cmax > 0.f && isfinite(cmax)) {
            int ex;
            frexpf(cmax, &ex);   // cmax = f * 2^ex, f in [0.5, 1)
            e = 6 - ex;          // cmax * 2^e in [32, 64)
            if (ldexpf(cmax, e + 1) <= SC_LIMIT) e++;   // up to (48, 96]
            e = max(-40, min(40, e));
        }
        float S = ldexpf(1.0f, e);
        float Sx = S;
        if (ext_sx) {
            // an attached row image fixes Sx: the c-side norm weights 4096 R^2 must stay fp16 numbers, R = S / Sx <= 1
            // keeps them exact; a smaller S only costs accumulator resolution
            Sx = ext_sx[0];
            if (!(Sx > 0.f) || !isfinite(Sx)) Sx = S;
            if (S > Sx) S = Sx;
            if (S < Sx * 0.0009765625f) S = Sx * 0.0009765625f;
        }
        scale[SC_S] = S;
        scale[SC_EMAX] = 0.f;   // max_j |fp16(-2 Sc c_j) - (-2 Sc c_j)|, raised by k_tc_prep
        scale[SC_INV_S2] = 1.0f / (S * S);
        // the three-piece norms carry ~2^-21 of S^2 (|x|^2 + |c|^2), fp32 accumulation a few 2^-24 of the partial sums
        // (|.| <= 2^17): 2^-19 S^2 64 m^2 (>= 2^-19 S^2 |c|^2) plus 8 ulps of the accumulator leaves a factor ~4
        // (d > 64: the accumulator collects d / 16 + 1 products instead of 5, its rounding grows in proportion)
        scale[SC_TAU] = ldexpf(S * S * (float)d * m * m, -19) + 0.0625f * (float)(d / 16 + 1) / 5.0f;
        scale[SC_CMAX] = S * cmax;
        scale[SC_SX] = Sx;
        scale[SC_RATIO] = S / Sx;
        scale[SC_CANON] = S * (sqrtf(m2) + sqrtf(n2_orig)) * 1.001f;
        // wide rows: the exact kernel (k_assign_gemm) sums the inner product as ONE d-term FMA chain, error <= d u |x||c| with
        // u = 2^-24; |c|^2 comes from 16 chains of d / 16 terms + a 4-level tree; |x|^2 is the same number for both
        // candidates of a comparison and cancels.  Difference of two candidates: 2 u (2 d |x||c| + (d / 16 + 5) |c|^2 + |x|^2)
        // <= 1.25 d u (|x| + |c|)^2 for d >= 128
        scale[SC_CCOEF] = 1.25f * ldexpf((float)d, -24);
    }
}

// eight lanes per padded centroid: lane `chunk` rounds 8 elements of -2 Sc c to fp16 and writes its 16-byte chunk of the
// K-major tile; the eight partial sums of the squared rounding error are combined in the group and the largest
// per-centroid error norm goes to scale[SC_EMAX] (non-negative floats order like their bit patterns); lane 0 also
// writes the aug chunk.  Padding columns (j >= k) repeat centroid k-1 with PAD_BUMP added to the norm: always behind the
// real column by far more than any threshold, inside the key range, skipped by the tail kernel.
__global__ void __launch_bounds__(256) k_tc_prep(const float *__restrict__ c, const float *__restrict__ shift, int k, int ktiles,
                                                 float *__restrict__ scale, unsigned char *__restrict__ op) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = idx >> 3, chunk = idx & 7;
    if (j >= ktiles * TN) return;   // whole warps: ktiles * TN * 8 is a multiple of 256
    const float S = scale[SC_S], R = scale[SC_RATIO];
    const float Sc = S * R;   // x~ carries Sx: -2 Sc c with Sx Sc = S^2
    unsigned char *tile = op + (size_t)(j / TN) * B_TILE_BYTES;
    const int r = j % TN;
    const int js = j < k ? j : k - 1;
    __align__(16) __half hi[8], lo[8];
    float e2 = 0.f, cn2 = 0.f;   // cn2: |c_j - m|^2 (need not be canonical: its rounding is inside tau_abs)
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const float cs = c[(size_t)js * 64 + chunk * 8 + e] - shift[chunk * 8 + e];
        cn2 = fmaf(cs, cs, cn2);
        const float v = -2.0f * Sc * cs;
        hi[e] = __float2half_rn(v);
        float err = v - __half2float(hi[e]);
        if (SPLIT_C) {
            lo[e] = __float2half_rn(err);
            err -= __half2float(lo[e]);
        }
        e2 = fmaf(err, err, e2);
    }
    *reinterpret_cast<uint4 *>(tile + sw128_off(r, chunk)) = *reinterpret_cast<uint4 *>(hi);
    if (SPLIT_C) *reinterpret_cast<uint4 *>(tile + TN * 128 + sw128_off(r, chunk)) = *reinterpret_cast<uint4 *>(lo);
    e2 += __shfl_xor_sync(0xffffffffu, e2, 1);
    e2 += __shfl_xor_sync(0xffffffffu, e2, 2);
    e2 += __shfl_xor_sync(0xffffffffu, e2, 4);
    cn2 += __shfl_xor_sync(0xffffffffu, cn2, 1);
    cn2 += __shfl_xor_sync(0xffffffffu, cn2, 2);
    cn2 += __shfl_xor_sync(0xffffffffu, cn2, 4);
    float en = sqrtf(e2) * 1.001f;
    if (!(en == en)) en = INFINITY;   // a non-finite centroid: nothing is certified
    en = fmaxf(en, __shfl_xor_sync(0xffffffffu, en, 8));
    en = fmaxf(en, __shfl_xor_sync(0xffffffffu, en, 16));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int *>(scale + SC_EMAX), __float_as_uint(en));
    if (chunk == 0) {
        // row side: [pieces of xn Sx^2 / 4096 | 4096 4096 4096]; this side: [w w w | pieces of (cn S^2 + BIAS) / 4096],
        // w = 4096 R^2 (a power of two <= 4096)
        __align__(16) __half a[8];
        const __half w = __float2half_rn(AUG_ONE * R * R), zero = __float2half_rn(0.f);
        a[0] = a[1] = a[2] = w;
        split3(fmaf(cn2, S * S * AUG_INV, (j < k ? BIAS : BIAS + PAD_BUMP) * AUG_INV), a[3], a[4], a[5]);
        a[6] = a[7] = zero;
        unsigned char *aug = tile + B_MAIN_BYTES;
        *reinterpret_cast<uint4 *>(aug + aug_off(r, 0)) = *reinterpret_cast<uint4 *>(a);
        *reinterpret_cast<uint4 *>(aug + aug_off(r, 1)) = make_uint4(0, 0, 0, 0);
    }
}

// ------------------------------------------------------------------------------------------ row image
// 4 lanes per row (lane jq holds floats 16 jq .. 16 jq + 15), 8 rows per warp step.  Rows beyond n are zero.
// The reductions here need not follow the canonical association: a row that differs from the canonically normalised
// one in the last bit is inside the certification threshold (the 1e-6 |Sx x| term of erow).
// The image is that of x - m (m = shift, see at_index.cuh): xns = Sx^2 |x - m|^2; erow bounds the image's distance from the
// exact shifted row: fp16 rounding + the fp32 rounding of the subtraction (<= 2^-24 |x - m|) + the normalisation slack, which
// scales with the UNSHIFTED norm.
__global__ void __launch_bounds__(256) k_tc_rows(const float *__restrict__ x, int64_t n, int64_t n_pad, int l2norm,
                                                 const float *__restrict__ sx_ptr, const float *__restrict__ shift,
                                                 unsigned char *__restrict__ img, float *__restrict__ erow,
                                                 float *__restrict__ xns) {
    const int lane = threadIdx.x & 31, rsub = lane >> 2, jq = lane & 3;
    const float Sx = sx_ptr[0];
    float ms[16];
#pragma unroll
    for (int t = 0; t < 16; t++) ms[t] = shift[16 * jq + t];
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * 8; r0 < n_pad; r0 += nwarps * 8) {   // warp-uniform trip count
        const int64_t r = r0 + rsub;
        float xs[16];
        if (r < n) {
            const float4 *xp = reinterpret_cast<const float4 *>(x + r * 64) + 4 * jq;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const float4 v = __ldg(xp + t);
                xs[4 * t] = v.x, xs[4 * t + 1] = v.y, xs[4 * t + 2] = v.z, xs[4 * t + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 16; t++) xs[t] = 0.f;
        }
        float q = 0.f;
#pragma unroll
        for (int t = 0; t < 16; t++) q = fmaf(xs[t], xs[t], q);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        float xn = q;
        if (l2norm) {
            const float inv = 1.0f / l2_denominator(q);
            float q2 = 0.f;
#pragma unroll
            for (int t = 0; t < 16; t++) xs[t] *= inv, q2 = fmaf(xs[t], xs[t], q2);
            q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
            q2 += __shfl_xor_sync(0xffffffffu, q2, 2);
            xn = q2;
        }
        const float xn_orig = xn;   // |x|^2 of the (normalised) row itself
        if (r < n) {
            float q3 = 0.f;
#pragma unroll
            for (int t = 0; t < 16; t++) xs[t] -= ms[t], q3 = fmaf(xs[t], xs[t], q3);
            q3 += __shfl_xor_sync(0xffffffffu, q3, 1);
            q3 += __shfl_xor_sync(0xffffffffu, q3, 2);
            xn = q3;
        }
        __align__(16) __half2 hh[8];
        float e2 = 0.f;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const float s0 = Sx * xs[2 * t], s1 = Sx * xs[2 * t + 1];
            hh[t] = __floats2half2_rn(s0, s1);
            const float2 f = __half22float2(hh[t]);
            const float e0 = s0 - f.x, e1 = s1 - f.y;
            e2 = fmaf(e0, e0, e2), e2 = fmaf(e1, e1, e2);
        }
        e2 += __shfl_xor_sync(0xffffffffu, e2, 1);
        e2 += __shfl_xor_sync(0xffffffffu, e2, 2);
        const int rr = (int)(r & (TM - 1));
        unsigned char *a_tile = img + (size_t)(r >> 7) * A_TILE_BYTES;
        // floats 16 jq .. 16 jq + 15 = fp16 chunks 2 jq, 2 jq + 1 of the row
        *reinterpret_cast<uint4 *>(a_tile + sw128_off(rr, 2 * jq)) = *reinterpret_cast<uint4 *>(&hh[0]);
        *reinterpret_cast<uint4 *>(a_tile + sw128_off(rr, 2 * jq + 1)) = *reinterpret_cast<uint4 *>(&hh[4]);
        if (jq == 0) {
            const float xnS = Sx * Sx * xn;   // inf / NaN for rows the fp16 image cannot hold: the epilogue sends them to the tail
            __align__(16) __half a[8];
            const __half one = __float2half_rn(AUG_ONE), zero = __float2half_rn(0.f);
            split3(xnS * AUG_INV, a[0], a[1], a[2]);
            a[3] = a[4] = a[5] = one;
            a[6] = a[7] = zero;
            unsigned char *a_aug = a_tile + A_MAIN_BYTES;
            *reinterpret_cast<uint4 *>(a_aug + aug_off(rr, 0)) = *reinterpret_cast<uint4 *>(a);
            *reinterpret_cast<uint4 *>(a_aug + aug_off(rr, 1)) = make_uint4(0, 0, 0, 0);
            erow[r] = sqrtf(e2) * 1.001f + 1e-6f * Sx * sqrtf(xn_orig) + 2e-7f * sqrtf(xnS);
            xns[r] = xnS;
        }
    }
}

// canonical fp32 distance of the row to a centroid, evaluated by a 16-lane group: lane l holds chunk l of the row
// (xv) and of the centroid (v) -- 4-element FMA chain per chunk, xor butterfly over the group: bit-identical to
// at_index.cuh's tree16 form.
__device__ __forceinline__ float coop_dist(const float4 xv, float xn, const float4 v, float cnj) {
    float s = xv.x * v.x;
    s = fmaf(xv.y, v.y, s);
    s = fmaf(xv.z, v.z, s);
    s = fmaf(xv.w, v.w, s);
    return l2_expanded(xn, cnj, half16_sum(s));
}


// min of 16 keys: 8 alu instructions
__device__ __forceinline__ uint32_t umin16(const uint32_t *k) {
    const uint32_t a = umin3(k[0], k[1], k[2]), b = umin3(k[3], k[4], k[5]), c = umin3(k[6], k[7], k[8]);
    const uint32_t d = umin3(k[9], k[10], k[11]), e = umin3(k[12], k[13], k[14]);
    return min(umin3(a, b, c), umin3(d, e, k[15]));
}

// 32 accumulator columns (two 16-column loads = two A groups, ga the index of the first one inside its tile).  The raw
// fp32 patterns are scanned as unsigned integers (every accumulator lies in [2^13, 2^17), where the pattern order is the
// value order); NO per-score key is built.  Two orthogonal groupings:
//   A: the tile's 8 groups of 16 adjacent columns -> the group minimum gets (tile mod 16, group) appended
//      (pattern * 128 + id: one IMAD per 16 scores; the seven pattern bits pushed out are the same for the whole range)
//      -> exact running top-3 over the group minima of a BLOCK of 16 tiles (2,048 centroids), merged into the row's
//      top-3 once per block;
//   B: 16 classes = column mod 16, running over ALL tiles of the row's sweep -> bp[h], the minimum pattern of class h
//      (the class is the register index; it is appended once per row, after the sweep).
// A group (tile, columns 16a .. 16a+15) and a B class share exactly one column, so two columns never share both: the
// second smallest COLUMN of a row is exactly min(second smallest A-group minimum, second smallest B-class minimum) -- a
// runner-up can hide behind the winner in one grouping, never in both -- and the arg-min COLUMN is where the best A
// group meets the best B class.
// alu pipe: 16 (A minima) + 8 (A top-3) + 16 (B) per 32 columns = 1.25 per score; fma pipe: 2 IMAD per 32 columns.
template <int G>   // G: index of the first of the two groups inside its tile (compile time: 0, 2, 4, 6)
__device__ __forceinline__ void fold32(const uint32_t (&ra)[16], const uint32_t (&rb)[16], const uint32_t tk, const uint32_t m16,
                                       const uint32_t m8, uint32_t &t1, uint32_t &t2, uint32_t &t3, uint32_t (&bp)[16]) {
    // key = pattern * 128 + (tile mod 16) * 8 + group as two chained IMADs (fma pipe, idle here) with the group as an
    // immediate: a register addend (tile * 8 + group) costs one alu-pipe VIADD per group, and the alu pipe is the one the
    // scan saturates (m16 = 16, m8 = 8 are kernel parameters so that the multiplications stay IMADs)
    const uint32_t ga0 = (umin16(ra) * m16 + tk) * m8 + (uint32_t)G, ga1 = (umin16(rb) * m16 + tk) * m8 + (uint32_t)(G + 1);
    const uint32_t lo = min(ga0, ga1), hi = max(ga0, ga1);
    t3 = umin3(t3, max(t2, lo), max(t1, hi));
    t2 = umin3(t2, hi, max(t1, lo));
    t1 = min(t1, lo);
#pragma unroll
    for (int h = 0; h < 16; h++) bp[h] = umin3(bp[h], ra[h], rb[h]);
}

// Candidate queue of scanning warp `sw` (0 .. 4 RT - 1) of CTA `worker`: the CTA scans super tiles worker, worker + workers,
// ... so the warp sees 32 rows of each of the CTA's my_tiles super tiles; the queues tile the n_pad-entry array in (CTA,
// warp) order.  Returns the first entry; the capacity is my_tiles * 32.
__device__ __host__ __forceinline__ int64_t tc_queue_base(int64_t nsuper, int workers, int worker, int sw) {
    const int64_t qd = nsuper / workers, rem = nsuper % workers;
    const int64_t before = worker * qd + (worker < rem ? worker : rem);   // super tiles of the CTAs in front
    const int64_t mine = qd + (worker < rem ? 1 : 0);
    return (before * (RT * 4) + mine * sw) * 32;
}

// ------------------------------------------------------------------------------------------ main kernel
// RESIDENT: all ktiles (<= B_SLOTS) operand tiles are loaded once and stay in shared memory.
template <bool RESIDENT>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_assign_tc(const unsigned char *__restrict__ img, const float *__restrict__ erow, const float *__restrict__ xns, int64_t n,
            const unsigned char *__restrict__ op, int ktiles, int k, const float *__restrict__ scale, uint32_t key_mul,
            int32_t *__restrict__ labels32, int64_t *__restrict__ labels64, float *__restrict__ dist,
            uint4 *__restrict__ tail, uint32_t *__restrict__ full, unsigned int *__restrict__ tail_count) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint32_t s_tmem_base;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bar0 = base + OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int worker = (int)blockIdx.x, workers = (int)gridDim.x;
    const int64_t nsuper = (n + SROWS - 1) / SROWS;
    const int64_t my_tiles = worker < nsuper ? (nsuper - worker + workers - 1) / workers : 0;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(BAR(BAR_A_FULL + i), 1);
            mbar_init(BAR(BAR_A_EMPTY + i), 1);
        }
        for (int i = 0; i < 4; i++) {
            mbar_init(BAR(BAR_ACC_FULL + i), 1);
            mbar_init(BAR(BAR_ACC_EMPTY + i), 4);
        }
        for (int i = 0; i < B_SLOTS; i++) {
            mbar_init(BAR(BAR_B_FULL + i), 1);
            mbar_init(BAR(BAR_B_EMPTY + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    // register re-partition: the producer / MMA / allocator warps (warp group 0) keep 56 registers each, the three
    // scanning warp groups take 152 (128 * 56 + 384 * 152 = the 65,536 the CTA was launched with)
    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ================================================================== centroid-tile producer (whole warp, elected lane)
        if (my_tiles > 0) {
            if (RESIDENT) {
                for (int jt = 0; jt < ktiles; jt++)
                    bulk_g2s_elect(base + OFF_B + jt * B_TILE_BYTES, op + (size_t)jt * B_TILE_BYTES, B_TILE_BYTES,
                                   BAR(BAR_B_FULL + jt));
            } else {
                uint32_t st = 0, ph = 0;
                for (int64_t i = 0; i < my_tiles; i++) {
                    for (int jt = 0; jt < ktiles; jt++) {
                        mbar_waitx(BAR(BAR_B_EMPTY + st), ph ^ 1);
                        bulk_g2s_elect(base + OFF_B + st * B_TILE_BYTES, op + (size_t)jt * B_TILE_BYTES, B_TILE_BYTES,
                                       BAR(BAR_B_FULL + st));
                        if (++st == B_SLOTS) st = 0, ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ================================================================== row-image producer
        for (int64_t i = 0; i < my_tiles; i++) {
            const uint32_t ab = (uint32_t)(i & 1);
            mbar_waitx(BAR(BAR_A_EMPTY + ab), (uint32_t)(((i >> 1) & 1) ^ 1));
            bulk_g2s_elect(base + OFF_A + ab * A_BUF_BYTES, img + (size_t)(worker + i * workers) * A_BUF_BYTES, A_BUF_BYTES,
                           BAR(BAR_A_FULL + ab));
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ================================================================== MMA issuer (whole warp, elected lane)
        uint32_t st = 0, bph = 0, u = 0;
        for (int64_t i = 0; i < my_tiles; i++) {
            const uint32_t ab = (uint32_t)(i & 1);
            mbar_waitx(BAR(BAR_A_FULL + ab), (uint32_t)((i >> 1) & 1));
            const uint32_t a0 = base + OFF_A + ab * A_BUF_BYTES;
            for (int jt = 0; jt < ktiles; jt++, u++) {
                const uint32_t slot = RESIDENT ? (uint32_t)jt : st;
                if (!RESIDENT || i == 0) mbar_waitx(BAR(BAR_B_FULL + slot), RESIDENT ? 0u : bph);
                const uint32_t b_hi = base + OFF_B + slot * B_TILE_BYTES;
                const uint64_t dB_hi = desc_sw128(b_hi), dB_lo = desc_sw128(b_hi + TN * 128), dB_aug = desc_nosw(b_hi + B_MAIN_BYTES);
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    const uint32_t a_hi = a0 + rt * A_TILE_BYTES;
                    const uint64_t dA = desc_sw128(a_hi), dA_aug = desc_nosw(a_hi + A_MAIN_BYTES);
                    const uint32_t v = u * RT + rt, acc = v % ACC_SLOTS, aph = (v / ACC_SLOTS) & 1;
                    mbar_waitx(BAR(BAR_ACC_EMPTY + acc), aph ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem + acc * TN;
                    // descriptor start addresses are in 16-byte units: a K step of 16 fp16 = 32 bytes = +2
#ifdef AT_TC_ONE_MMA   // timing experiment only (wrong results): how much of the scan's time is spent waiting for the MMAs
                    umma_f16(d, dA_aug, dB_aug, IDESC, 0);
#else
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) umma_f16(d, dA + 2 * kk, dB_hi + 2 * kk, IDESC, kk > 0);
                    if (SPLIT_C) {
#pragma unroll
                        for (int kk = 0; kk < 4; kk++) umma_f16(d, dA + 2 * kk, dB_lo + 2 * kk, IDESC, 1);
                    }
#ifndef AT_TC_NO_AUG   // timing experiment only (wrong results): the cost of the K step that carries the norms
                    umma_f16(d, dA_aug, dB_aug, IDESC, 1);
#endif
#endif
                    umma_commit(BAR(BAR_ACC_FULL + acc));
                }
                if (!RESIDENT) {
                    umma_commit(BAR(BAR_B_EMPTY + slot));
                    if (++st == B_SLOTS) st = 0, bph ^= 1;
                }
            }
            umma_commit(BAR(BAR_A_EMPTY + ab));
        }
    } else if (warp >= 4) {
        #if AT_TC_RT == 3
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
#else
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#endif
#define AT_SCAN_WIDE 0
#include "at_tc_scan.inc"
#undef AT_SCAN_WIDE
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

// ------------------------------------------------------------------------------------------ tail
// canonical row chunk (float4 chunk g of the row, normalised like normalize_vectors when asked) and its |x|^2
__device__ __forceinline__ float4 tail_row(const float *__restrict__ x, int64_t row, int l2norm, int g, float &xn) {
    float4 xv = __ldg(reinterpret_cast<const float4 *>(x + row * 64) + g);
    if (l2norm) {
        float q = xv.x * xv.x;
        q = fmaf(xv.y, xv.y, q), q = fmaf(xv.z, xv.z, q), q = fmaf(xv.w, xv.w, q);
        const float den = l2_denominator(half16_sum(q));
        xv.x = __fdiv_rn(xv.x, den), xv.y = __fdiv_rn(xv.y, den), xv.z = __fdiv_rn(xv.z, den), xv.w = __fdiv_rn(xv.w, den);
    }
    float q = xv.x * xv.x;
    q = fmaf(xv.y, xv.y, q), q = fmaf(xv.z, xv.z, q), q = fmaf(xv.w, xv.w, q);
    xn = half16_sum(q);
    return xv;
}

// Uncertified rows with a candidate list: block q works through queue q (one per scanning warp of k_assign_tc), a
// 16-lane group per row evaluating the (at most four) candidate columns with the canonical fp32 arithmetic; the lowest
// index wins exact ties like the exact kernel's scan.
__global__ void __launch_bounds__(256) k_tc_tail(const float *__restrict__ x, int l2norm, const float *__restrict__ c,
                                                 const float *__restrict__ cn, const uint4 *__restrict__ tail_all,
                                                 const unsigned int *__restrict__ tail_count, int64_t nsuper, int workers,
                                                 int32_t *__restrict__ labels32, int64_t *__restrict__ labels64,
                                                 float *__restrict__ dist, unsigned long long *__restrict__ counters) {
    const int lane = threadIdx.x & 31, half = lane >> 4, g = lane & 15;
    const int q = (int)blockIdx.x;
    const unsigned int n_cand = tail_count[2 + q];
    const uint4 *__restrict__ tail = tail_all + tc_queue_base(nsuper, workers, q / (RT * 4), q % (RT * 4));
    const unsigned int nwarps = blockDim.x >> 5;
    // warp-uniform trip count: the 16-lane reductions shuffle with the full mask
    for (unsigned int e0 = (threadIdx.x >> 5) * 2; e0 < n_cand; e0 += nwarps * 2) {
        const unsigned int e = e0 + half;
        const bool live = e < n_cand;
        const uint4 cc = tail[live ? e : e0];
        const int64_t row = (int64_t)cc.x;
        int cols[4] = {(int)(cc.y & 0xFFFFu), (int)(cc.y >> 16), (int)(cc.z & 0xFFFFu), (int)(cc.z >> 16)};
        float4 cv[4];
        float cnv[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {   // issue the candidate loads before anything depends on them
            cv[t] = __ldg(reinterpret_cast<const float4 *>(c + (size_t)cols[t] * 64) + g);
            cnv[t] = __ldg(cn + cols[t]);
        }
        float xn;
        const float4 xv = tail_row(x, row, l2norm, g, xn);
        float bd = INFINITY;
        int best = 0x7FFFFFFF;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const float dj = coop_dist(xv, xn, cv[t], cnv[t]);
            if (dj < bd || (dj == bd && cols[t] < best)) bd = dj, best = cols[t];
        }
        if (live && g == 0) {
            if (labels32) labels32[row] = best;
            if (labels64) labels64[row] = best;
            if (dist) dist[row] = bd;
        }
    }
    if (counters && threadIdx.x == 0) {
        if (n_cand) atomicAdd(&counters[0], (unsigned long long)n_cand);
        if (blockIdx.x == 0) atomicAdd(&counters[1], (unsigned long long)tail_count[1]);
    }
}

// Uncertified rows without a usable candidate list (back of the tail array): exact scan of every centroid.  A block takes
// 32 listed rows; its eight warps share them (lane = row, the row in registers in each warp), warp q scanning centroids
// 4q .. 4q+3 of every 32-centroid tile streamed through shared memory (the next tile is fetched into registers while the
// current one is scanned) -- the exact SIMT kernel's arithmetic (canonical chunk partials + xor tree); the lowest index
// wins exact ties.
// Packed fp32 (FMUL2 / FFMA2 / FADD2, sm_100): the canonical sum is 16 independent 4-term FMA chains (chunk l = elements
// 4l .. 4l+3) joined by the tree q[i] + q[i+8], a[i] + a[i+4], b[i] + b[i+2], c[0] + c[1].  Chunks 2p and 2p+1 ride in the
// two halves of a register pair, so each chain step is one packed instruction for two chunks and the first three tree
// levels are packed additions of pairs -- the same IEEE operations on the same operands in the same association, 40 issue
// slots per (row, centroid) instead of 79.  The centroid tile is stored in shared memory already interleaved for that:
// element 4l+e of a centroid sits at 8 (l >> 1) + 2e + (l & 1).
// fp32x2 on 64-bit registers (the halves are independent IEEE fp32 lanes): keeping the operands as b64 values from the load
// on makes the register pairs explicit -- built from float2 temporaries the compiler re-packs every operand with MOVs
typedef unsigned long long u64x;
__device__ __forceinline__ u64x pack2(float lo, float hi) {
    u64x r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64x v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64x mul2(u64x a, u64x b) {
    u64x r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64x fma2(u64x a, u64x b, u64x c) {
    u64x r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64x addp2(u64x a, u64x b) {
    u64x r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
constexpr int FULL_ROWS = 32, FULL_KT = 32, FULL_SLICES = 8;
__global__ void __launch_bounds__(FULL_ROWS * FULL_SLICES) k_tc_full(
    const float *__restrict__ x, int l2norm, const float *__restrict__ c, const float *__restrict__ cn, int k,
    const uint32_t *__restrict__ full, const unsigned int *__restrict__ tail_count,
    int32_t *__restrict__ labels32, int64_t *__restrict__ labels64, float *__restrict__ dist) {
    __shared__ __align__(16) float ctile[FULL_KT][64];
    __shared__ float cns[FULL_KT];
    __shared__ float s_bd[FULL_SLICES][FULL_ROWS];
    __shared__ int s_best[FULL_SLICES][FULL_ROWS];
    const int tid = threadIdx.x, r = tid & (FULL_ROWS - 1), q = tid / FULL_ROWS;
    const unsigned int n_full = tail_count[1];
    constexpr int PER = FULL_KT / FULL_SLICES;                      // centroids per warp per tile
    constexpr int LD = FULL_KT * 16 / (FULL_ROWS * FULL_SLICES);    // float4 per thread per tile
    const float4 *c4 = reinterpret_cast<const float4 *>(c);
    const int64_t c4_total = (int64_t)k * 16;
    for (unsigned int base = blockIdx.x * FULL_ROWS; base < n_full; base += gridDim.x * FULL_ROWS) {
        const unsigned int e = base + r;
        const bool live = e < n_full;
        const int64_t row = (int64_t)full[live ? e : base];
        // xp[p][e] = (x[8p + e], x[8p + 4 + e]): chunks 2p and 2p+1 side by side.  (A plain mov.b64 pack is transparent to
        // ptxas, which then keeps the 64 scalars where the loads put them and re-packs both halves before every use: 61
        // MOVs per centroid.  x + (-0) = x for every x.)
        u64x negzero2 = 0x8000000080000000ull;
        asm volatile("" : "+l"(negzero2));
        u64x xp[8][4];
        float xn;
        {
            float xr[64];
            const float4 *xq = reinterpret_cast<const float4 *>(x + row * 64);
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const float4 v = __ldg(xq + t);
                xr[4 * t] = v.x, xr[4 * t + 1] = v.y, xr[4 * t + 2] = v.z, xr[4 * t + 3] = v.w;
            }
            float qq[16];
            if (l2norm) {
#pragma unroll
                for (int l = 0; l < 16; l++) qq[l] = 0.f;
#pragma unroll
                for (int t = 0; t < 64; t++) qq[t >> 2] = fmaf(xr[t], xr[t], qq[t >> 2]);
                const float den = l2_denominator(tree16(qq));
#pragma unroll
                for (int t = 0; t < 64; t++) xr[t] = __fdiv_rn(xr[t], den);
            }
#pragma unroll
            for (int l = 0; l < 16; l++) qq[l] = 0.f;
#pragma unroll
            for (int t = 0; t < 64; t++) qq[t >> 2] = fmaf(xr[t], xr[t], qq[t >> 2]);
            xn = tree16(qq);
#pragma unroll
            for (int pp = 0; pp < 8; pp++)
#pragma unroll
                for (int ee = 0; ee < 4; ee++)   // + (-0, -0): the identity, as an instruction whose result IS a register pair
                    xp[pp][ee] = addp2(pack2(xr[8 * pp + ee], xr[8 * pp + 4 + ee]), negzero2);
        }
        float bd = INFINITY;
        int best = 0x7FFFFFFF;
        float4 pre[LD];
        float pre_cn = INFINITY;
#pragma unroll
        for (int t = 0; t < LD; t++) {
            const int64_t i4 = tid + t * (FULL_ROWS * FULL_SLICES);
            pre[t] = i4 < c4_total ? __ldg(c4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (tid < FULL_KT) pre_cn = tid < k ? cn[tid] : INFINITY;
        for (int j0 = 0; j0 < k; j0 += FULL_KT) {
            __syncthreads();   // the previous tile has been consumed
#pragma unroll
            for (int t = 0; t < LD; t++) {
                // float4 `i4` of the tile = chunk l of centroid jj -> interleaved positions 8 (l >> 1) + 2e + (l & 1)
                const int i4 = tid + t * (FULL_ROWS * FULL_SLICES), jj = i4 >> 4, l = i4 & 15;
                float *dstp = &ctile[jj][8 * (l >> 1) + (l & 1)];
                dstp[0] = pre[t].x, dstp[2] = pre[t].y, dstp[4] = pre[t].z, dstp[6] = pre[t].w;
            }
            if (tid < FULL_KT) cns[tid] = pre_cn;
            __syncthreads();
            const int jn = j0 + FULL_KT;   // fetch the next tile while this one is scanned
            if (jn < k) {
#pragma unroll
                for (int t = 0; t < LD; t++) {
                    const int64_t i4 = (int64_t)jn * 16 + tid + t * (FULL_ROWS * FULL_SLICES);
                    pre[t] = i4 < c4_total ? __ldg(c4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (tid < FULL_KT) pre_cn = jn + tid < k ? cn[jn + tid] : INFINITY;
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int jj = q * PER + u;
                const ulonglong2 *ct = reinterpret_cast<const ulonglong2 *>(&ctile[jj][0]);
                u64x sp[8];   // sp[p] = (q[2p], q[2p+1])
#pragma unroll
                for (int pp = 0; pp < 8; pp++) {
                    const ulonglong2 v0 = ct[2 * pp], v1 = ct[2 * pp + 1];
                    u64x acc = mul2(xp[pp][0], v0.x);
                    acc = fma2(xp[pp][1], v0.y, acc);
                    acc = fma2(xp[pp][2], v1.x, acc);
                    acc = fma2(xp[pp][3], v1.y, acc);
                    sp[pp] = acc;
                }
                // tree16 on pairs: a[i] = q[i] + q[i+8], b[i] = a[i] + a[i+4], c[i] = b[i] + b[i+2], c[0] + c[1]
                const u64x a0 = addp2(sp[0], sp[4]), a1 = addp2(sp[1], sp[5]);
                const u64x a2 = addp2(sp[2], sp[6]), a3 = addp2(sp[3], sp[7]);
                const u64x cc2 = addp2(addp2(a0, a2), addp2(a1, a3));
                float2 cc;
                unpack2(cc2, cc.x, cc.y);
                const float dj = l2_expanded(xn, cns[jj], __fadd_rn(cc.x, cc.y));
                if (j0 + jj < k && dj < bd) bd = dj, best = j0 + jj;   // ascending index within the slice: strict '<'
            }
        }
        s_bd[q][r] = bd, s_best[q][r] = best;
        __syncthreads();
        if (q == 0 && live) {
#pragma unroll
            for (int o = 1; o < FULL_SLICES; o++) {
                const float od = s_bd[o][r];
                const int ob = s_best[o][r];
                if (od < bd || (od == bd && ob < best)) bd = od, best = ob;
            }
            if (best == 0x7FFFFFFF) best = 0;   // every distance NaN: label 0 like the exact kernel
            if (labels32) labels32[row] = best;
            if (labels64) labels64[row] = best;
            if (dist) dist[row] = bd;
        }
    }
}

// canonical fp32 distance of every row to its label (second pass of at_index_search when distances are requested):
// a 16-lane group per row.
__global__ void __launch_bounds__(256) k_exact_dist(const float *__restrict__ x, int64_t n, int l2norm,
                                                    const float *__restrict__ c, const float *__restrict__ cn,
                                                    const int32_t *__restrict__ labels32,
                                                    const int64_t *__restrict__ labels64, float *__restrict__ dist) {
    const int g = threadIdx.x & 15;
    const int64_t groups = ((int64_t)gridDim.x * blockDim.x) >> 4;
    const int64_t nloop = (n + groups - 1) / groups;   // warp-uniform trip count (full-mask shuffles)
    int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    for (int64_t it = 0; it < nloop; it++, row += groups) {
        const bool live = row < n;
        const int64_t rr = live ? row : 0;
        float4 xv = __ldg(reinterpret_cast<const float4 *>(x + rr * 64) + g);
        if (l2norm) {
            float q = xv.x * xv.x;
            q = fmaf(xv.y, xv.y, q), q = fmaf(xv.z, xv.z, q), q = fmaf(xv.w, xv.w, q);
            const float den = l2_denominator(half16_sum(q));
            xv.x = __fdiv_rn(xv.x, den), xv.y = __fdiv_rn(xv.y, den), xv.z = __fdiv_rn(xv.z, den), xv.w = __fdiv_rn(xv.w, den);
        }
        float q = xv.x * xv.x;
        q = fmaf(xv.y, xv.y, q), q = fmaf(xv.z, xv.z, q), q = fmaf(xv.w, xv.w, q);
        const float xn = half16_sum(q);
        const int j = labels32 ? labels32[rr] : (int)labels64[rr];
        const float dj = coop_dist(xv, xn, __ldg(reinterpret_cast<const float4 *>(c + (size_t)j * 64) + g), __ldg(cn + j));
        if (live && g == 0) dist[row] = dj;
    }
}


// ====================================================================================================== wide rows
// d = 64 NS (2 <= NS <= 16: the use_convolution branch, d = 640).  Same arithmetic, same scan, same certification; the
// product is accumulated over NS slices of 64 values: an accumulator receives 4 NS + 1 MMAs (the last one carries the norms)
// before it is committed to the scanning warps.  Operand images: per 128 rows / 128 centroids NS tiles of 16 KB (the d = 64
// layout per slice) followed by the 4 KB aug tile.  Neither operand fits shared memory whole, so BOTH are streamed slice by
// slice: warp 3 brings the RT row tiles' slice (RT x 16 KB per stage, 2 stages), warp 0 the centroid tile's slice (16 KB,
// 4 stages), in the order (super tile, centroid tile, slice) -- the row slices are fetched again for every centroid tile
// (L2 / HBM traffic NS x 16 KB x RT per centroid tile; the tensor pipe, 41 MMAs per accumulator at NS = 10, is what binds).
// Uncertified rows: the candidate columns are re-checked by k_tc_tail_wide with the exact wide kernel's arithmetic, the rows
// without a usable candidate list go through k_assign_gemm over the list.
constexpr uint32_t W_MAIN = TM * 128;                    // one 64-value slice of 128 rows / centroids: 16,384 B
constexpr uint32_t W_AUG = TM * 32;                      // 4,096 B
constexpr int WA_SLOTS = 2, WB_SLOTS = 4;
constexpr uint32_t WOFF_A = 0;
constexpr uint32_t WOFF_B = WOFF_A + WA_SLOTS * RT * W_MAIN;
constexpr uint32_t WOFF_BAR = WOFF_B + WB_SLOTS * W_MAIN;
constexpr uint32_t TCW_SMEM = WOFF_BAR + 256 + 1024;
static_assert(TCW_SMEM <= 232448 && WB_SLOTS == B_SLOTS, "wide kernel shared memory / barrier layout");
__host__ __device__ inline size_t wide_tile_bytes(int ns) { return (size_t)ns * W_MAIN + W_AUG; }

__device__ __forceinline__ void bulk_expect_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_elect(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
        "}" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// rows -> wide image.  8 lanes per row (lane `chunk` owns the 16-byte chunk `chunk` of every slice), 4 rows per warp step.
__global__ void __launch_bounds__(256) k_tc_rows_wide(const float *__restrict__ x, int64_t n, int64_t n_pad, int d,
                                                      const float *__restrict__ sx_ptr, const float *__restrict__ shift,
                                                      unsigned char *__restrict__ img, float *__restrict__ erow,
                                                      float *__restrict__ xns) {
    const int lane = threadIdx.x & 31, rsub = lane >> 3, chunk = lane & 7;
    const int ns = d >> 6;
    const size_t wt = wide_tile_bytes(ns);
    const float Sx = sx_ptr[0];
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r0 = warp * 4; r0 < n_pad; r0 += nwarps * 4) {   // warp-uniform trip count
        const int64_t r = r0 + rsub;
        const int rr = (int)(r & (TM - 1));
        unsigned char *tile = img + (size_t)(r >> 7) * wt;
        float q = 0.f, q0 = 0.f, e2 = 0.f;
        for (int sl = 0; sl < ns; sl++) {
            float xs[8];
            if (r < n) {
                const float4 *xp = reinterpret_cast<const float4 *>(x + r * d + sl * 64 + chunk * 8);
                const float4 a = __ldg(xp), b = __ldg(xp + 1);
                xs[0] = a.x, xs[1] = a.y, xs[2] = a.z, xs[3] = a.w, xs[4] = b.x, xs[5] = b.y, xs[6] = b.z, xs[7] = b.w;
            } else {
#pragma unroll
                for (int t = 0; t < 8; t++) xs[t] = 0.f;
            }
            __align__(16) __half2 hh[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                float v0 = xs[2 * t], v1 = xs[2 * t + 1];
                q0 = fmaf(v0, v0, q0), q0 = fmaf(v1, v1, q0);
                if (r < n) v0 -= shift[sl * 64 + chunk * 8 + 2 * t], v1 -= shift[sl * 64 + chunk * 8 + 2 * t + 1];
                q = fmaf(v0, v0, q), q = fmaf(v1, v1, q);
                const float s0 = Sx * v0, s1 = Sx * v1;
                hh[t] = __floats2half2_rn(s0, s1);
                const float2 f = __half22float2(hh[t]);
                const float e0 = s0 - f.x, e1 = s1 - f.y;
                e2 = fmaf(e0, e0, e2), e2 = fmaf(e1, e1, e2);
            }
            *reinterpret_cast<uint4 *>(tile + (size_t)sl * W_MAIN + sw128_off(rr, chunk)) = *reinterpret_cast<uint4 *>(hh);
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            q += __shfl_xor_sync(0xffffffffu, q, o);
            q0 += __shfl_xor_sync(0xffffffffu, q0, o);
            e2 += __shfl_xor_sync(0xffffffffu, e2, o);
        }
        if (chunk == 0) {
            const float xnS = Sx * Sx * q;
            __align__(16) __half a[8];
            const __half one = __float2half_rn(AUG_ONE), zero = __float2half_rn(0.f);
            split3(xnS * AUG_INV, a[0], a[1], a[2]);
            a[3] = a[4] = a[5] = one;
            a[6] = a[7] = zero;
            unsigned char *a_aug = tile + (size_t)ns * W_MAIN;
            *reinterpret_cast<uint4 *>(a_aug + aug_off(rr, 0)) = *reinterpret_cast<uint4 *>(a);
            *reinterpret_cast<uint4 *>(a_aug + aug_off(rr, 1)) = make_uint4(0, 0, 0, 0);
            erow[r] = sqrtf(e2) * 1.001f + 1e-6f * Sx * sqrtf(q0) + 2e-7f * sqrtf(xnS);
            xns[r] = xnS;
        }
    }
}

// centroids -> wide operand image (8 lanes per padded centroid, like k_tc_prep, looping over the slices)
__global__ void __launch_bounds__(256) k_tc_prep_wide(const float *__restrict__ c, const float *__restrict__ shift, int k,
                                                      int ktiles, int d, float *__restrict__ scale,
                                                      unsigned char *__restrict__ op) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = idx >> 3, chunk = idx & 7;
    if (j >= ktiles * TN) return;
    const int ns = d >> 6;
    const float S = scale[SC_S], R = scale[SC_RATIO];
    const float Sc = S * R;
    unsigned char *tile = op + (size_t)(j / TN) * wide_tile_bytes(ns);
    const int r = j % TN;
    const int js = j < k ? j : k - 1;
    float e2 = 0.f, cn2 = 0.f;
    for (int sl = 0; sl < ns; sl++) {
        __align__(16) __half hi[8];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const float cs = c[(size_t)js * d + sl * 64 + chunk * 8 + e] - shift[sl * 64 + chunk * 8 + e];
            cn2 = fmaf(cs, cs, cn2);
            const float v = -2.0f * Sc * cs;
            hi[e] = __float2half_rn(v);
            const float err = v - __half2float(hi[e]);
            e2 = fmaf(err, err, e2);
        }
        *reinterpret_cast<uint4 *>(tile + (size_t)sl * W_MAIN + sw128_off(r, chunk)) = *reinterpret_cast<uint4 *>(hi);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
        cn2 += __shfl_xor_sync(0xffffffffu, cn2, o);
    }
    float en = sqrtf(e2) * 1.001f;
    if (!(en == en)) en = INFINITY;
    en = fmaxf(en, __shfl_xor_sync(0xffffffffu, en, 8));
    en = fmaxf(en, __shfl_xor_sync(0xffffffffu, en, 16));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int *>(scale + SC_EMAX), __float_as_uint(en));
    if (chunk == 0) {
        __align__(16) __half a[8];
        const __half w = __float2half_rn(AUG_ONE * R * R), zero = __float2half_rn(0.f);
        a[0] = a[1] = a[2] = w;
        split3(fmaf(cn2, S * S * AUG_INV, (j < k ? BIAS : BIAS + PAD_BUMP) * AUG_INV), a[3], a[4], a[5]);
        a[6] = a[7] = zero;
        unsigned char *aug = tile + (size_t)ns * W_MAIN;
        *reinterpret_cast<uint4 *>(aug + aug_off(r, 0)) = *reinterpret_cast<uint4 *>(a);
        *reinterpret_cast<uint4 *>(aug + aug_off(r, 1)) = make_uint4(0, 0, 0, 0);
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_assign_tc_wide(const unsigned char *__restrict__ img, const float *__restrict__ erow, const float *__restrict__ xns, int64_t n,
                 const unsigned char *__restrict__ op, int ktiles, int k, int ns, const float *__restrict__ scale,
                 uint32_t key_mul, int32_t *__restrict__ labels32, int64_t *__restrict__ labels64, float *__restrict__ dist,
                 uint4 *__restrict__ tail, uint32_t *__restrict__ full, unsigned int *__restrict__ tail_count) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ uint32_t s_tmem_base;
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bar0 = base + WOFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int worker = (int)blockIdx.x, workers = (int)gridDim.x;
    const int64_t nsuper = (n + SROWS - 1) / SROWS;
    const int64_t my_tiles = worker < nsuper ? (nsuper - worker + workers - 1) / workers : 0;
    const size_t wt = wide_tile_bytes(ns);

    if (tid == 0) {
        for (int i = 0; i < WA_SLOTS; i++) {
            mbar_init(BAR(BAR_A_FULL + i), 1);
            mbar_init(BAR(BAR_A_EMPTY + i), 1);
        }
        for (int i = 0; i < 4; i++) {
            mbar_init(BAR(BAR_ACC_FULL + i), 1);
            mbar_init(BAR(BAR_ACC_EMPTY + i), 4);
        }
        for (int i = 0; i < WB_SLOTS; i++) {
            mbar_init(BAR(BAR_B_FULL + i), 1);
            mbar_init(BAR(BAR_B_EMPTY + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;
    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ================================================================== centroid slices (whole warp, elected lane)
        uint32_t st = 0, ph = 0;
        for (int64_t i = 0; i < my_tiles; i++)
            for (int jt = 0; jt < ktiles; jt++)
                for (int sl = 0; sl <= ns; sl++) {
                    mbar_waitx(BAR(BAR_B_EMPTY + st), ph ^ 1);
                    bulk_g2s_elect(base + WOFF_B + st * W_MAIN, op + (size_t)jt * wt + (size_t)sl * W_MAIN, sl < ns ? W_MAIN : W_AUG,
                                   BAR(BAR_B_FULL + st));
                    if (++st == WB_SLOTS) st = 0, ph ^= 1;
                }
    } else if (warp == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ================================================================== row slices of the super tile's RT row tiles
        uint32_t st = 0, ph = 0;
        for (int64_t i = 0; i < my_tiles; i++) {
            const unsigned char *t0 = img + (size_t)(worker + i * workers) * RT * wt;
            for (int jt = 0; jt < ktiles; jt++)
                for (int sl = 0; sl <= ns; sl++) {
                    const uint32_t bytes = sl < ns ? W_MAIN : W_AUG;
                    mbar_waitx(BAR(BAR_A_EMPTY + st), ph ^ 1);
                    bulk_expect_elect(BAR(BAR_A_FULL + st), RT * bytes);
#pragma unroll
                    for (int rt = 0; rt < RT; rt++)
                        bulk_copy_elect(base + WOFF_A + st * (RT * W_MAIN) + rt * W_MAIN, t0 + (size_t)rt * wt + (size_t)sl * W_MAIN,
                                        bytes, BAR(BAR_A_FULL + st));
                    if (++st == WA_SLOTS) st = 0, ph ^= 1;
                }
        }
    } else if (warp == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // ================================================================== MMA issuer (whole warp, elected lane)
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0, u = 0;
        for (int64_t i = 0; i < my_tiles; i++) {
            for (int jt = 0; jt < ktiles; jt++, u++) {
                for (int sl = 0; sl <= ns; sl++) {
                    mbar_waitx(BAR(BAR_A_FULL + sa), pa);
                    mbar_waitx(BAR(BAR_B_FULL + sb), pb);
                    tc_fence_after();
                    const uint32_t a0 = base + WOFF_A + sa * (RT * W_MAIN), b0 = base + WOFF_B + sb * W_MAIN;
#pragma unroll
                    for (int rt = 0; rt < RT; rt++) {
                        const uint32_t v = u * RT + rt, acc = v % ACC_SLOTS, aph = (v / ACC_SLOTS) & 1;
                        if (sl == 0) {
                            mbar_waitx(BAR(BAR_ACC_EMPTY + acc), aph ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t dt = tmem + acc * TN;
                        if (sl < ns) {
                            const uint64_t dA = desc_sw128(a0 + rt * W_MAIN), dB = desc_sw128(b0);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) umma_f16(dt, dA + 2 * kk, dB + 2 * kk, IDESC, (sl > 0 || kk > 0) ? 1u : 0u);
                        } else {
                            umma_f16(dt, desc_nosw(a0 + rt * W_MAIN), desc_nosw(b0), IDESC, 1u);
                            umma_commit(BAR(BAR_ACC_FULL + acc));
                        }
                    }
                    umma_commit(BAR(BAR_A_EMPTY + sa));
                    umma_commit(BAR(BAR_B_EMPTY + sb));
                    if (++sa == WA_SLOTS) sa = 0, pa ^= 1;
                    if (++sb == WB_SLOTS) sb = 0, pb ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
#if AT_TC_RT == 3
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
#else
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#endif
#define AT_SCAN_WIDE 1
#include "at_tc_scan.inc"
#undef AT_SCAN_WIDE
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

// ------------------------------------------------------------------------------------------ host side
bool assign_tc_supported(const at_index *ix) { return ix->d == 64 && ix->k >= 16 && ix->k <= 65536 && ix->op != nullptr; }
bool assign_tc_wide_supported(const at_index *ix) {
    return ix->d > 64 && ix->d % 64 == 0 && ix->d <= 1024 && ix->k >= 16 && ix->k <= 65536 && ix->op != nullptr;
}

int assign_tc_prepare(at_index *ix, cudaStream_t st) {
    if (!ix->tc_scale) AT_CUDA_OK(cudaMalloc(&ix->tc_scale, SC_COUNT * sizeof(float)));
    const float *shift = ix->ext_shift ? ix->ext_shift : ix->shift;
    k_tc_scale<<<1, 32, 0, st>>>(ix->tc_max, ix->ext_sx, shift, ix->d, ix->tc_scale);
    AT_LAUNCH_OK();
    const int total = ix->ktiles * TN * 8;
    if (ix->d == 64)
        k_tc_prep<<<(total + 255) / 256, 256, 0, st>>>(ix->c, shift, ix->k, ix->ktiles, ix->tc_scale,
                                                      reinterpret_cast<unsigned char *>(ix->op));
    else
        k_tc_prep_wide<<<(total + 255) / 256, 256, 0, st>>>(ix->c, shift, ix->k, ix->ktiles, ix->d, ix->tc_scale,
                                                           reinterpret_cast<unsigned char *>(ix->op));
    AT_LAUNCH_OK();
    return AT_OK;
}

void tc_rows_free(at_tc_rows *r) {
    cudaFree(r->img), cudaFree(r->erow), cudaFree(r->xns), cudaFree(r->tail), cudaFree(r->full), cudaFree(r->tail_count);
    cudaFree(r->keys);
    *r = at_tc_rows();
}

// (Re)builds the operand image of x.  sx: device float, the image scale (a power of two).
// mean of the centroids (d == 64): 16 row slices x 64 columns, summed in a fixed order
// (d = 64 NS: block b takes columns 64 b .. 64 b + 63)
__global__ void __launch_bounds__(1024) k_tc_mean(const float *__restrict__ c, int k, int d, float *__restrict__ shift_all) {
    __shared__ float part[16][64];
    const int col = threadIdx.x & 63, sl = threadIdx.x >> 6;
    float *shift = shift_all + blockIdx.x * 64;
    c += blockIdx.x * 64;
    float s = 0.f;
    for (int j = sl; j < k; j += 16) s += c[(size_t)j * d + col];
    part[sl][col] = s;
    __syncthreads();
    if (sl == 0) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 16; q++) t += part[q][col];
        t /= (float)k;
        shift[col] = isfinite(t) ? t : 0.f;   // any vector is a valid centre; a non-finite one would poison every row
    }
}

int tc_mean(const float *c, int k, int d, float *shift, cudaStream_t st) {
    k_tc_mean<<<d / 64, 1024, 0, st>>>(c, k, d, shift);
    AT_LAUNCH_OK();
    return AT_OK;
}

size_t tc_operand_bytes(int d, int ktiles) {
    return d == 64 ? (size_t)ktiles * 36864 /* room for the hi | lo | aug form */ : (size_t)ktiles * wide_tile_bytes(d / 64);
}

int tc_rows_build(at_tc_rows *r, const float *x, int64_t n, int l2norm, const float *sx, const float *shift, cudaStream_t st,
                  int d) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) {
        set_error("search: tensor path needs 16-byte aligned rows");
        return AT_ERR_UNSUPPORTED;
    }
    if (n >= (1LL << 31)) {
        set_error("search: tensor path takes fewer than 2^31 rows per call");
        return AT_ERR_UNSUPPORTED;
    }
    const int64_t n_pad = (n + SROWS - 1) / SROWS * SROWS;
    if (d != 64 && l2norm) {
        set_error("search: the wide-row tensor path takes pre-normalised rows");
        return AT_ERR_UNSUPPORTED;
    }
    if (n_pad > r->cap || d != r->d) {
        AT_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(r->img), cudaFree(r->erow), cudaFree(r->xns), cudaFree(r->tail), cudaFree(r->full), cudaFree(r->keys);
        r->img = nullptr, r->erow = r->xns = nullptr, r->tail = nullptr, r->full = nullptr, r->keys = nullptr, r->cap = 0;
        AT_CUDA_OK(cudaMalloc(&r->img, (size_t)(n_pad / TM) * (d == 64 ? (size_t)A_TILE_BYTES : wide_tile_bytes(d / 64))));
        r->d = d;
        AT_CUDA_OK(cudaMalloc(&r->erow, sizeof(float) * (size_t)n_pad));
        AT_CUDA_OK(cudaMalloc(&r->xns, sizeof(float) * (size_t)n_pad));
        AT_CUDA_OK(cudaMalloc(&r->tail, sizeof(uint4) * (size_t)n_pad));
        AT_CUDA_OK(cudaMalloc(&r->full, sizeof(uint32_t) * (size_t)n_pad));
        if (d != 64) AT_CUDA_OK(cudaMalloc(&r->keys, sizeof(unsigned long long) * (size_t)n_pad));
        r->cap = n_pad;
    }
    const int sms = sm_count() > 0 ? sm_count() : 1;
    if (r->tail_queues < sms * RT * 4) {   // one candidate queue per scanning warp of the search grid (<= one CTA per SM)
        AT_CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(r->tail_count);
        r->tail_count = nullptr, r->tail_queues = 0;
        AT_CUDA_OK(cudaMalloc(&r->tail_count, sizeof(unsigned int) * (size_t)(2 + sms * RT * 4)));
        r->tail_queues = sms * RT * 4;
    }
    if (d == 64)
        k_tc_rows<<<sms * 8, 256, 0, st>>>(x, n, n_pad, l2norm, sx, shift, reinterpret_cast<unsigned char *>(r->img), r->erow,
                                           r->xns);
    else
        k_tc_rows_wide<<<sms * 8, 256, 0, st>>>(x, n, n_pad, d, sx, shift, reinterpret_cast<unsigned char *>(r->img), r->erow,
                                                r->xns);
    AT_LAUNCH_OK();
    r->x = x, r->n = n, r->l2norm = l2norm;
    return AT_OK;
}

// rows: a prepared image of exactly these rows (k-means), or nullptr to build one in the index's own workspace.
int assign_tc_search(at_index *ix, const float *x, int64_t n, int l2norm_rows, int32_t *labels32, int64_t *labels64,
                     float *dist, int exact_dist, at_tc_rows *rows, cudaStream_t st) {
    static bool configured[MAX_DEVICES] = {};   // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        AT_CUDA_OK(cudaFuncSetAttribute(k_assign_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        AT_CUDA_OK(cudaFuncSetAttribute(k_assign_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        configured[dev] = true;
    }
    if (!rows) {
        rows = &ix->rows;
        int rc = tc_rows_build(rows, x, n, l2norm_rows, ix->tc_scale + SC_SX, ix->ext_shift ? ix->ext_shift : ix->shift, st);
        if (rc != AT_OK) return rc;
    }
    // a distance request without labels still needs labels for the second pass
    int32_t *l32 = labels32;
    if (dist && exact_dist && !labels32 && !labels64) {
        if (n > ix->part_cap) {
            AT_CUDA_OK(cudaStreamSynchronize(st));
            cudaFree(ix->part_lab);
            ix->part_lab = nullptr, ix->part_cap = 0;
            AT_CUDA_OK(cudaMalloc(&ix->part_lab, sizeof(int32_t) * (size_t)n));
            ix->part_cap = n;
        }
        l32 = ix->part_lab;
    }
    const int64_t nsuper = (n + SROWS - 1) / SROWS;
    const int sms = sm_count() > 0 ? sm_count() : 1;
    int grid = sms;
    if (grid > nsuper) grid = (int)nsuper;
    if (grid < 1) grid = 1;
    const unsigned char *op = reinterpret_cast<const unsigned char *>(ix->op);
    const unsigned char *img = reinterpret_cast<const unsigned char *>(rows->img);
    const int mode = ix->tc_mode;  // 0 auto, 1 force stream, 2 resident when the tiles fit
    const bool resident = mode != 1 && ix->ktiles <= B_SLOTS;
    float *kdist = (dist && exact_dist) ? nullptr : dist;
    AT_CUDA_OK(cudaMemsetAsync(rows->tail_count, 0, 2 * sizeof(unsigned int), st));
    if (resident)
        k_assign_tc<true><<<grid, TC_THREADS, TC_SMEM, st>>>(img, rows->erow, rows->xns, n, op, ix->ktiles, ix->k,
                                                            ix->tc_scale, 128u, l32, labels64, kdist, rows->tail,
                                                            rows->full, rows->tail_count);
    else
        k_assign_tc<false><<<grid, TC_THREADS, TC_SMEM, st>>>(img, rows->erow, rows->xns, n, op, ix->ktiles, ix->k,
                                                             ix->tc_scale, 128u, l32, labels64, kdist, rows->tail,
                                                             rows->full, rows->tail_count);
    AT_LAUNCH_OK();
    // The two tail kernels touch disjoint rows: the exact scans (few rows, FMA-bound) run on the index's side stream while
    // the candidate re-checks (many rows, latency-bound) run on the caller's, joined again before anything reads the labels.
#ifndef AT_TC_TAIL_FORK
#define AT_TC_TAIL_FORK 1
#endif
    // timing experiments only (wrong labels; experiment builds with -DAT_TC_EXPERIMENTS, never the product library):
    // AT_TC_SKIP bit 0 drops the exact scans, bit 1 the candidate re-checks
#ifdef AT_TC_EXPERIMENTS
    static const int skip = getenv("AT_TC_SKIP") ? atoi(getenv("AT_TC_SKIP")) : 0;
#else
    constexpr int skip = 0;
#endif
    const bool fork = AT_TC_TAIL_FORK && ix->side_ok() && !skip;
    cudaStream_t fs = fork ? ix->side : st;
    if (fork) {
        AT_CUDA_OK(cudaEventRecord(ix->ev_fork, st));
        AT_CUDA_OK(cudaStreamWaitEvent(ix->side, ix->ev_fork, 0));
    }
    if (!(skip & 1))
    k_tc_full<<<sms * 8, FULL_ROWS * FULL_SLICES, 0, fs>>>(x, l2norm_rows, ix->c, ix->cn, ix->k, rows->full, rows->tail_count,
                                            l32, labels64, kdist);
    AT_LAUNCH_OK();
    if (fork) AT_CUDA_OK(cudaEventRecord(ix->ev_join, ix->side));
    if (!(skip & 2))
    k_tc_tail<<<grid * RT * 4, 256, 0, st>>>(x, l2norm_rows, ix->c, ix->cn, rows->tail, rows->tail_count, nsuper, grid, l32,
                                            labels64, kdist, ix->tc_counters);
    AT_LAUNCH_OK();
    if (fork) AT_CUDA_OK(cudaStreamWaitEvent(st, ix->ev_join, 0));
    if (dist && exact_dist) {
        k_exact_dist<<<sms * 8, 256, 0, st>>>(x, n, l2norm_rows, ix->c, ix->cn, l32, l32 ? nullptr : labels64, dist);
        AT_LAUNCH_OK();
    }
    return AT_OK;
}

// Candidate re-check of the wide search: block q works through candidate queue q, a warp per row.  The warp stages the row
// and its (at most four) candidate centroids in shared memory with coalesced loads (a thread walking its own 2.5 KB row
// touches a new sector with every 16 bytes and the kernel becomes sector-bound: measured 0.22 ms), then lanes 0-3 evaluate
// one candidate each with the EXACT wide kernel's arithmetic (k_assign_gemm: |x|^2 and <x, c> each ONE sequential fp32 FMA
// chain over the d values, |c|^2 from the index, (|x|^2 + |c|^2) - 2 <x, c> clamped at 0); the lowest index wins exact ties.
constexpr int TAILW_WARPS = 8;
__global__ void __launch_bounds__(TAILW_WARPS * 32) k_tc_tail_wide(
    const float *__restrict__ x, int d, const float *__restrict__ c, const float *__restrict__ cn,
    const uint4 *__restrict__ tail_all, const unsigned int *__restrict__ tail_count, int64_t nsuper, int workers,
    int32_t *__restrict__ labels32, int64_t *__restrict__ labels64, float *__restrict__ dist,
    unsigned long long *__restrict__ counters) {
    extern __shared__ __align__(16) float s_stage[];   // per warp: 5 rows of d floats (the row, then the candidates)
    const int q = (int)blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned int n_cand = tail_count[2 + q];
    const uint4 *__restrict__ tail = tail_all + tc_queue_base(nsuper, workers, q / (RT * 4), q % (RT * 4));
    float *sx = s_stage + (size_t)w * 5 * d;
    for (unsigned int e = w; e < n_cand; e += TAILW_WARPS) {
        const uint4 cc = tail[e];
        const int64_t row = (int64_t)cc.x;
        const int cols[4] = {(int)(cc.y & 0xFFFFu), (int)(cc.y >> 16), (int)(cc.z & 0xFFFFu), (int)(cc.z >> 16)};
        __syncwarp();   // the previous entry's chains are done with the buffer
        for (int i = lane; i < d / 4; i += 32) {
            reinterpret_cast<float4 *>(sx)[i] = __ldg(reinterpret_cast<const float4 *>(x + row * d) + i);
#pragma unroll
            for (int t = 0; t < 4; t++)
                reinterpret_cast<float4 *>(sx + (size_t)(t + 1) * d)[i] = __ldg(reinterpret_cast<const float4 *>(c + (size_t)cols[t] * d) + i);
        }
        __syncwarp();
        float bd = INFINITY;
        int best = 0x7FFFFFFF;
        if (lane < 4) {
            const float *cr = sx + (size_t)(lane + 1) * d;
            float xn = 0.f, ip = 0.f;
            for (int i = 0; i < d; i += 4) {
                const float4 xv = *reinterpret_cast<const float4 *>(sx + i), cv = *reinterpret_cast<const float4 *>(cr + i);
                xn = fmaf(xv.x, xv.x, xn), xn = fmaf(xv.y, xv.y, xn), xn = fmaf(xv.z, xv.z, xn), xn = fmaf(xv.w, xv.w, xn);
                ip = fmaf(xv.x, cv.x, ip), ip = fmaf(xv.y, cv.y, ip), ip = fmaf(xv.z, cv.z, ip), ip = fmaf(xv.w, cv.w, ip);
            }
            best = cols[lane];
            bd = l2_expanded(xn, __ldg(cn + best), ip);
        }
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int ob = __shfl_xor_sync(0xffffffffu, best, o);
            if (od < bd || (od == bd && ob < best)) bd = od, best = ob;
        }
        if (lane == 0) {
            if (labels32) labels32[row] = best;
            if (labels64) labels64[row] = best;
            if (dist) dist[row] = bd;
        }
    }
    if (counters && threadIdx.x == 0 && n_cand) atomicAdd(&counters[0], (unsigned long long)n_cand);
}

// rows: a prepared image of exactly these rows (k-means), or nullptr to build one in the index's own workspace.  Labels only
// (callers that want exact distances of wide rows use the exact kernel).
int assign_tc_wide_search(at_index *ix, const float *x, int64_t n, int32_t *labels32, int64_t *labels64, float *dist,
                          at_tc_rows *rows, cudaStream_t st) {
    static bool configured[MAX_DEVICES] = {};   // the attribute is per device
    const int dev = current_device();
    if (!configured[dev]) {
        AT_CUDA_OK(cudaFuncSetAttribute(k_assign_tc_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCW_SMEM));
        configured[dev] = true;
    }
    if (!rows) {
        rows = &ix->rows;
        int rc = tc_rows_build(rows, x, n, 0, ix->tc_scale + SC_SX, ix->ext_shift ? ix->ext_shift : ix->shift, st, ix->d);
        if (rc != AT_OK) return rc;
    }
    const int64_t nsuper = (n + SROWS - 1) / SROWS;
    const int sms = sm_count() > 0 ? sm_count() : 1;
    int grid = sms;
    if (grid > nsuper) grid = (int)nsuper;
    if (grid < 1) grid = 1;
    AT_CUDA_OK(cudaMemsetAsync(rows->tail_count, 0, 2 * sizeof(unsigned int), st));
    k_assign_tc_wide<<<grid, TC_THREADS, TCW_SMEM, st>>>(reinterpret_cast<const unsigned char *>(rows->img), rows->erow, rows->xns,
                                                        n, reinterpret_cast<const unsigned char *>(ix->op), ix->ktiles, ix->k,
                                                        ix->d / 64, ix->tc_scale, 128u, labels32, labels64, dist, rows->tail,
                                                        rows->full, rows->tail_count);
    AT_LAUNCH_OK();
    // uncertified rows: the (at most four) candidate columns with the exact kernel's arithmetic, the rest (rare) through the
    // exact wide-row kernel over the list
    const size_t tail_smem = sizeof(float) * (size_t)TAILW_WARPS * 5 * (size_t)ix->d;   // <= 160 KB at d = 1024
    static bool tail_configured[MAX_DEVICES] = {};
    if (!tail_configured[dev]) {
        AT_CUDA_OK(cudaFuncSetAttribute(k_tc_tail_wide, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(float) * TAILW_WARPS * 5 * 1024)));
        tail_configured[dev] = true;
    }
    k_tc_tail_wide<<<grid * RT * 4, TAILW_WARPS * 32, tail_smem, st>>>(x, ix->d, ix->c, ix->cn, rows->tail, rows->tail_count, nsuper,
                                                                       grid, labels32, labels64, dist, ix->tc_counters);
    AT_LAUNCH_OK();
    return launch_assign_gemm_list(ix, x, rows->full, rows->tail_count + 1, n, rows->keys, labels32, labels64, dist, st);
}

}  // namespace at
