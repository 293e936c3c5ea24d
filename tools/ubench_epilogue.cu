// Micro-benchmarks behind the search kernel's epilogue design: TMEM read rate, alu / fma pipe rates
// and the cost of the key scan, per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_epilogue tools/ubench_epilogue.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ uint32_t umin3(uint32_t a, uint32_t b, uint32_t c) { return min(min(a, b), c); }

__device__ __forceinline__ void fold32(const uint32_t (&r)[32], const int cb, const uint32_t mul, uint32_t &t1, uint32_t &t2, uint32_t &t3) {
#pragma unroll
    for (int gp = 0; gp < 4; gp++) {
        const int e = 8 * gp;
        uint32_t kx[8];
#pragma unroll
        for (int i = 0; i < 8; i++) kx[i] = r[e + i] * mul + (uint32_t)(cb + e + i);
        const uint32_t a = min(umin3(kx[0], kx[1], kx[2]), kx[3]);
        const uint32_t b = min(umin3(kx[4], kx[5], kx[6]), kx[7]);
        const uint32_t lo = min(a, b), hi = max(a, b);
        t3 = umin3(t3, max(t2, lo), max(t1, hi));
        t2 = umin3(t2, hi, max(t1, lo));
        t1 = min(t1, lo);
    }
}

// mode 0: TMEM loads only (x32, each result xor-folded: 31 LOP3 -> dominated?) -- use minimal consumption: one xor of 2 regs
// mode 1: TMEM loads + fold32;  mode 2: fold32 on registers only (no TMEM);  mode 3: IMAD only; mode 4: VIMNMX3 only
__global__ void __launch_bounds__(512, 1) k_bench(int mode, int iters, uint32_t mul, int nwarps, unsigned long long *out_clk, uint32_t *sink) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    uint32_t acc = 0;
    unsigned long long t0 = 0, t1c = 0;
    if (warp < nwarps) {
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;
        uint32_t t1 = 0xffffffffu, t2 = 0xffffffffu, t3 = 0xffffffffu;
        uint32_t ra[32], rb[32];
#pragma unroll
        for (int i = 0; i < 32; i++) ra[i] = lane * 977 + i * 131 + warp, rb[i] = lane * 31 + i * 7 + warp;
        __syncwarp();
        t0 = clock64();
        if (mode == 0) {
            for (int it = 0; it < iters; it++) {
                tmem_ld32(ta, ra);
                tmem_ld32(ta + 32, rb);
                tmem_ld_wait();
                acc ^= ra[0] ^ rb[31];
                tmem_ld32(ta + 64, ra);
                tmem_ld32(ta + 96, rb);
                tmem_ld_wait();
                acc ^= ra[5] ^ rb[7];
            }
        } else if (mode == 1) {
            for (int it = 0; it < iters; it++) {
                tmem_ld32(ta, ra);
                tmem_ld32(ta + 32, rb);
                tmem_ld_wait();
                fold32(ra, 0, mul, t1, t2, t3);
                tmem_ld32(ta + 64, ra);
                fold32(rb, 32, mul, t1, t2, t3);
                tmem_ld32(ta + 96, rb);
                tmem_ld_wait();
                fold32(ra, 64, mul, t1, t2, t3);
                fold32(rb, 96, mul, t1, t2, t3);
            }
        } else if (mode == 2) {
            for (int it = 0; it < iters; it++) {
                fold32(ra, 0, mul, t1, t2, t3);
                fold32(rb, 32, mul, t1, t2, t3);
                ra[it & 31] += t1;   // keep the loop from being hoisted
                fold32(ra, 64, mul, t1, t2, t3);
                fold32(rb, 96, mul, t1, t2, t3);
                rb[it & 31] ^= t2;
            }
        } else if (mode == 3) {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int r4 = 0; r4 < 4; r4++)
#pragma unroll
                    for (int i = 0; i < 32; i++) ra[i] = ra[i] * mul + (uint32_t)(i + r4);
            }
        } else if (mode == 4) {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int r4 = 0; r4 < 4; r4++)
#pragma unroll
                    for (int i = 0; i < 32; i++) ra[i] = umin3(ra[i], rb[(i + 1) & 31], rb[(i + 7 + r4) & 31]) + 0;
#pragma unroll
                for (int i = 0; i < 32; i++) rb[i] ^= ra[(i + 3) & 31];
            }
        }
        t1c = clock64();
#pragma unroll
        for (int i = 0; i < 32; i++) acc ^= ra[i] ^ rb[i];
        acc ^= t1 ^ t2 ^ t3;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
    if (lane == 0 && warp < nwarps) out_clk[blockIdx.x * 16 + warp] = t1c - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    unsigned long long *d_clk, h_clk[16];
    uint32_t *d_sink;
    cudaMalloc(&d_clk, 148 * 16 * 8);
    cudaMalloc(&d_sink, 4);
    const char *names[] = {"tmem loads only (4 x x32 per iter = one 128x128 fp32 tile per 4 warps)", "tmem loads + key scan", "key scan on registers",
                           "128 IMAD per thread-iter", "128 VIMNMX3 (+32 LOP3) per thread-iter"};
    const int iters = 2000;
    for (int mode = 0; mode < 5; mode++) {
        for (int nw : {4, 8, 16}) {
            k_bench<<<148, 512>>>(mode, iters, 128u, nw, d_clk, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h_clk, d_clk, sizeof(h_clk), cudaMemcpyDeviceToHost);
            unsigned long long mx = 0;
            for (int w = 0; w < nw; w++) mx = h_clk[w] > mx ? h_clk[w] : mx;
            // per iteration each warp handles 32 rows x 128 columns = 4096 scores
            const double clk_iter = (double)mx / iters;
            printf("mode %d (%s) warps %2d: %.1f clk per warp-iteration; SM rate %.1f scores/clk\n", mode, names[mode], nw,
                   clk_iter, 4096.0 * nw / clk_iter);
        }
    }
    return 0;
}
