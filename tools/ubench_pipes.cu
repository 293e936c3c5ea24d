// Per-SM instruction throughput of the ops an accumulator scan can be built from (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes tools/ubench_pipes.cu
// 16 warps per SM (4 per scheduler), 16 independent dependency chains per thread: the figure printed is
// warp-instructions per clock per SM (4.0 = every scheduler issues one such instruction every clock).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 16
#define BODY(NAME, ASM_STMT)                                                                                  \
    __global__ void __launch_bounds__(512, 1) k_##NAME(int iters, uint32_t p0, uint32_t p1,                   \
                                                       unsigned long long *out, uint32_t *sink) {             \
        uint32_t r[CHAINS];                                                                                   \
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) r[i] = threadIdx.x * 2654435761u + i * 40503u + p1; \
        uint32_t a = p0, b = p1 ^ threadIdx.x;                                                                \
        __syncthreads();                                                                                      \
        unsigned long long t0 = clock64();                                                                    \
        for (int it = 0; it < iters; it++) {                                                                  \
            _Pragma("unroll") for (int u = 0; u < 4; u++) {                                                   \
                _Pragma("unroll") for (int i = 0; i < CHAINS; i++) { ASM_STMT; }                              \
            }                                                                                                 \
        }                                                                                                     \
        unsigned long long t1 = clock64();                                                                    \
        uint32_t acc = a ^ b;                                                                                 \
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) acc ^= r[i];                                       \
        if (acc == 0x12345u) sink[0] = acc;                                                                   \
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;                                                      \
    }

BODY(fmnmx, asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float *)&r[i]) : "f"(*(float *)&r[(i + 1) % CHAINS])))
BODY(fmnmx3, asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r[i]) : "f"(*(float *)&r[(i + 1) % CHAINS]), "f"(*(float *)&r[(i + 5) % CHAINS])))
BODY(lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]), "r"(a)))
BODY(lop3imm, asm volatile("lop3.b32 %0, %0, 0xFFFFFF80, %1, 0xEA;" : "+r"(r[i]) : "r"(a)))
BODY(imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b)))
BODY(imadimm, asm volatile("mad.lo.u32 %0, %0, %1, 77;" : "+r"(r[i]) : "r"(a)))
BODY(lea, asm volatile("{.reg .u32 t; shl.b32 t, %0, 7; add.u32 %0, t, %1;}" : "+r"(r[i]) : "r"(b)))
BODY(iadd, asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(umin, asm volatile("min.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(umin3, asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 5) % CHAINS])))
BODY(smin, asm volatile("min.s32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(fadd, asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float *)&r[i]) : "f"(*(float *)&r[(i + 1) % CHAINS])))
BODY(ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r[i]) : "f"(*(float *)&a), "f"(*(float *)&r[(i + 1) % CHAINS])))
BODY(hmnmx2, asm volatile("min.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(umin16x2, asm volatile("min.u16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(hadd2, asm volatile("add.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(prmt, asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(shf, asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
BODY(f2fp, asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "+r"(r[i]) : "f"(*(float *)&r[(i + 1) % CHAINS]), "f"(*(float *)&r[(i + 5) % CHAINS])))
BODY(setpsel, asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.b32 %0, %0, %1, p;}" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS])))
// mixes (2 instructions per statement)
BODY(mix_fmnmx_ffma, asm volatile("min.f32 %0, %0, %2; fma.rn.f32 %1, %1, %3, %2;" : "+f"(*(float *)&r[i]), "+f"(*(float *)&r[(i + 8) % CHAINS]) : "f"(*(float *)&b), "f"(*(float *)&a)))
BODY(mix_fmnmx_lop3, asm volatile("min.f32 %0, %0, %2; lop3.b32 %1, %1, 0xFFFFFF80, %3, 0xEA;" : "+f"(*(float *)&r[i]), "+r"(r[(i + 8) % CHAINS]) : "f"(*(float *)&b), "r"(a)))
BODY(mix_fmnmx_imad, asm volatile("min.f32 %0, %0, %2; mad.lo.u32 %1, %1, %3, 77;" : "+f"(*(float *)&r[i]), "+r"(r[(i + 8) % CHAINS]) : "f"(*(float *)&b), "r"(a)))
BODY(mix_fmnmx_fadd, asm volatile("min.f32 %0, %0, %2; add.f32 %1, %1, %3;" : "+f"(*(float *)&r[i]), "+f"(*(float *)&r[(i + 8) % CHAINS]) : "f"(*(float *)&b), "f"(*(float *)&a)))
BODY(mix_lop3_ffma, asm volatile("lop3.b32 %0, %0, 0xFFFFFF80, %2, 0xEA; fma.rn.f32 %1, %1, %3, %3;" : "+r"(r[i]), "+f"(*(float *)&r[(i + 8) % CHAINS]) : "r"(b), "f"(*(float *)&a)))
BODY(mix_umin_ffma, asm volatile("min.u32 %0, %0, %2; fma.rn.f32 %1, %1, %3, %3;" : "+r"(r[i]), "+f"(*(float *)&r[(i + 8) % CHAINS]) : "r"(b), "f"(*(float *)&a)))

typedef void (*kern_t)(int, uint32_t, uint32_t, unsigned long long *, uint32_t *);
struct Entry { const char *name; kern_t k; int per_stmt; };
#define E(NAME, N) {#NAME, k_##NAME, N}

int main() {
    unsigned long long *d_clk, h_clk[148];
    uint32_t *d_sink;
    cudaMalloc(&d_clk, 148 * 8);
    cudaMalloc(&d_sink, 4);
    Entry es[] = {E(fmnmx, 1), E(fmnmx3, 1), E(lop3, 1), E(lop3imm, 1), E(imad, 1), E(imadimm, 1), E(lea, 1), E(iadd, 1),
                  E(umin, 1), E(umin3, 1), E(smin, 1), E(fadd, 1), E(ffma, 1), E(hmnmx2, 1), E(umin16x2, 1),
                  E(hadd2, 1), E(prmt, 1), E(shf, 1), E(f2fp, 1), E(setpsel, 2), E(mix_fmnmx_ffma, 2), E(mix_fmnmx_lop3, 2),
                  E(mix_fmnmx_imad, 2), E(mix_fmnmx_fadd, 2), E(mix_lop3_ffma, 2), E(mix_umin_ffma, 2)};
    const int iters = 512;
    for (auto &e : es) {
        for (int threads : {128, 512}) {
            e.k<<<148, threads>>>(iters, 0x3f800001u, 0x40000003u, d_clk, d_sink);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("%s: error %s\n", e.name, cudaGetErrorString(err)); return 1; }
            cudaMemcpy(h_clk, d_clk, sizeof(h_clk), cudaMemcpyDeviceToHost);
            double clk = (double)h_clk[0];
            double stmts = (double)iters * 4 * CHAINS * (threads / 32);
            printf("%-16s warps/SM %2d: %.3f statements/clk/SM (%d instr per statement) -> %.2f clk per warp-statement per scheduler\n",
                   e.name, threads / 32, stmts / clk, e.per_stmt, clk / (stmts / 4));
        }
    }
    return 0;
}
