// Which pipe does an instruction share with VIMNMX3 (alu pipe)?  Each kernel interleaves one min3.u32 chain op with one op
// of another kind (independent chains); if the pair takes ~2 clk per scheduler the two run on different pipes, ~4 clk = same.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_mix tools/ubench_mix.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 8
#define BODY(NAME, ASM_STMT)                                                                                  \
    __global__ void __launch_bounds__(512, 1) k_##NAME(int iters, uint32_t p0, uint32_t p1,                   \
                                                       unsigned long long *out, uint32_t *sink) {             \
        uint32_t r[CHAINS], q[CHAINS];                                                                        \
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) {                                                  \
            r[i] = threadIdx.x * 2654435761u + i * 40503u + p1;                                               \
            q[i] = 0x3c003c00u + threadIdx.x + i;                                                             \
        }                                                                                                     \
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
        _Pragma("unroll") for (int i = 0; i < CHAINS; i++) acc ^= r[i] ^ q[i];                                \
        if (acc == 0x12345u) sink[0] = acc;                                                                   \
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;                                                      \
    }

#define MIN3 "{.reg .u32 t; min.u32 t, %0, %2; min.u32 %0, t, %3;}\n\t"
#define OPS : "+r"(r[i]), "+r"(q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "r"(q[(i + 1) % CHAINS]), "r"(a)
BODY(min3_only, asm volatile(MIN3 OPS))
BODY(min3_min3, asm volatile(MIN3 "{.reg .u32 t; min.u32 t, %1, %4; min.u32 %1, t, %5;}" OPS))
BODY(min3_hmnmx2, asm volatile(MIN3 "min.f16x2 %1, %1, %4;" OPS))
BODY(min3_hmnmx2_3, asm volatile(MIN3 "{.reg .b32 t; min.f16x2 t, %1, %4; min.f16x2 %1, t, %5;}" OPS))
BODY(min3_bmnmx2, asm volatile(MIN3 "min.bf16x2 %1, %1, %4;" OPS))
BODY(min3_umin16x2, asm volatile(MIN3 "min.u16x2 %1, %1, %4;" OPS))
BODY(min3_f2fp, asm volatile(MIN3 "cvt.rn.f16x2.f32 %1, %4, %5;" : "+r"(r[i]), "+r"(q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "f"(*(float *)&a)))
BODY(min3_f2bf, asm volatile(MIN3 "cvt.rn.bf16x2.f32 %1, %4, %5;" : "+r"(r[i]), "+r"(q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "f"(*(float *)&a)))
BODY(min3_hadd2, asm volatile(MIN3 "add.f16x2 %1, %1, %4;" OPS))
BODY(min3_hfma2, asm volatile(MIN3 "fma.rn.f16x2 %1, %1, %4, %5;" OPS))
BODY(min3_prmt, asm volatile(MIN3 "prmt.b32 %1, %1, %4, 0x7632;" OPS))
BODY(min3_lea, asm volatile(MIN3 "{.reg .u32 t; shl.b32 t, %1, 3; add.u32 %1, t, %4;}" OPS))
BODY(min3_iadd, asm volatile(MIN3 "add.u32 %1, %1, %4;" OPS))
BODY(min3_imad, asm volatile(MIN3 "mad.lo.u32 %1, %1, %5, %4;" OPS))
BODY(min3_fadd, asm volatile(MIN3 "add.f32 %1, %1, %4;" : "+r"(r[i]), "+f"(*(float *)&q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "r"(a)))
BODY(min3_ffmasat, asm volatile(MIN3 "fma.rn.sat.f32 %1, %1, %4, %4;" : "+r"(r[i]), "+f"(*(float *)&q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "r"(a)))
BODY(min3_2ffma, asm volatile(MIN3 "fma.rn.f32 %1, %1, %4, %4; fma.rn.f32 %1, %1, %4, %4;" : "+r"(r[i]), "+f"(*(float *)&q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "r"(a)))
BODY(min3_fmnmx, asm volatile(MIN3 "min.f32 %1, %1, %4;" : "+r"(r[i]), "+f"(*(float *)&q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "r"(a)))
BODY(min3_sel, asm volatile(MIN3 "{.reg .pred p; setp.lt.u32 p, %4, %5; selp.b32 %1, %1, %4, p;}" OPS))
BODY(min3_isetp, asm volatile(MIN3 "{.reg .pred p; setp.lt.u32 p, %4, %1; @p add.u32 %1, %1, 1;}" OPS))
BODY(min3_shf, asm volatile(MIN3 "shf.r.wrap.b32 %1, %1, %4, 3;" OPS))
BODY(min3_lop3, asm volatile(MIN3 "lop3.b32 %1, %1, %4, %5, 0xE8;" OPS))
BODY(min3_hmnmx2_f2fp, asm volatile(MIN3 "{.reg .b32 t; cvt.rn.f16x2.f32 t, %4, %5; min.f16x2 %1, %1, t;}" : "+r"(r[i]), "+r"(q[i]) : "r"(r[(i + 1) % CHAINS]), "r"(r[(i + 3) % CHAINS]), "f"(*(float *)&q[(i + 1) % CHAINS]), "f"(*(float *)&a)))

typedef void (*kern_t)(int, uint32_t, uint32_t, unsigned long long *, uint32_t *);
struct Entry { const char *name; kern_t k; };
#define E(NAME) {#NAME, k_##NAME}

int main() {
    unsigned long long *d_clk, h_clk[148];
    uint32_t *d_sink;
    cudaMalloc(&d_clk, 148 * 8);
    cudaMalloc(&d_sink, 4);
    Entry es[] = {E(min3_only), E(min3_min3), E(min3_hmnmx2), E(min3_hmnmx2_3), E(min3_bmnmx2), E(min3_umin16x2), E(min3_f2fp), E(min3_f2bf), E(min3_hadd2),
                  E(min3_hfma2), E(min3_prmt), E(min3_lea), E(min3_iadd), E(min3_imad), E(min3_fadd), E(min3_ffmasat),
                  E(min3_2ffma), E(min3_fmnmx), E(min3_sel), E(min3_isetp), E(min3_shf), E(min3_lop3), E(min3_hmnmx2_f2fp)};
    const int iters = 512;
    for (auto &e : es) {
        for (int threads : {512}) {
            e.k<<<148, threads>>>(iters, 0x3f800001u, 0x40000003u, d_clk, d_sink);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("%s: error %s\n", e.name, cudaGetErrorString(err)); return 1; }
            cudaMemcpy(h_clk, d_clk, sizeof(h_clk), cudaMemcpyDeviceToHost);
            double clk = (double)h_clk[0];
            double stmts = (double)iters * 4 * CHAINS * (threads / 32);
            printf("%-20s warps/SM %2d: %.2f clk per warp-statement per scheduler\n", e.name, threads / 32, clk / (stmts / 4));
        }
    }
    return 0;
}
