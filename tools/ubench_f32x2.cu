// Packed fp32x2 arithmetic (fma.rn.f32x2 / add.f32x2 / mul.f32x2, sm_100+) against scalar FFMA / FADD: issue slots per
// warp-instruction and per-lane flop rate.  16 independent chains per thread, 4 or 16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_f32x2 tools/ubench_f32x2.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 16
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float p0, float p1, unsigned long long *out, float *sink) {
    float r[2 * CHAINS];
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) r[i] = p0 * (float)(threadIdx.x + i) + p1;
    float a = p0, b = p1;
    unsigned long long *r2 = reinterpret_cast<unsigned long long *>(r);
    unsigned long long a2, b2;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a2) : "f"(a), "f"(b));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b2) : "f"(b), "f"(a));
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(a), "f"(b));
                if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r2[i]) : "l"(a2), "l"(b2));
                if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r2[i]) : "l"(a2));
                if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(r2[i]) : "l"(a2));
                if (MODE == 4) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(a));
                if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r2[i]) : "l"(r2[(i + 1) % CHAINS]));
                if (MODE == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(r[(i + 3) % CHAINS]), "f"(r[(i + 7) % CHAINS]));
                if (MODE == 7) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r2[i]) : "l"(r2[(i + 3) % CHAINS]), "l"(r2[(i + 7) % CHAINS]));
            }
        }
    }
    unsigned long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * CHAINS; i++) acc += r[i];
    if (acc == 12345.678f) sink[0] = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int flops_per_lane_stmt, unsigned long long *d_clk, float *d_sink) {
    unsigned long long h[148];
    const int iters = 512;
    for (int threads : {128, 512}) {
        k<MODE><<<148, threads>>>(iters, 1.0000001f, 0.9999999f, d_clk, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
        double stmts = (double)iters * 4 * CHAINS * (threads / 32);
        double per_clk = stmts / (double)h[0];
        printf("%-28s warps/SM %2d: %.3f warp-instr/clk/SM = %.2f clk per warp-instr per scheduler; %.1f lane-flops/clk/SM\n", name,
               threads / 32, per_clk, 4.0 / per_clk, per_clk * 32 * flops_per_lane_stmt);
    }
}

int main() {
    unsigned long long *d_clk;
    float *d_sink;
    cudaMalloc(&d_clk, 148 * 8);
    cudaMalloc(&d_sink, 4);
    run<0>("ffma (uniform operands)", 2, d_clk, d_sink);
    run<6>("ffma (3 registers)", 2, d_clk, d_sink);
    run<1>("ffma2 (uniform operands)", 4, d_clk, d_sink);
    run<7>("ffma2 (3 register pairs)", 4, d_clk, d_sink);
    run<4>("fadd", 1, d_clk, d_sink);
    run<2>("fadd2 (uniform operand)", 2, d_clk, d_sink);
    run<5>("fadd2 (2 register pairs)", 2, d_clk, d_sink);
    run<3>("fmul2", 2, d_clk, d_sink);
    return 0;
}
