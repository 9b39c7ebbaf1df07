// FFMA vs FFMA2 (fma.rn.f32x2) issue rate on one SM: W warps per scheduler, 8 independent accumulator chains per thread, operands in
// registers.  Prints cycles per warp-instruction per scheduler and FMA lanes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma_microbench tools/ffma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

template <int MODE>   // 0: FFMA, one operand shared; 1: FFMA2, all operands distinct registers; 2: FFMA2, one operand shared; 3: FFMA, all distinct
__global__ void k(float* out, int iters, long long* cyc) {
    float2 acc[8], w[8], g[8];
    for (int i = 0; i < 8; ++i) { acc[i] = make_float2(0.f, 0.f); w[i] = make_float2(1.f + threadIdx.x * 1e-6f + i, 0.5f + i); g[i] = make_float2(0.999f + i * 1e-5f, 1.001f - i * 1e-5f); }
    float2 h = make_float2(1.0001f, 0.9999f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) {
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(w[i].x), "f"(h.x));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(w[i].y), "f"(h.y));
                } else if (MODE == 1) {
                    acc[i] = ffma2(w[i], g[i], acc[i]);
                } else if (MODE == 2) {
                    acc[i] = ffma2(w[i], h, acc[i]);
                } else {
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(w[i].x), "f"(g[i].x));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(w[i].y), "f"(g[i].y));
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y + g[i].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
    const int iters = 2000;
    for (int warps_per_sched : {1, 2, 4}) {
        const int threads = 128 * warps_per_sched;
        for (int mode = 0; mode < 4; ++mode) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<1, threads>>>(out, iters, cyc);
                else if (mode == 1) k<1><<<1, threads>>>(out, iters, cyc);
                else if (mode == 2) k<2><<<1, threads>>>(out, iters, cyc);
                else k<3><<<1, threads>>>(out, iters, cyc);
                cudaDeviceSynchronize();
                cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            }
            const double instr_per_warp = (double)iters * 64 * ((mode == 0 || mode == 3) ? 2 : 1);
            const double cyc_per_instr_sched = (double)h / (instr_per_warp * warps_per_sched);
            const double fma_per_clk_sm = (double)iters * 128 * 32 * 4 * warps_per_sched / (double)h;
            printf("%d warp(s) per scheduler, %s: %.2f cycles per warp-instruction per scheduler, %.1f FMA per clock per SM\n", warps_per_sched,
                   mode == 0 ? "FFMA  (one shared operand)" : (mode == 1 ? "FFMA2 (distinct operands) " : (mode == 2 ? "FFMA2 (one shared operand)" : "FFMA  (distinct operands) ")), cyc_per_instr_sched, fma_per_clk_sm);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
