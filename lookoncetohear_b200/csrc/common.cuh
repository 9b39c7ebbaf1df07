// Shared device helpers for the lookonce-b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace l2h {

#define L2H_DEVINL __device__ __forceinline__

// ---- packed fp32x2 FMA (Blackwell FFMA2): d.xy = a.xy * b.xy + c.xy -------------------------
L2H_DEVINL float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)),
          "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

L2H_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
L2H_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (red must hold 32 floats). All threads get the result.
L2H_DEVINL float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect red from a previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (lane < nw) ? red[lane] : 0.f;
    t = warp_sum(t);
    return t;
}

// sigmoid / tanh through ex2.approx + rcp.approx (abs err ~1e-7; MUFU.TANH's 5e-4 is too coarse
// for the 1e-3 rel-L2 gate across 97 recurrent steps x 3 blocks).
L2H_DEVINL float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
L2H_DEVINL float fast_tanh(float x) { return 2.f * __fdividef(1.f, 1.f + __expf(-2.f * x)) - 1.f; }

// 2^x, flush-to-zero (plain ex2.approx carries extra denormal-range fix-up instructions)
L2H_DEVINL float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

L2H_DEVINL float prelu(float x, float a) { return x >= 0.f ? x : a * x; }

// cp.async 16 B (LDGSTS)
L2H_DEVINL void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
L2H_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
L2H_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization
// attribute may start while its predecessor drains; griddep_wait() blocks until the predecessor
// grid has completed and flushed.  Every kernel of a chain calls griddep_launch() first (lets the
// successor's CTAs become resident early and run their weight-only prologue) and griddep_wait()
// before touching anything an earlier kernel wrote -- on every path, so completion stays
// transitive.  Both are no-ops for launches without the attribute.
L2H_DEVINL void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
L2H_DEVINL void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace l2h

#include <utility>
namespace l2h {
// host: launch with (or without) the PDL attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace l2h
