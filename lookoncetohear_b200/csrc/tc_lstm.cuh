// tc_lstm: the LSTM recurrence for MANY sequences on the tensor cores (throughput regime: offline batches,
// enrollment; the few-sequence latency regime keeps lstm.cuh's CUDA-core kernels).
//
// Per step and direction the recurrent product is gates^T[256 x NS] = W_hh[256 x 64] . h^T[64 x NS]: "swap-AB" --
// the 256 gate rows are the MMA's M dimension (two M = 128 tiles, W_hh bf16 hi/lo resident in shared memory for the
// CTA's whole life), the CTA's NS = 32 sequences are its N dimension, so one tcgen05.mma costs N/2 = 16 cycles
// instead of a full 128-row tile.  Products are bf16x3 split (W_hi h_hi + W_lo h_hi + W_hi h_lo; `passes` = 2 drops
// the W_lo term for the bf16 configuration), accumulators live in TENSOR MEMORY (2 x NS columns).
//
// 128 compute threads = TMEM lanes: lane r of tile m is gate column p = m*128 + r of the packed order (hidden unit
// j = p/4, gate q = p%4 in i,f,g,o).  A step:
//   1. every thread reads its 2 x NS accumulator values (tcgen05.ld), adds the precomputed input projection gx
//      (prefetched one step ahead into registers; coalesced: a warp reads 32 consecutive gate columns of one row)
//      and applies ITS gate's activation -- 2 x NS activations per thread, perfectly balanced over the four MUFUs;
//   2. the four gates of a (unit, sequence) meet through a padded shared-memory tile inside the warp (lanes 4u..4u+3):
//      lane q of a quad owns the sequences n = q (mod 4) and does their cell updates (c in registers);
//   3. h goes out as fp32 rows (the layer's output) and as bf16 hi/lo into the K-major SWIZZLE_128B h^T operand tile;
//      fence.proxy.async + mbarrier hand it to the MMA warp, which issues the next step's 24 MMAs and commits.
// Reference semantics: torch.nn.LSTM cell, gate order i,f,g,o (tfgridnet_causal.py:336-346, :512, :529).
#pragma once
#include "lstm.cuh"
#include "umma_ptx.cuh"

namespace l2h {
namespace tcl {

constexpr int NS = 32;                 // sequences per CTA (= MMA N)
constexpr int XLD = 136;               // padded row of the gate-exchange tile (words): conflict-free LDS.128 / STS.32
constexpr size_t W_BYTES = 2 * 256 * 128;                    // hi + lo planes, 256 rows x 128 B
constexpr size_t H_BYTES = 2 * NS * 128;                     // h^T hi + lo
constexpr size_t X_BYTES = (size_t)2 * NS * XLD * 4;         // [2 tiles][NS][XLD]
constexpr size_t SMEM = 1024 + W_BYTES + H_BYTES + X_BYTES;
constexpr int THREADS = 160;           // warps 0-3: compute (TMEM lanes), warp 4: MMA issuer + TMEM allocation

static __global__ void __launch_bounds__(THREADS, 2)
tc_lstm_kernel(const LstmArgs a, int passes) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bar_h, bar_d;
    __shared__ unsigned tmem_base_s;
    __shared__ long long gx_row[NS], out_row[NS], hc_off[NS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int dir = blockIdx.y, seq0 = blockIdx.x * NS;
    const unsigned sm0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned w_sm = sm0, h_sm = sm0 + (unsigned)W_BYTES;
    float* xt = reinterpret_cast<float*>(smem_raw + (sm0 - smem_u32(smem_raw)) + W_BYTES + H_BYTES);
    griddep_launch();
    if (tid == 0) { mbar_init(&bar_h, 128); mbar_init(&bar_d, 1); mbar_fence_init(); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2 * NS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // sequence -> first row of gx / out (in rows), state offset; -1: no such sequence
    const bool same_out = (a.out_outer_stride | a.out_inner_stride | a.out_step_stride) == 0;
    const long long o_outer = same_out ? a.outer_stride : a.out_outer_stride, o_inner = same_out ? a.inner_stride : a.out_inner_stride;
    const long long o_step = same_out ? a.step_stride : a.out_step_stride;
    if (tid < NS) {
        const int seq = seq0 + tid;
        if (seq < a.nseq) {
            const long long o = seq / a.inner_count, i = seq % a.inner_count;
            gx_row[tid] = o * a.outer_stride + i * a.inner_stride;
            out_row[tid] = o * o_outer + i * o_inner;
            hc_off[tid] = o * a.hc_outer_stride + i * 64;
        } else {
            gx_row[tid] = -1; out_row[tid] = -1; hc_off[tid] = -1;
        }
    }
    // W_hh of this direction: fp32 [256 (j*4+q)][64] -> bf16 hi/lo, K-major SWIZZLE_128B rows (weights: before the wait)
    if (tid < 128) {
        for (int row = tid; row < 256; row += 128) {
            const float4* src = reinterpret_cast<const float4*>(a.whh + ((size_t)dir * 256 + row) * 64);
            const unsigned dst = w_sm + (unsigned)row * 128u;
#pragma unroll
            for (unsigned c = 0; c < 8; ++c) {
                const float4 v0 = __ldg(src + 2 * c), v1 = __ldg(src + 2 * c + 1);
                const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                unsigned hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                    hi[e] = *reinterpret_cast<const unsigned*>(&h2);
                    const float2 hf = __bfloat1622float2(h2);
                    lo[e] = umma::pack_bf16x2(f[2 * e] - hf.x, f[2 * e + 1] - hf.y);
                }
                const unsigned off = dst + ((c ^ (unsigned)(row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + 256 * 128), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
            }
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const unsigned tmem = tmem_base_s;
    griddep_wait();

    if (warp == 4) {
        // ===================== MMA issuer =============================================================
        if (lane == 0) {
            const unsigned idesc = umma::make_idesc_bf16(NS, 0);
            for (int s = 0; s < a.L; ++s) {
                umma::mbar_wait_to(&bar_h, (unsigned)(s & 1), 80);
                umma::tc_fence_after();
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const unsigned long long w_hi = umma::smem_desc(w_sm + m * 16384, 16, 1024);
                    const unsigned long long w_lo = umma::smem_desc(w_sm + 256 * 128 + m * 16384, 16, 1024);
                    const unsigned long long h_hi = umma::smem_desc(h_sm, 16, 1024);
                    const unsigned long long h_lo = umma::smem_desc(h_sm + NS * 128, 16, 1024);
                    for (int ps = 0; ps < 3; ++ps) {
                        if (ps == 1 && passes < 3) continue;          // W_lo term only for the fp32-grade split
                        const unsigned long long da = (ps == 1) ? w_lo : w_hi, db = (ps == 2) ? h_lo : h_hi;
#pragma unroll
                        for (unsigned kk = 0; kk < 4; ++kk)
                            umma::tc_mma_bf16(tmem + m * NS, da + kk * 2, db + kk * 2, idesc, (ps | (int)kk) != 0);
                    }
                }
                umma::tc_commit(&bar_d);
            }
        }
    } else {
        // ===================== compute: activations, cell, h ===========================================
        const int r = tid, q = r & 3, u = r >> 2;                 // TMEM lane, gate, hidden unit inside the tile
        constexpr float LOG2E = 1.4426950408889634f;
        const float S0 = (q == 2) ? -2.f * LOG2E : -LOG2E, A0 = (q == 2) ? 2.f : 1.f, B0 = (q == 2) ? -1.f : 0.f;
        const int gcol = dir * 256 + r;                            // + m*128: this lane's gx column
        float c[2][NS / 4];
        // initial state
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int k = 0; k < NS / 4; ++k) {
                const int n = 4 * k + q, j = m * 32 + u;
                float h0 = 0.f, c0 = 0.f;
                if (a.h_state != nullptr && hc_off[n] >= 0) { h0 = a.h_state[hc_off[n] + j]; c0 = a.c_state[hc_off[n] + j]; }
                c[m][k] = c0;
                const __nv_bfloat16 hh = __float2bfloat16_rn(h0);
                const __nv_bfloat16 hl = __float2bfloat16_rn(h0 - __bfloat162float(hh));
                const unsigned off = h_sm + (unsigned)n * 128u + ((((unsigned)j >> 3) ^ (unsigned)(n & 7)) << 4) + ((unsigned)j & 7u) * 2u;
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(off), "h"(*reinterpret_cast<const unsigned short*>(&hh)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(off + NS * 128), "h"(*reinterpret_cast<const unsigned short*>(&hl)) : "memory");
            }
        fence_proxy_async();
        umma::mbar_arrive(&bar_h);
        // gx of the first step
        float gxn[2][NS];
        const long long sgn = dir ? -1 : 1;
        long long st = dir ? (long long)(a.L - 1) : 0;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            const long long row = gx_row[n];
#pragma unroll
            for (int m = 0; m < 2; ++m)
                gxn[m][n] = row >= 0 ? __ldg(a.gx + (row + st * a.step_stride) * a.gx_ld + gcol + m * 128) : 0.f;
        }
        for (int s = 0; s < a.L; ++s) {
            float pre[2][NS];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int n = 0; n < NS; ++n) pre[m][n] = gxn[m][n];
            if (s + 1 < a.L) {                                     // prefetch the next step's input projection
                const long long st1 = st + sgn;
#pragma unroll
                for (int n = 0; n < NS; ++n) {
                    const long long row = gx_row[n];
#pragma unroll
                    for (int m = 0; m < 2; ++m)
                        gxn[m][n] = row >= 0 ? __ldg(a.gx + (row + st1 * a.step_stride) * a.gx_ld + gcol + m * 128) : 0.f;
                }
            }
            umma::mbar_wait_to(&bar_d, (unsigned)(s & 1), 81);
            umma::tc_fence_after();
            {
                float d[NS];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const unsigned ta = tmem + ((unsigned)(warp * 32) << 16) + m * NS;
                    umma::tc_ld16(ta, d);
                    umma::tc_ld16(ta + 16, d + 16);
                    umma::tc_wait_ld();
#pragma unroll
                    for (int n = 0; n < NS; ++n) {
                        const float x = pre[m][n] + d[n];
                        const float v = __fdividef(A0, 1.f + ex2_ftz(S0 * x)) + B0;        // this lane's gate activation
                        xt[(m * NS + n) * XLD + r] = v;
                    }
                }
            }
            umma::tc_fence_before();
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int k = 0; k < NS / 4; ++k) {
                    const int n = 4 * k + q, j = m * 32 + u;
                    const float4 g = *reinterpret_cast<const float4*>(xt + (m * NS + n) * XLD + 4 * u);     // i, f, g, o
                    const float cc = g.y * c[m][k] + g.x * g.z;
                    c[m][k] = cc;
                    const float h = g.w * (__fdividef(2.f, 1.f + ex2_ftz(-2.f * LOG2E * cc)) - 1.f);
                    const long long orow = out_row[n];
                    if (orow >= 0) a.out[(orow + st * o_step) * a.out_ld + dir * 64 + j] = h;
                    const __nv_bfloat16 hh = __float2bfloat16_rn(h);
                    const __nv_bfloat16 hl = __float2bfloat16_rn(h - __bfloat162float(hh));
                    const unsigned off = h_sm + (unsigned)n * 128u + ((((unsigned)j >> 3) ^ (unsigned)(n & 7)) << 4) + ((unsigned)j & 7u) * 2u;
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(off), "h"(*reinterpret_cast<const unsigned short*>(&hh)) : "memory");
                    asm volatile("st.shared.b16 [%0], %1;" ::"r"(off + NS * 128), "h"(*reinterpret_cast<const unsigned short*>(&hl)) : "memory");
                    if (s + 1 == a.L && a.h_state != nullptr && hc_off[n] >= 0) {
                        a.h_state[hc_off[n] + j] = h;
                        a.c_state[hc_off[n] + j] = cc;
                    }
                }
            __syncwarp();                                          // the exchange tile is free for the next step
            if (s + 1 < a.L) {
                fence_proxy_async();
                umma::mbar_arrive(&bar_h);
            }
            st += sgn;
        }
    }
    umma::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        __syncwarp();
        umma::tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * NS) : "memory");
    }
}

}  // namespace tcl

// (static: every translation unit that includes this header owns its copy of the kernel and configures it itself)
static inline cudaError_t configure_tc_lstm() {
    return cudaFuncSetAttribute(tcl::tc_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcl::SMEM);
}
// many sequences: the recurrence on the tensor cores
static inline cudaError_t launch_tc_lstm(const LstmArgs& a, int passes, cudaStream_t st, bool pdl = false) {
    if (a.nseq <= 0 || a.L <= 0) return cudaErrorInvalidValue;
    dim3 grid((a.nseq + tcl::NS - 1) / tcl::NS, a.ndir);
    return launch_k(pdl, tcl::tc_lstm_kernel, grid, dim3(tcl::THREADS), tcl::SMEM, st, a, passes);
}

}  // namespace l2h
