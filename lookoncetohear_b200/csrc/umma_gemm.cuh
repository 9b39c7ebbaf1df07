// umma_gemm: the dense contractions of both networks on the 5th-generation tensor cores.
//
//     C[m, n] = epi( alpha * sum_k A(m)[k] * B[n][k] + bias[n] )          (fp32 in, fp32 out)
//
// One persistent CTA per SM, 128-row tiles, the Blackwell way:
//   * operands arrive through TENSOR-MAP TMA (cp.async.bulk.tensor, SASS UTMALDG) with 128-byte swizzle:
//     the fp32 activation tile [128 rows x 64 k] into a staging ring, the weight tile (bf16, split on the host
//     side of the engine into hi/lo planes) straight into the operand ring;
//   * a converter warpgroup (thread = row) turns the staged fp32 rows into bf16 hi/lo operand tiles in the
//     canonical K-major SWIZZLE_128B layout (optionally applying LayerNorm over the 64 channels of the row first
//     -- nn.LayerNorm semantics -- so that LN -> Linear pairs are one kernel);
//   * ONE thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, M = 128, N = BN <= 256) with the accumulator in
//     TENSOR MEMORY.  passes = 3 gives fp32-grade products from three bf16 MMAs
//     (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, relative error ~2^-16 per product: the "bf16x3" split), passes = 1 is plain
//     bf16 (the offline bf16 configuration);
//   * the epilogue warpgroup reads the accumulator back with tcgen05.ld (thread = row = TMEM lane), applies
//     bias / PReLU / residual / scale and stores fp32 rows; accumulators are double-buffered in TMEM so the epilogue
//     of tile i overlaps the MMAs of tile i+1.
//
// "Rows" (the M space) are (sequence, position) pairs described by a 4-D tensor map (channel, position,
// seq_inner, seq_outer); a k-chunk of 64 channels may come from a position offset `dp` (overlapping windows: the
// enrollment net's unfold / ConvTranspose1d become plain GEMMs, out-of-range positions are zero-filled by TMA)
// and from one of two source tensors (concatenated K).  B is K-major [n][k] or MN-major [k][n] (attention P.V).
//
// Reference call sites this replaces (addmm / conv / bmm): tfgridnet_causal.py:510-516, :524-538, :547-551,
// :583-588; tfgridnet_orig/tfgridnet.py:117-125 (espnet2 GridNetBlock GEMMs and the T x T attention products).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include "common.cuh"

namespace l2h {
namespace umma {

constexpr int BM = 128;          // rows per tile == TMEM lanes
constexpr int KC = 64;           // k-chunk: 64 bf16 = one 128-byte swizzle row
constexpr int NTHREADS = 320;    // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: converter, warps 6-9: epilogue
constexpr int MAX_NSTG = 2;      // fp32 staging ring slots (32 KB each)
constexpr int STG_BYTES = BM * KC * 4;
constexpr int OPA_PLANE = BM * KC * 2;   // 16 KB per bf16 plane
constexpr int MAX_CHUNKS = 96;

struct KChunk {
    short c0;             // first channel of the chunk inside the source row
    signed char dp;       // position offset of the chunk (windows)
    unsigned char flags;  // bit0: source tensor 1, bit1: LayerNorm over the 64 channels
};

struct Params {
    CUtensorMap tmA0, tmA1, tmB;
    KChunk chunks[MAX_CHUNKS];
    int n_chunks;
    int P_TILE, S_TILE;        // positions x sequences per tile (P_TILE * S_TILE <= 128)
    int rows_per_seq, nseq;    // valid M space
    int seq_inner;             // sequence -> tensor-map coordinates (seq % seq_inner, seq / seq_inner)
    int pos_bias;              // added to every position coordinate (may be negative: zero-filled halo)
    int N, BN, n_tiles_n, passes, b_mn_major, b_by_seq;
    unsigned idesc;
    int nop;                   // operand ring slots
    int nstg;                  // fp32 staging ring slots (1 or 2)
    int tmem_cols;
    int vec_ok;                // C/R rows 16-byte aligned: float4 epilogue accesses
    // epilogue
    float* C;
    const float* R;
    long long ldc, c_seq_stride, c_inner_stride;
    int c_inner;
    const float* bias;
    const float* prelu;
    const float* prelu_vec;
    const float* ln_g;
    const float* ln_b;
    float alpha;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
L2H_DEVINL void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug must not hang the GPU (a hung box is a lost lease) -- trap after ~2 s
L2H_DEVINL void mbar_wait_to(unsigned long long* bar, unsigned parity, int id) {
    const unsigned addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) {
            printf("umma_gemm: barrier %d timed out (block %d thread %d parity %u)\n", id, blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
L2H_DEVINL void tma_load_4d(unsigned dst, const CUtensorMap* tm, unsigned long long* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
L2H_DEVINL void tmap_prefetch(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
L2H_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
L2H_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
L2H_DEVINL void tc_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
L2H_DEVINL void tc_mma_bf16(unsigned d_tmem, unsigned long long da, unsigned long long db, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
L2H_DEVINL void tc_ld16(unsigned taddr, float* v) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
L2H_DEVINL void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (sm_100): start address, leading/stride byte offsets (>>4), version 1,
// SWIZZLE_128B.  K-major operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1).
// MN-major operand tile: k rows of 128 B (64 bf16 along N), 8-row groups 1024 B apart (SBO), the next 64
// columns `lbo_bytes` further (LBO).
L2H_DEVINL unsigned long long smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFFu) >> 4);
    d |= (unsigned long long)(lbo_bytes >> 4) << 16;
    d |= (unsigned long long)(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

__host__ __device__ inline unsigned make_idesc_bf16(int n, int b_mn_major) {
    // c_format F32 (bits 4-5 = 1), a/b format BF16 (bits 7-9, 10-12 = 1), b_major bit 16, N>>3 at 17, M>>4 at 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(b_mn_major ? 1 : 0) << 16) | ((unsigned)(n >> 3) << 17) |
           ((unsigned)(BM >> 4) << 24);
}

L2H_DEVINL unsigned pack_bf16x2(float lo_elem, float hi_elem) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<const unsigned*>(&h);
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
umma_gemm_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bar_stg_full[MAX_NSTG], bar_stg_empty[MAX_NSTG];
    __shared__ __align__(8) unsigned long long bar_op_full[4], bar_op_empty[4];
    __shared__ __align__(8) unsigned long long bar_acc_full[2], bar_acc_empty[2];
    __shared__ unsigned tmem_base_s;
    __shared__ float ln_s[128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms need 1024-byte alignment
    const int passes = p.passes, BN = p.BN, nop = p.nop, NSTG = p.nstg;
    const unsigned opA_bytes = (passes > 1 ? 2 : 1) * OPA_PLANE;
    const int nb64 = (BN + 63) >> 6;
    const unsigned opB_plane = p.b_mn_major ? (unsigned)nb64 * 8192u : (unsigned)BN * 128u;
    const unsigned op_bytes = opA_bytes + (passes > 1 ? 2 : 1) * opB_plane;
    const unsigned stg0 = smem0, op0 = smem0 + NSTG * STG_BYTES;

    if (tid == 0) {
        for (int i = 0; i < MAX_NSTG; ++i) { mbar_init(&bar_stg_full[i], 1); mbar_init(&bar_stg_empty[i], 4); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_op_full[i], 5); mbar_init(&bar_op_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 4); }
        mbar_fence_init();
        tmap_prefetch(&p.tmA0); tmap_prefetch(&p.tmA1); tmap_prefetch(&p.tmB);
    }
    if (tid < 128 && p.ln_g != nullptr) ln_s[tid] = tid < 64 ? __ldg(p.ln_g + tid) : __ldg(p.ln_b + tid - 64);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = tmem_base_s;

    const int p_tiles = (p.rows_per_seq + p.P_TILE - 1) / p.P_TILE;
    const int s_tiles = (p.nseq + p.S_TILE - 1) / p.S_TILE;
    const int n_tiles = p_tiles * s_tiles * p.n_tiles_n;
    const int tile_rows = p.P_TILE * p.S_TILE;

    if (warp == 0) {
        // ===================== TMA producer ============================================================
        if (lane == 0) {
            const unsigned stg_tx = 2u * 128u * (unsigned)tile_rows;                       // two 32-float half boxes
            const unsigned b_tx = (passes > 1 ? 2u : 1u) * (p.b_mn_major ? (unsigned)nb64 * 8192u : (unsigned)BN * 128u);
            unsigned it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
                const int p0 = (mt % p_tiles) * p.P_TILE, seq0 = (mt / p_tiles) * p.S_TILE;
                const int n0 = nt * BN;
                const int s_in = seq0 % p.seq_inner, s_out = seq0 / p.seq_inner;
                const int bz = p.b_by_seq ? seq0 : 0;
                for (int j = 0; j < p.n_chunks; ++j, ++it) {
                    const KChunk kc = p.chunks[j];
                    const int s = it % NSTG, o = it % nop;
                    mbar_wait_to(&bar_stg_empty[s], ((it / NSTG) & 1) ^ 1, 10 + s);
                    mbar_expect_tx(&bar_stg_full[s], stg_tx);
                    const CUtensorMap* tm = (kc.flags & 1) ? &p.tmA1 : &p.tmA0;
                    const unsigned dst = stg0 + s * STG_BYTES;
                    tma_load_4d(dst, tm, &bar_stg_full[s], kc.c0, p0 + kc.dp + p.pos_bias, s_in, s_out);
                    tma_load_4d(dst + STG_BYTES / 2, tm, &bar_stg_full[s], kc.c0 + 32, p0 + kc.dp + p.pos_bias, s_in, s_out);
                    mbar_wait_to(&bar_op_empty[o], ((it / nop) & 1) ^ 1, 20 + o);
                    mbar_expect_tx(&bar_op_full[o], b_tx);
                    const unsigned bdst = op0 + o * op_bytes + opA_bytes;
                    for (int pl = 0; pl < (passes > 1 ? 2 : 1); ++pl) {
                        if (!p.b_mn_major) {
                            tma_load_4d(bdst + pl * opB_plane, &p.tmB, &bar_op_full[o], j * KC, n0, bz, pl);
                        } else {
                            for (int nb = 0; nb < nb64; ++nb)
                                tma_load_4d(bdst + pl * opB_plane + nb * 8192, &p.tmB, &bar_op_full[o], n0 + nb * 64, j * KC, bz, pl);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer ==============================================================
        if (lane == 0) {
            unsigned it = 0, tl = 0;
            const unsigned kstep_a = 2;                                   // 16 bf16 = 32 B along K (>>4)
            const unsigned kstep_b = p.b_mn_major ? 128u : 2u;            // 16 k rows = 2048 B (>>4) when MN-major
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
                const unsigned a = tl & 1;
                mbar_wait_to(&bar_acc_empty[a], ((tl >> 1) & 1) ^ 1, 30 + a);
                tc_fence_after();
                const unsigned d_tmem = tmem_base + a * (unsigned)BN;
                for (int j = 0; j < p.n_chunks; ++j, ++it) {
                    const int o = it % nop;
                    mbar_wait_to(&bar_op_full[o], (it / nop) & 1, 40 + o);
                    tc_fence_after();
                    const unsigned abase = op0 + o * op_bytes, bbase = abase + opA_bytes;
                    const unsigned long long da_hi = smem_desc(abase, 16, 1024);
                    const unsigned long long da_lo = smem_desc(abase + OPA_PLANE, 16, 1024);
                    const unsigned long long db_hi = p.b_mn_major ? smem_desc(bbase, 8192, 1024) : smem_desc(bbase, 16, 1024);
                    const unsigned long long db_lo = p.b_mn_major ? smem_desc(bbase + opB_plane, 8192, 1024) : smem_desc(bbase + opB_plane, 16, 1024);
                    for (int ps = 0; ps < passes; ++ps) {
                        const unsigned long long da = (ps == 1) ? da_lo : da_hi;
                        const unsigned long long db = (ps == 2) ? db_lo : db_hi;
#pragma unroll
                        for (unsigned kk = 0; kk < 4; ++kk)
                            tc_mma_bf16(d_tmem, da + kk * kstep_a, db + kk * kstep_b, p.idesc, (j | ps | (int)kk) != 0);
                    }
                    tc_commit(&bar_op_empty[o]);          // frees the operand slot when these MMAs have read it
                }
                tc_commit(&bar_acc_full[a]);              // accumulator complete -> epilogue
            }
        }
    } else if (warp < 6) {
        // ===================== converter: fp32 staging -> bf16 hi/lo operand tile (thread = row) =======
        const int r = (warp - 2) * 32 + lane;
        const unsigned sw = (unsigned)(r & 7);
        unsigned it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int j = 0; j < p.n_chunks; ++j, ++it) {
                const int s = it % NSTG, o = it % nop;
                mbar_wait_to(&bar_stg_full[s], (it / NSTG) & 1, 50 + s);
                float v[64];
                const unsigned src = stg0 + s * STG_BYTES + (unsigned)r * 128u;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (unsigned c = 0; c < 8; ++c) {
                        float4 t;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                     : "r"(src + h * (STG_BYTES / 2) + ((c ^ sw) << 4)));
                        v[h * 32 + c * 4 + 0] = t.x; v[h * 32 + c * 4 + 1] = t.y;
                        v[h * 32 + c * 4 + 2] = t.z; v[h * 32 + c * 4 + 3] = t.w;
                    }
                if (p.chunks[j].flags & 2) {        // LayerNorm over the 64 channels of this row (two-pass, biased var)
                    float sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) sum += v[i];
                    const float mu = sum * (1.f / 64.f);
                    float q = 0.f;
#pragma unroll
                    for (int i = 0; i < 64; ++i) { v[i] -= mu; q = fmaf(v[i], v[i], q); }
                    const float rs = rsqrtf(q * (1.f / 64.f) + 1e-5f);
#pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] = fmaf(v[i] * rs, ln_s[i], ln_s[64 + i]);
                }
                mbar_wait_to(&bar_op_empty[o], ((it / nop) & 1) ^ 1, 60 + o);
                const unsigned dst = op0 + o * op_bytes + (unsigned)r * 128u;
#pragma unroll
                for (unsigned c = 0; c < 8; ++c) {
                    unsigned hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = v[c * 8 + 2 * e], a1 = v[c * 8 + 2 * e + 1];
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);
                        hi[e] = *reinterpret_cast<const unsigned*>(&h2);
                        const float2 hf = __bfloat1622float2(h2);
                        lo[e] = pack_bf16x2(a0 - hf.x, a1 - hf.y);
                    }
                    const unsigned off = dst + ((c ^ sw) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
                    if (passes > 1)
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(off + OPA_PLANE), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
                }
                fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) { mbar_arrive(&bar_op_full[o]); mbar_arrive(&bar_stg_empty[s]); }
            }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> global (thread = row = TMEM lane) ========
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;
        const float slope = p.prelu ? __ldg(p.prelu) : 0.f;
        unsigned tl = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
            const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
            const int p0 = (mt % p_tiles) * p.P_TILE, seq0 = (mt / p_tiles) * p.S_TILE;
            const int n0 = nt * BN;
            const int sl = r / p.P_TILE, pos = p0 + r % p.P_TILE, seq = seq0 + sl;
            const bool valid = r < tile_rows && pos < p.rows_per_seq && seq < p.nseq;
            long long coff;
            if (p.c_inner > 1)
                coff = (long long)(seq / p.c_inner) * p.c_seq_stride + (long long)(seq % p.c_inner) * p.c_inner_stride + (long long)pos * p.ldc;
            else
                coff = (long long)seq * p.c_seq_stride + (long long)pos * p.ldc;
            const unsigned a = tl & 1;
            mbar_wait_to(&bar_acc_full[a], (tl >> 1) & 1, 70 + a);
            tc_fence_after();
            const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + a * (unsigned)BN;
            for (int cb = 0; cb < BN; cb += 32) {
                float v[32];
                __syncwarp();                          // tcgen05.ld is warp-collective (.sync.aligned)
                tc_ld16(taddr + cb, v);
                if (cb + 16 < BN) tc_ld16(taddr + cb + 16, v + 16);
                tc_wait_ld();
                if (cb + 32 >= BN) {                   // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_acc_empty[a]);
                }
                if (!valid) continue;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int n = n0 + cb + g * 4;
                    if (cb + g * 4 >= BN || n >= p.N) break;
                    float4 o = make_float4(v[g * 4] * p.alpha, v[g * 4 + 1] * p.alpha, v[g * 4 + 2] * p.alpha, v[g * 4 + 3] * p.alpha);
                    if (n + 3 < p.N && p.vec_ok) {
                        if (p.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
                        if (p.prelu) { o.x = prelu(o.x, slope); o.y = prelu(o.y, slope); o.z = prelu(o.z, slope); o.w = prelu(o.w, slope); }
                        if (p.prelu_vec) { const float4 sv = __ldg(reinterpret_cast<const float4*>(p.prelu_vec + n)); o.x = prelu(o.x, sv.x); o.y = prelu(o.y, sv.y); o.z = prelu(o.z, sv.z); o.w = prelu(o.w, sv.w); }
                        if (p.R) { const float4 rr = *reinterpret_cast<const float4*>(p.R + coff + n); o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
                        *reinterpret_cast<float4*>(p.C + coff + n) = o;
                    } else {
                        const float ov[4] = {o.x, o.y, o.z, o.w};
                        for (int e = 0; e < 4 && n + e < p.N; ++e) {
                            float x = ov[e];
                            if (p.bias) x += __ldg(p.bias + n + e);
                            if (p.prelu) x = prelu(x, slope);
                            if (p.prelu_vec) x = prelu(x, __ldg(p.prelu_vec + n + e));
                            if (p.R) x += p.R[coff + n + e];
                            p.C[coff + n + e] = x;
                        }
                    }
                }
            }
        }
    }
    // ---- teardown ------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace umma
}  // namespace l2h
