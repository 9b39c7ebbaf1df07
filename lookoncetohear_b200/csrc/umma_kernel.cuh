// Device side of umma_gemm: PTX wrappers (mbarrier, tensor-map TMA, tcgen05 / TMEM) and the kernel.  Included by
// umma_gemm.cu only.
#pragma once
#include "umma_ptx.cuh"

namespace l2h {
namespace umma {

// ---------------------------------------------------------------------------------------------------------------
// Shared-memory map (all regions 1024-byte aligned):
//   [fp32 staging ring: nstg x 32 KB][A operand ring: nop x planes x 16 KB][B: resident (all k-chunks) or ring of nop]
//   [epilogue transpose buffers: 4 warps x 4 KB]
// Tile schedule: CTA c owns column tile nt = c % n_tiles_n for its whole life (so a resident B is loaded once) and
// walks the row tiles gi, gi + groups, ... with gi = c / n_tiles_n; the n_tiles_n CTAs of a group read the same A tile
// at about the same time (L2 hits).
__global__ void __launch_bounds__(NTHREADS, 1)
umma_gemm_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bar_stg_full[MAX_NSTG], bar_stg_empty[MAX_NSTG];
    __shared__ __align__(8) unsigned long long bar_op_full[4], bar_op_empty[4];
    __shared__ __align__(8) unsigned long long bar_acc_full[2], bar_acc_empty[2];
    __shared__ __align__(8) unsigned long long bar_b_full;
    __shared__ unsigned tmem_base_s;
    __shared__ float ln_s[128];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // programmatic dependent launch: the successor may become resident now; THIS kernel's barrier / TMEM set-up and its
    // weight-slab loads run under the predecessor's tail, and only the roles that touch chain data (the TMA producer
    // before the first activation tile, the epilogue before residual reads / stores) wait for the predecessor
    griddep_launch();
    const unsigned smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms need 1024-byte alignment
    const int passes = p.passes, BN = p.BN, nop = p.nop, NSTG = p.nstg;
    const int planes_a = passes > 1 ? 2 : 1;          // passes: 1 = a_hi b_hi; 2 = + a_lo b_hi (bf16 weights, split activations);
    const int planes = passes > 2 ? 2 : 1;            //         3 = + a_hi b_lo (bf16x3: fp32-grade products).  `planes` = B planes
    const unsigned opA_bytes = (unsigned)planes_a * OPA_PLANE;
    const int nb64 = (BN + 63) >> 6;
    const unsigned opB_plane = p.b_mn_major ? (unsigned)nb64 * 8192u : (unsigned)BN * 128u;
    const unsigned opB_bytes = (unsigned)planes * opB_plane;
    const bool resident = p.b_resident != 0;
    const unsigned stg0 = smem0, opA0 = stg0 + (unsigned)NSTG * STG_BYTES, opB0 = opA0 + (unsigned)nop * opA_bytes;
    const unsigned epi0 = opB0 + (unsigned)(resident ? p.n_chunks : nop) * opB_bytes;

    if (tid == 0) {
        for (int i = 0; i < MAX_NSTG; ++i) { mbar_init(&bar_stg_full[i], 1); mbar_init(&bar_stg_empty[i], 4); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_op_full[i], resident ? 4 : 5); mbar_init(&bar_op_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 4); }
        mbar_init(&bar_b_full, 1);
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
    const int m_tiles = p_tiles * s_tiles;
    const int tile_rows = p.P_TILE * p.S_TILE;
    const int nt = blockIdx.x % p.n_tiles_n, gi = blockIdx.x / p.n_tiles_n, groups = gridDim.x / p.n_tiles_n;
    const int n0 = nt * BN;

    if (warp == 0) {
        // ===================== TMA producer ============================================================
        if (lane == 0) {
            const unsigned stg_tx = 2u * 128u * (unsigned)tile_rows;                       // two 32-float half boxes
            auto load_b = [&](unsigned dst, unsigned long long* bar, int j, int bz) {
                for (int pl = 0; pl < planes; ++pl) {
                    if (!p.b_mn_major) {
                        tma_load_4d(dst + pl * opB_plane, &p.tmB, bar, j * KC, n0, bz, pl);
                    } else {
                        for (int nb = 0; nb < nb64; ++nb)
                            tma_load_4d(dst + pl * opB_plane + nb * 8192, &p.tmB, bar, n0 + nb * 64, j * KC, bz, pl);
                    }
                }
            };
            if (resident) {                           // the whole [BN x K] weight slab, once
                mbar_expect_tx(&bar_b_full, opB_bytes * (unsigned)p.n_chunks);
                for (int j = 0; j < p.n_chunks; ++j) load_b(opB0 + j * opB_bytes, &bar_b_full, j, 0);
            }
            griddep_wait();                           // activations (and a batched B) come from the predecessor
            unsigned it = 0;
            for (int mt = gi; mt < m_tiles; mt += groups) {
                const int p0 = (mt % p_tiles) * p.P_TILE, seq0 = (mt / p_tiles) * p.S_TILE;
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
                    if (!resident) {
                        mbar_wait_to(&bar_op_empty[o], ((it / nop) & 1) ^ 1, 20 + o);
                        mbar_expect_tx(&bar_op_full[o], opB_bytes);
                        load_b(opB0 + o * opB_bytes, &bar_op_full[o], j, bz);
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
            const unsigned b_lbo = p.b_mn_major ? 8192u : 16u;
            if (resident) mbar_wait_to(&bar_b_full, 0, 25);
            for (int mt = gi; mt < m_tiles; mt += groups, ++tl) {
                const unsigned a = tl & 1;
                mbar_wait_to(&bar_acc_empty[a], ((tl >> 1) & 1) ^ 1, 30 + a);
                tc_fence_after();
                const unsigned d_tmem = tmem_base + a * (unsigned)BN;
                for (int j = 0; j < p.n_chunks; ++j, ++it) {
                    const int o = it % nop;
                    mbar_wait_to(&bar_op_full[o], (it / nop) & 1, 40 + o);
                    tc_fence_after();
                    const unsigned abase = opA0 + o * opA_bytes;
                    const unsigned bbase = opB0 + (unsigned)(resident ? j : o) * opB_bytes;
                    const unsigned long long da_hi = smem_desc(abase, 16, 1024);
                    const unsigned long long da_lo = smem_desc(abase + OPA_PLANE, 16, 1024);
                    const unsigned long long db_hi = smem_desc(bbase, b_lbo, 1024);
                    const unsigned long long db_lo = smem_desc(bbase + opB_plane, b_lbo, 1024);
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
        for (int mt = gi; mt < m_tiles; mt += groups) {
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
                const unsigned dst = opA0 + o * opA_bytes + (unsigned)r * 128u;
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
        // ===================== epilogue: TMEM -> registers -> (smem transpose) -> coalesced global rows ====
        // tcgen05.ld gives thread = row, 32 consecutive columns; stored like that every warp store would touch 32
        // different rows (16 B each).  Each warp therefore transposes its 32 x 32 block through 4 KB of shared memory
        // (128-byte rows, 16-byte chunks XOR-swizzled by row: conflict-free both ways) and stores with 8 lanes per row:
        // one warp instruction writes four full 128-byte lines.  Bias / PReLU / residual are applied after the
        // transpose, where a lane's columns are fixed and the residual is read with the same coalesced pattern.
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;
        const unsigned tb = epi0 + (unsigned)q * 4096u;
        const float slope1 = p.prelu ? __ldg(p.prelu) : 1.f;      // scalar PReLU slope (1 = identity)
        const bool has_prelu = p.prelu != nullptr || p.prelu_vec != nullptr;
        const int chunk = lane & 7, rsub = lane >> 3;
        griddep_wait();                               // C may still be read, R still be written by the predecessor
        unsigned tl = 0;
        for (int mt = gi; mt < m_tiles; mt += groups, ++tl) {
            const int p0 = (mt % p_tiles) * p.P_TILE, seq0 = (mt / p_tiles) * p.S_TILE;
            const int sl = r / p.P_TILE, pos = p0 + r % p.P_TILE, seq = seq0 + sl;
            const int valid = (r < tile_rows && pos < p.rows_per_seq && seq < p.nseq) ? 1 : 0;
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
                const int nl = cb + chunk * 4;         // this lane's four columns inside the tile (after the transpose)
                const int n = n0 + nl;
                const bool col_ok = nl < BN && n < p.N;
                const bool full4 = col_ok && (n + 3 < p.N) && p.vec_ok;
                // per-column epilogue operands of this lane (L1 hits after the first tile: same columns every tile)
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(slope1, slope1, slope1, slope1);
                if (full4) {
                    if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                    if (p.prelu_vec) s4 = __ldg(reinterpret_cast<const float4*>(p.prelu_vec + n));
                }
                __syncwarp();                          // tcgen05.ld is warp-collective; also: transpose buffer free again
                tc_ld16(taddr + cb, v);
                if (cb + 16 < BN) tc_ld16(taddr + cb + 16, v + 16);
                tc_wait_ld();
                if (cb + 32 >= BN) {                   // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_acc_empty[a]);
                }
#pragma unroll
                for (unsigned c = 0; c < 8; ++c)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(tb + (unsigned)lane * 128u + ((c ^ (unsigned)(lane & 7)) << 4)),
                                 "f"(v[c * 4]), "f"(v[c * 4 + 1]), "f"(v[c * 4 + 2]), "f"(v[c * 4 + 3]) : "memory");
                __syncwarp();
                if (__all_sync(0xffffffffu, full4 || !col_ok)) {
                    // ---- fast path: whole float4 groups.  Row offsets first (and the residual loads in flight), then math
                    long long co[8];
                    unsigned okm = 0;
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        const int row = itr * 4 + rsub;
                        co[itr] = __shfl_sync(0xffffffffu, coff, row) + n;
                        okm |= (unsigned)(__shfl_sync(0xffffffffu, valid, row) & (col_ok ? 1 : 0)) << itr;
                    }
                    float4 rr[8];
                    if (p.R) {
#pragma unroll
                        for (int itr = 0; itr < 8; ++itr)
                            rr[itr] = ((okm >> itr) & 1u) ? *reinterpret_cast<const float4*>(p.R + co[itr]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        const int row = itr * 4 + rsub;
                        float4 o;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                                     : "r"(tb + (unsigned)row * 128u + (((unsigned)chunk ^ (unsigned)(row & 7)) << 4)));
                        o.x = fmaf(o.x, p.alpha, b4.x); o.y = fmaf(o.y, p.alpha, b4.y); o.z = fmaf(o.z, p.alpha, b4.z); o.w = fmaf(o.w, p.alpha, b4.w);
                        if (has_prelu) {               // max(x,0) + slope * min(x,0)
                            o.x = fmaf(s4.x, fminf(o.x, 0.f), fmaxf(o.x, 0.f)); o.y = fmaf(s4.y, fminf(o.y, 0.f), fmaxf(o.y, 0.f));
                            o.z = fmaf(s4.z, fminf(o.z, 0.f), fmaxf(o.z, 0.f)); o.w = fmaf(s4.w, fminf(o.w, 0.f), fmaxf(o.w, 0.f));
                        }
                        if (p.R) { o.x += rr[itr].x; o.y += rr[itr].y; o.z += rr[itr].z; o.w += rr[itr].w; }
                        if ((okm >> itr) & 1u) *reinterpret_cast<float4*>(p.C + co[itr]) = o;
                    }
                } else {
                    // ---- cold path: ragged N or unaligned rows, element by element
#pragma unroll 1
                    for (int itr = 0; itr < 8; ++itr) {
                        const int row = itr * 4 + rsub;
                        float4 o;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                                     : "r"(tb + (unsigned)row * 128u + (((unsigned)chunk ^ (unsigned)(row & 7)) << 4)));
                        const long long cor = __shfl_sync(0xffffffffu, coff, row);
                        const int ok = __shfl_sync(0xffffffffu, valid, row);
                        if (!ok || !col_ok) continue;
                        const float ov[4] = {o.x, o.y, o.z, o.w};
                        for (int e = 0; e < 4 && n + e < p.N; ++e) {
                            float x = ov[e] * p.alpha;
                            if (p.bias) x += __ldg(p.bias + n + e);
                            if (p.prelu) x = prelu(x, slope1);
                            if (p.prelu_vec) x = prelu(x, __ldg(p.prelu_vec + n + e));
                            if (p.R) x += p.R[cor + n + e];
                            p.C[cor + n + e] = x;
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
