// PTX wrappers shared by the tcgen05 kernels (umma_gemm_kernel, tc_lstm_kernel): mbarrier, tensor-map TMA, tcgen05 / TMEM,
// shared-memory matrix descriptors.  Device-inline only: safe to include from every translation unit.
#pragma once
#include <cstdio>
#include "umma_gemm.cuh"

namespace l2h {
namespace umma {

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

L2H_DEVINL unsigned pack_bf16x2(float lo_elem, float hi_elem) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo_elem, hi_elem);
    return *reinterpret_cast<const unsigned*>(&h);
}

}  // namespace umma
}  // namespace l2h
