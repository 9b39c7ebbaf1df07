// rows_gemm: C[m, n] = epi( sum_k A(m)[k] * Wt[k][n] + bias[n] ),  fp32 on the CUDA cores with
// packed FFMA2.  "Rows" are (stream, frame, freq-bin) activations: M is huge or tiny, K and N are
// small (64..512), so W^T lives in shared memory and A streams through once.
//
// A rows may be overlapping windows of a [seq][pos][lda] tensor (the enrollment net's unfold /
// ConvTranspose1d become plain GEMMs this way): row m -> seq = m / a_rows_per_seq,
// p = m % a_rows_per_seq, address A + seq*a_seq_stride + p*lda, K contiguous floats.
//
// Optional prologue: LayerNorm over the K==64 channels of each row (nn.LayerNorm semantics:
// biased variance, eps inside the sqrt).  Epilogues: bias | bias+PReLU | bias+residual.
#pragma once
#include "common.cuh"

namespace l2h {

struct GemmArgs {
    const float* A;
    int64_t lda;
    int a_rows_per_seq;      // 0 => plain row-major
    int64_t a_seq_stride;
    const float* Wt;         // [K][N]
    const float* bias;       // [N] or null
    float* C;
    int64_t ldc;
    int c_rows_per_seq;      // 0 => plain; else row m -> (seq = m / c_rows_per_seq, p = m % c_rows_per_seq)
    int64_t c_seq_stride;    //   offset = (seq / c_inner)*c_seq_stride + (seq % c_inner)*c_inner_stride + p*ldc
    int c_inner;             //   (c_inner <= 1 => offset = seq*c_seq_stride + p*ldc)
    int64_t c_inner_stride;
    const float* R;          // residual, indexed like C (may alias C); null => none
    const float* ln_g;       // LN prologue (requires K == 64); null => none
    const float* ln_b;
    const float* prelu;      // scalar slope pointer; null => none
    const float* prelu_vec;  // per-output-column slopes [N]; null => none
    int M, N, K;
};

constexpr int GK = 64;  // K tile

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
rows_gemm_kernel(const GemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int APAD = 4;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                      // [GK][BM + APAD]  (transposed: k-major)
    float* Bs = smem + GK * (BM + APAD);   // [GK][BN]

    griddep_launch();
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);

    float2 acc[TM][TN / 2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN / 2; ++j) acc[i][j] = make_float2(0.f, 0.f);

    griddep_wait();
    for (int k0 = 0; k0 < g.K; k0 += GK) {
        // ---- A tile: 16 lanes per row, one float4 each (64 floats / row) ------------------
        for (int r = tid / 16; r < BM; r += NT / 16) {
            const int m = m0 + r;
            const int c4 = tid % 16;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < g.M) {
                const float* ap;
                if (g.a_rows_per_seq > 0) {
                    const int seq = m / g.a_rows_per_seq, p = m % g.a_rows_per_seq;
                    ap = g.A + (int64_t)seq * g.a_seq_stride + (int64_t)p * g.lda;
                } else {
                    ap = g.A + (int64_t)m * g.lda;
                }
                v = *reinterpret_cast<const float4*>(ap + k0 + c4 * 4);
            }
            if (g.ln_g != nullptr) {   // LayerNorm over the 64 channels held by these 16 lanes
                float s = v.x + v.y + v.z + v.w;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mu = s * (1.f / 64.f);
                const float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
                float q = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
                const float rs = rsqrtf(q * (1.f / 64.f) + 1e-5f);
                const float4 gg = *reinterpret_cast<const float4*>(g.ln_g + c4 * 4);
                const float4 bb = *reinterpret_cast<const float4*>(g.ln_b + c4 * 4);
                v.x = dx * rs * gg.x + bb.x;
                v.y = dy * rs * gg.y + bb.y;
                v.z = dz * rs * gg.z + bb.z;
                v.w = dw * rs * gg.w + bb.w;
            }
            As[(c4 * 4 + 0) * (BM + APAD) + r] = v.x;
            As[(c4 * 4 + 1) * (BM + APAD) + r] = v.y;
            As[(c4 * 4 + 2) * (BM + APAD) + r] = v.z;
            As[(c4 * 4 + 3) * (BM + APAD) + r] = v.w;
        }
        // ---- B tile: Wt[k0 .. k0+63][n0 .. n0+BN) -----------------------------------------
        for (int i = tid; i < GK * (BN / 4); i += NT) {
            const int kk = i / (BN / 4), c4 = i % (BN / 4);
            *reinterpret_cast<float4*>(Bs + kk * BN + c4 * 4) =
                *reinterpret_cast<const float4*>(g.Wt + (int64_t)(k0 + kk) * g.N + n0 + c4 * 4);
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < GK; ++kk) {
            float a[TM];
            float2 b[TN / 2];
            if constexpr (TM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(As + kk * (BM + APAD) + ty * TM + i);
                    a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; i += 2) {
                    const float2 t = *reinterpret_cast<const float2*>(As + kk * (BM + APAD) + ty * TM + i);
                    a[i] = t.x; a[i + 1] = t.y;
                }
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(Bs + kk * BN + tx * TN + j);
                b[j / 2] = make_float2(t.x, t.y);
                b[j / 2 + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < TN / 2; ++j) acc[i][j] = ffma2(aa, b[j], acc[i][j]);
            }
        }
        __syncthreads();
    }

    // ---- epilogue -----------------------------------------------------------------------
    const float slope = g.prelu ? __ldg(g.prelu) : 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
        int64_t coff;
        if (g.c_rows_per_seq > 0) {
            const int seq = m / g.c_rows_per_seq, p = m % g.c_rows_per_seq;
            if (g.c_inner > 1)
                coff = (int64_t)(seq / g.c_inner) * g.c_seq_stride + (int64_t)(seq % g.c_inner) * g.c_inner_stride +
                       (int64_t)p * g.ldc;
            else
                coff = (int64_t)seq * g.c_seq_stride + (int64_t)p * g.ldc;
        } else {
            coff = (int64_t)m * g.ldc;
        }
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const int n = n0 + tx * TN + j;
            float4 o = make_float4(acc[i][j / 2].x, acc[i][j / 2].y, acc[i][j / 2 + 1].x, acc[i][j / 2 + 1].y);
            if (g.bias) {
                const float4 bb = *reinterpret_cast<const float4*>(g.bias + n);
                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (g.prelu) {
                o.x = prelu(o.x, slope); o.y = prelu(o.y, slope);
                o.z = prelu(o.z, slope); o.w = prelu(o.w, slope);
            }
            if (g.prelu_vec) {
                const float4 sv = *reinterpret_cast<const float4*>(g.prelu_vec + n);
                o.x = prelu(o.x, sv.x); o.y = prelu(o.y, sv.y);
                o.z = prelu(o.z, sv.z); o.w = prelu(o.w, sv.w);
            }
            if (g.R) {
                const float4 rr = *reinterpret_cast<const float4*>(g.R + coff + n);
                o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            *reinterpret_cast<float4*>(g.C + coff + n) = o;
        }
    }
}

template <int BM, int BN, int TM, int TN>
inline cudaError_t configure_rows_gemm_cfg() {
    const size_t smem = (size_t)(GK * (BM + 4) + GK * BN) * sizeof(float);
    return cudaFuncSetAttribute(rows_gemm_kernel<BM, BN, TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem);
}

// call once per process before the first launch (and outside stream capture)
inline cudaError_t configure_rows_gemm() {
    cudaError_t e = configure_rows_gemm_cfg<16, 64, 2, 4>();
    if (e == cudaSuccess) e = configure_rows_gemm_cfg<64, 128, 4, 8>();
    if (e == cudaSuccess) e = configure_rows_gemm_cfg<64, 64, 4, 4>();
    return e;
}

template <int BM, int BN, int TM, int TN>
inline cudaError_t launch_rows_gemm_cfg(const GemmArgs& g, cudaStream_t st, bool pdl) {
    constexpr int NT = (BM / TM) * (BN / TN);
    const size_t smem = (size_t)(GK * (BM + 4) + GK * BN) * sizeof(float);
    dim3 grid((g.M + BM - 1) / BM, g.N / BN);
    return launch_k(pdl, rows_gemm_kernel<BM, BN, TM, TN>, grid, dim3(NT), smem, st, g);
}

// Pick a tile by problem size: small M (one streaming frame = 97 rows) wants many small CTAs,
// large M wants the 64x128 tile.  N must be a multiple of 64; K a multiple of 64.
inline cudaError_t launch_rows_gemm(const GemmArgs& g, cudaStream_t st, bool pdl = false) {
    if (g.N % 64 != 0 || g.K % 64 != 0 || g.M <= 0) return cudaErrorInvalidValue;
    if (g.ln_g && g.K != 64) return cudaErrorInvalidValue;
    if (g.M <= 2048) return launch_rows_gemm_cfg<16, 64, 2, 4>(g, st, pdl);     // 128 threads
    if (g.N % 128 == 0) return launch_rows_gemm_cfg<64, 128, 4, 8>(g, st, pdl); // 256 threads
    return launch_rows_gemm_cfg<64, 64, 4, 4>(g, st, pdl);                      // 256 threads
}

}  // namespace l2h
