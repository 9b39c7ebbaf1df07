// mid_kernel: for one-frame calls (T == 1) everything between the intra BiLSTM and the attention
// is ROW-LOCAL (one row = one frequency bin of one stream):
//     X1 = X + Y W_l1^T + b                      intra Linear(128->64) + residual   (tfgridnet_causal.py:513-516)
//     g  = W_ih LN(X1) + W_hh h + b ; (h,c) cell   inter LSTM, one step, carried state   (:521-532)
//     X2 = X1 + h' W_l2^T + b                     inter Linear(64->64) + residual      (:534-538)
//     P  = PReLU(X2 W_qkv^T + b)                  the three attention projections      (:547-551, :354-387)
// The generic chain runs this as four row-GEMMs + one recurrence launch + the GEMM half of
// qkv_kernel.  Here one CTA keeps all five weight matrices (204 KB) in shared memory -- loaded
// BEFORE griddepcontrol.wait, i.e. while the 97-step intra recurrence is still running -- and
// pushes a tile of RT rows through the whole section.  grid (ceil(97/RT), B), 256 threads.
#pragma once
#include "common.cuh"
#include "sep_kernels.cuh"

namespace l2h {

constexpr int MID_RT = 8;
constexpr size_t MID_SMEM = (size_t)(128 * 64 + 64 * 256 + 64 * 256 + 64 * 64 + 64 * NQKV   // weights
                                     + 128 * MID_RT + 4 * 64 * MID_RT + 256 * MID_RT) * sizeof(float);

__global__ void __launch_bounds__(256)
mid_kernel(const float* __restrict__ Y, float* __restrict__ X, float* __restrict__ QKV, float* __restrict__ state,
           int64_t sstride, int blk, BlockWeights w) {
    extern __shared__ __align__(16) float sm[];
    float* W1 = sm;                       // [128][64]   intra linear, k-major
    float* W2 = W1 + 128 * 64;            // [64][256]   inter W_ih, k-major, gate cols j*4+q
    float* W3 = W2 + 64 * 256;            // [64][256]   inter W_hh, k-major
    float* W4 = W3 + 64 * 256;            // [64][64]    inter linear
    float* W5 = W4 + 64 * 64;             // [64][112]   q|k|v projections
    float* yT = W5 + 64 * NQKV;           // [128][RT]   intra LSTM outputs, k-major
    float* x1 = yT + 128 * MID_RT;        // [RT][64]
    float* aT = x1 + 64 * MID_RT;         // [64][RT]    LN(X1), k-major
    float* hT = aT + 64 * MID_RT;         // [64][RT]    h (old, then new), k-major
    float* x2T = hT + 64 * MID_RT;        // [64][RT]    X2, k-major
    float* gt = x2T + 64 * MID_RT;        // [RT][256]   gate pre-activations

    __shared__ __align__(8) unsigned long long wbar;
    griddep_launch();
    const int tid = threadIdx.x, b = blockIdx.y;
    const int r0 = blockIdx.x * MID_RT;
    const int nr = min(MID_RT, NF - r0);
    // ---- weights -> smem: five TMA bulk copies (independent of the chain, so issued before the wait)
    if (tid == 0) { mbar_init(&wbar, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&wbar, (128 * 64 + 64 * 256 + 64 * 256 + 64 * 64 + 64 * NQKV) * 4);
    __syncthreads();
    tma_load_split(W1, w.wl1_t, 128 * 64 * 4, &wbar, tid, 256);
    tma_load_split(W2, w.wih2_t, 64 * 256 * 4, &wbar, tid, 256);
    tma_load_split(W3, w.whh2_t, 64 * 256 * 4, &wbar, tid, 256);
    tma_load_split(W4, w.wl2_t, 64 * 64 * 4, &wbar, tid, 256);
    tma_load_split(W5, w.wqkv_t, 64 * NQKV * 4, &wbar, tid, 256);
    griddep_wait();
    float* sb = state + sizeof(StateHeader) / 4 + (int64_t)b * sstride + ST_BLK + (int64_t)blk * BK_STRIDE;
    float* hst = sb + BK_H;
    float* cst = sb + BK_C;
    const int64_t row0 = (int64_t)b * NF + r0;
    // ---- tile loads: Y -> yT (k-major), h -> hT ------------------------------------------------
    for (int i = tid; i < MID_RT * 32; i += 256) {           // float4 over 128 k
        const int r = i / 32, k4 = i % 32;
        const float4 v = (r < nr) ? *reinterpret_cast<const float4*>(Y + (row0 + r) * 128 + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        yT[(k4 * 4 + 0) * MID_RT + r] = v.x; yT[(k4 * 4 + 1) * MID_RT + r] = v.y;
        yT[(k4 * 4 + 2) * MID_RT + r] = v.z; yT[(k4 * 4 + 3) * MID_RT + r] = v.w;
    }
    for (int i = tid; i < MID_RT * 16; i += 256) {
        const int r = i / 16, k4 = i % 16;
        const float4 v = (r < nr) ? *reinterpret_cast<const float4*>(hst + (r0 + r) * 64 + k4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        hT[(k4 * 4 + 0) * MID_RT + r] = v.x; hT[(k4 * 4 + 1) * MID_RT + r] = v.y;
        hT[(k4 * 4 + 2) * MID_RT + r] = v.z; hT[(k4 * 4 + 3) * MID_RT + r] = v.w;
    }
    mbar_wait(&wbar, 0);
    __syncthreads();
    const int c = tid & 63, rp = tid >> 6;                  // column c, row pair (2rp, 2rp+1)
    // ---- phase 1: X1 = X + Y W1 + b -----------------------------------------------------------
    {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < 128; ++k) {
            const float wv = W1[k * 64 + c];
            const float2 yv = *reinterpret_cast<const float2*>(yT + k * MID_RT + rp * 2);
            acc = ffma2(make_float2(wv, wv), yv, acc);
        }
        const float bias = __ldg(w.bl1 + c);
        const float xa = (rp * 2 < nr) ? X[(row0 + rp * 2) * 64 + c] : 0.f;
        const float xb = (rp * 2 + 1 < nr) ? X[(row0 + rp * 2 + 1) * 64 + c] : 0.f;
        x1[(rp * 2) * 64 + c] = xa + acc.x + bias;
        x1[(rp * 2 + 1) * 64 + c] = xb + acc.y + bias;
    }
    __syncthreads();
    // ---- phase 2: LayerNorm over channels, one warp per row -> aT --------------------------------
    {
        const int r = tid >> 5, lane = tid & 31;
        const float v0 = x1[r * 64 + lane], v1 = x1[r * 64 + lane + 32];
        const float mu = warp_sum(v0 + v1) * (1.f / 64.f);
        const float d0 = v0 - mu, d1 = v1 - mu;
        const float rs = rsqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.f / 64.f) + 1e-5f);
        aT[lane * MID_RT + r] = d0 * rs * __ldg(w.ln2_g + lane) + __ldg(w.ln2_b + lane);
        aT[(lane + 32) * MID_RT + r] = d1 * rs * __ldg(w.ln2_g + lane + 32) + __ldg(w.ln2_b + lane + 32);
    }
    __syncthreads();
    // ---- phase 3: gate pre-activations, thread = gate column, all RT rows --------------------------
    {
        float2 acc[MID_RT / 2];
        const float bias = __ldg(w.b2 + tid);
#pragma unroll
        for (int i = 0; i < MID_RT / 2; ++i) acc[i] = make_float2(bias, bias);
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
            const float w2 = W2[k * 256 + tid], w3 = W3[k * 256 + tid];
            const float2 ww2 = make_float2(w2, w2), ww3 = make_float2(w3, w3);
            const float4 a0 = *reinterpret_cast<const float4*>(aT + k * MID_RT);
            const float4 a1 = *reinterpret_cast<const float4*>(aT + k * MID_RT + 4);
            const float4 h0 = *reinterpret_cast<const float4*>(hT + k * MID_RT);
            const float4 h1 = *reinterpret_cast<const float4*>(hT + k * MID_RT + 4);
            acc[0] = ffma2(ww2, make_float2(a0.x, a0.y), acc[0]); acc[1] = ffma2(ww2, make_float2(a0.z, a0.w), acc[1]);
            acc[2] = ffma2(ww2, make_float2(a1.x, a1.y), acc[2]); acc[3] = ffma2(ww2, make_float2(a1.z, a1.w), acc[3]);
            acc[0] = ffma2(ww3, make_float2(h0.x, h0.y), acc[0]); acc[1] = ffma2(ww3, make_float2(h0.z, h0.w), acc[1]);
            acc[2] = ffma2(ww3, make_float2(h1.x, h1.y), acc[2]); acc[3] = ffma2(ww3, make_float2(h1.z, h1.w), acc[3]);
        }
#pragma unroll
        for (int i = 0; i < MID_RT / 2; ++i) {
            gt[(2 * i) * 256 + tid] = acc[i].x;
            gt[(2 * i + 1) * 256 + tid] = acc[i].y;
        }
    }
    __syncthreads();
    // ---- phase 4: LSTM cell; thread (hidden unit c, row pair) -------------------------------------
    {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = rp * 2 + u;
            const float4 g4 = *reinterpret_cast<const float4*>(gt + r * 256 + c * 4);      // i, f, g, o
            const float gi = fast_sigmoid(g4.x), gf = fast_sigmoid(g4.y), gg = fast_tanh(g4.z), go = fast_sigmoid(g4.w);
            const float cold = (r < nr) ? cst[(r0 + r) * 64 + c] : 0.f;
            const float cn = gf * cold + gi * gg;
            const float hn = go * fast_tanh(cn);
            if (r < nr) { cst[(r0 + r) * 64 + c] = cn; hst[(r0 + r) * 64 + c] = hn; }
            x2T[c * MID_RT + r] = hn;       // stage h' k-major (x2T is free until phase 5 writes it)
        }
    }
    __syncthreads();
    // ---- phase 5: X2 = X1 + h' W4 + b ----------------------------------------------------------------
    float2 x2v;
    {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float wv = W4[k * 64 + c];
            const float2 hv = *reinterpret_cast<const float2*>(x2T + k * MID_RT + rp * 2);
            acc = ffma2(make_float2(wv, wv), hv, acc);
        }
        const float bias = __ldg(w.bl2 + c);
        x2v = make_float2(x1[(rp * 2) * 64 + c] + acc.x + bias, x1[(rp * 2 + 1) * 64 + c] + acc.y + bias);
        if (rp * 2 < nr) X[(row0 + rp * 2) * 64 + c] = x2v.x;
        if (rp * 2 + 1 < nr) X[(row0 + rp * 2 + 1) * 64 + c] = x2v.y;
    }
    __syncthreads();                      // everyone has consumed h' from x2T
    *reinterpret_cast<float2*>(x2T + c * MID_RT + rp * 2) = x2v;
    __syncthreads();
    // ---- phase 6: P = PReLU(X2 W5 + b); item = (column n, row pair) ---------------------------------
    for (int it = tid; it < NQKV * (MID_RT / 2); it += 256) {
        const int n = it % NQKV, q = it / NQKV;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll 8
        for (int k = 0; k < 64; ++k) {
            const float wv = W5[k * NQKV + n];
            const float2 xv = *reinterpret_cast<const float2*>(x2T + k * MID_RT + q * 2);
            acc = ffma2(make_float2(wv, wv), xv, acc);
        }
        const float bias = __ldg(w.bqkv + n);
        const float slope = __ldg(w.slopes + (n < 24 ? 0 : (n < 48 ? 1 : 2)));
        if (q * 2 < nr) QKV[(row0 + q * 2) * NQKV + n] = prelu(acc.x + bias, slope);
        if (q * 2 + 1 < nr) QKV[(row0 + q * 2 + 1) * NQKV + n] = prelu(acc.y + bias, slope);
    }
}

}  // namespace l2h
