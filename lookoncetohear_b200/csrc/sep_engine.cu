// Separation engine: weight packing, state, the per-call kernel chain and the C ABI
// (include/lookonce_b200.h).  Host side of the hot path = this file; no torch anywhere.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <tuple>
#include <string>
#include <vector>

#include "../../include/lookonce_b200.h"
#include "gemm.cuh"
#include "umma_host.cuh"
#include "tc_lstm.cuh"
#include "lstm.cuh"
#include "sep_kernels.cuh"
#include "mid_kernel.cuh"
#include "hop_kernels.cuh"

namespace l2h {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(3, std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
    } while (0)

struct Slot {
    int64_t off;    // floats into the packed buffer
    int64_t numel;  // of the reference tensor
    // repack(src_host, dst_host_base)
    std::function<void(const float*, float*)> repack;
    bool loaded = false;
    bool accumulate = false;   // LSTM biases: b_ih and b_hh of one direction sum into one destination
    std::vector<float> raw;    // accumulate slots keep their tensor; the sum is formed at commit
};

// (sequence, direction) pairs from which the recurrence runs on the tensor cores (tc_lstm: 32 sequences per CTA, 9 472 per wave of
// 296 CTAs at 16.8 us per step; CUDA cores: 4 per CTA, 1 184 per wave at 2.7 us).  Measured at the offline shape
// (tools/offline_split_experiment.py): 3 492 inter sequences are faster as three CUDA-core waves than as 110 tensor-core CTAs.
constexpr int TCL_MIN_SEQDIRS = 4096;

struct SepEngine {
    l2h_sep_config cfg;
    int n_blocks;
    std::vector<float> host;        // staging
    float* dev = nullptr;           // packed weights
    int device = -1;                // ordinal of the device that owns `dev`, the streams, events and cached graphs
    int64_t weight_gen = 0;         // bumped by every commit
    bool graph_stats = false;
    // tensor-core path: bf16 hi/lo planes [2][N][K] of the GEMM weights, built on the device at commit
    struct PlaneSrc { int64_t wt_off; int K, N, ld, col0; int64_t plane_off; };
    std::vector<PlaneSrc> plane_srcs;
    std::vector<int64_t> plane_of;      // per block: plane offsets of wih1, wl1, wih2, wl2, wqkv, [wih2|whh2], wp
    std::vector<int64_t> plane_stash;
    __nv_bfloat16* planes = nullptr;
    int64_t planes_total = 0;
    bool cur_pdl = false;               // PDL attribute for the tensor-core launches of the chain being enqueued
    bool fuse_ih = false;               // many sequences: W_ih + LayerNorm inside the tensor-core recurrence (option "fuse_ih").  Off:
                                        // measured slower than GEMM + tc_lstm (offline B=16: 5.48 vs 4.50 ms per chain, profiles/r02i)
    bool use_tc = true;                 // rows > TC_MIN_ROWS: dense contractions on tcgen05 (option "tensor_cores")
    int tc_passes = 3;                  // 3 = bf16x3 split products (fp32 configs); 2 = bf16 weights x split activations (option
                                        // "bf16" = 1: the offline bf16 configuration); 1 = plain bf16 operands ("bf16" = 2)
    int64_t total = 0;
    std::map<std::string, Slot> slots;
    SepWeights w;
    std::vector<BlockWeights> bw;
    bool committed = false;
    // CUDA-graph cache of whole kernel chains (the T=1 streaming chain is ~30 tiny kernels:
    // launch-bound unless replayed as a graph)
    std::map<std::vector<int64_t>, cudaGraphExec_t> graphs;
    std::map<cudaGraphExec_t, int> graph_kernels;   // kernel nodes per cached graph
    TraceRec* trace_dev = nullptr;                  // device trace buffer (l2h_sep_trace_start)
    int trace_cap = 0;
    int64_t launch_count = 0;                       // kernels launched so far (graph replays counted by their kernel nodes)
    cudaStream_t cap_stream = nullptr;
    struct MidSrc { int64_t wl1, wih2, whh2t, wl2, wqkv, dst, slopes, slope_vec; };
    std::vector<MidSrc> mid_src;   // per block: host offsets the packed mid_kernel weights are derived from at commit
    cudaStream_t pipe_streams[128] = {};
    std::vector<cudaEvent_t> pipe_events;
    int64_t pipe_budget = 0; // bytes the pipelined workspace may take (min(24 GB, half of the free memory at first use))
    int pipe_frames = 0;     // one-hop chains per pipelined graph (<= PIPE_MAX_FRAMES); 0 = auto: as many as a 24 GB workspace holds
    int pipe_alanes = 12;    // BiLSTM (stage A) hops in flight per block (<= PIPE_LANES)
    int pipe_gemm_shape = 0; // tile shape of the pipelined W_ih GEMM (gemm.cuh: launch_rows_gemm), option "pipeline_gemm_shape"
    int pipe_midb_hops = 4;       // pipeline: consecutive hops one mid_b launch takes (<= PIPE_MIDB_MAX)
    int pipe_skip = 0;            // DEBUG (timing experiments only): bit mask of pipeline stages NOT to launch
    int pipe_pdl = 16;            // pipeline: stages launched with programmatic dependent launch (bit mask; 16 = mid_b)
    bool pipe_split_mid = true;   // pipeline: mid section as mid_a | mid_b (serial) | mid_c
    int pipe_qlanes = 3;     // qkv hops in flight per block (<= PIPE_QLANES)
    int pipe_clanes = 2;     // mid_c hops in flight per block (<= PIPE_CLANES)
    int pipe_tlanes = 3;     // attention hops in flight per block (<= PIPE_TLANES)
    int pipe_olanes = 4;     // attn_out hops in flight per block (<= PIPE_OLANES)
    int pipe_flanes = 6;     // front_kernel hops in flight (<= PIPE_FLANES)
    int pipe_blanes = 6;     // back_kernel hops in flight (<= PIPE_BLANES)
    bool use_pipe = true;    // wavefront pipelining of one-frame calls inside a multi-frame graph (option "pipeline")
    bool fold_mid_c = false;      // no mid_c / no projection in the mid kernels: Linear in mid_b2, Q/K/V projection in qkv (untested)
    bool mid_split_large = true;  // many streams: run the fused mid section as mid_a | mid_b | mid_c (2-4 CTAs per SM)
    bool use_mid = true;     // fused row-local mid-section for one-frame calls (option "fused_mid")
    int tcl_min_seqdirs = TCL_MIN_SEQDIRS;  // (sequence, direction) pairs from which the recurrence runs on the tensor cores (option "tc_lstm_min")
    int tc_pdl = 7;              // programmatic dependent launch around the tensor-core GEMMs of many-row calls: bit 0 the many-stream mid section, bit 1 W_ih and the out projection + its LayerNorm kernel, bit 2 the persistent qkv kernel (256 streams: 0.847 -> 0.828 ms per hop-step; PDL on EVERY kernel of that chain was slower: 0.939; option "tc_pdl")
    bool use_back_many = true;   // calls of several frames: front_many_kernel / back_many_kernel (one CTA / cluster per chunk of frames) instead of one per frame (option "back_many")
    bool use_tail = true;    // one-hop calls of a few streams: mid + qkv + attention + attn_out (+ next W_ih) as ONE 16-CTA cluster kernel (option "fused_tail")
    bool use_pdl = true;     // programmatic dependent launch between the kernels of a chain (option "pdl")
};

static int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

// gate-row permutation: packed row j*4+q <- reference row q*64+j
static inline int perm_row(int p) { return (p & 3) * 64 + (p >> 2); }

static void build_layout(SepEngine* e) {
    int64_t cur = 0;
    auto alloc = [&](int64_t n) { int64_t o = cur; cur = align4(cur + n); return o; };
    auto& S = e->slots;
    auto plain = [&](const std::string& name, int64_t n) {
        int64_t o = alloc(n);
        S[name] = Slot{o, n, [o, n](const float* s, float* d) { memcpy(d + o, s, n * sizeof(float)); }};
        return o;
    };
    auto transposed = [&](const std::string& name, int rows, int cols, int ld_out) {
        // reference [rows][cols] -> packed [cols][ld_out] (k-major)
        int64_t o = alloc((int64_t)cols * ld_out);
        S[name] = Slot{o, (int64_t)rows * cols, [o, rows, cols, ld_out](const float* s, float* d) {
            for (int r = 0; r < rows; ++r)
                for (int c = 0; c < cols; ++c) d[o + (int64_t)c * ld_out + r] = s[(int64_t)r * cols + c];
        }};
        return o;
    };
    const std::string P = "tfgridnet.";
    std::vector<std::pair<const float**, int64_t>> fix;   // pointer fields to resolve after alloc
    auto bind = [&](const float** field, int64_t off) { fix.push_back({field, off}); };

    // STFT filterbanks [194][1][192]
    bind(&e->w.wat, transposed(P + "enc.filterbank._filters", NROW, NFFT, 196));
    bind(&e->w.ws, plain(P + "dec.filterbank._filters", (int64_t)NROW * NFFT));
    bind(&e->w.wc, plain(P + "conv.0.weight", 64 * 36));
    bind(&e->w.bc, plain(P + "conv.0.bias", 64));
    bind(&e->w.we, plain(P + "embed_to_feats_proj.0.weight", (int64_t)FC * SPK));
    bind(&e->w.be, plain(P + "embed_to_feats_proj.0.bias", FC));
    bind(&e->w.lne_g, plain(P + "embed_to_feats_proj.1.weight", FC));
    bind(&e->w.lne_b, plain(P + "embed_to_feats_proj.1.bias", FC));
    bind(&e->w.wd, plain(P + "deconv.weight", 64 * 36));
    bind(&e->w.bd, plain(P + "deconv.bias", 4));

    e->bw.resize(e->n_blocks);
    for (int b = 0; b < e->n_blocks; ++b) {
        BlockWeights& W = e->bw[b];
        const std::string B = P + "blocks." + std::to_string(b) + ".";
        bind(&W.ln1_g, plain(B + "intra_norm.norm.weight", 64));
        bind(&W.ln1_b, plain(B + "intra_norm.norm.bias", 64));
        bind(&W.ln2_g, plain(B + "inter_norm.norm.weight", 64));
        bind(&W.ln2_b, plain(B + "inter_norm.norm.bias", 64));
        // LSTM input weights [256][64] -> Wt[k][dirofs + p], p = j*4+q
        auto ih = [&](const std::string& name, int64_t base, int ld, int dirofs) {
            S[name] = Slot{base, 256 * 64, [base, ld, dirofs](const float* s, float* d) {
                for (int p = 0; p < 256; ++p) {
                    const int r = perm_row(p);
                    for (int k = 0; k < 64; ++k) d[base + (int64_t)k * ld + dirofs + p] = s[r * 64 + k];
                }
            }};
        };
        auto hh = [&](const std::string& name, int64_t base) {
            S[name] = Slot{base, 256 * 64, [base](const float* s, float* d) {
                for (int p = 0; p < 256; ++p) memcpy(d + base + (int64_t)p * 64, s + perm_row(p) * 64, 64 * sizeof(float));
            }};
        };
        auto bias = [&](const std::string& name, int64_t base, bool) {
            Slot sl{base, 256, [base](const float* s, float* d) {
                for (int p = 0; p < 256; ++p) d[base + p] += s[perm_row(p)];
            }};
            sl.accumulate = true;
            S[name] = sl;
        };
        const int64_t wih1 = alloc(64 * 512), b1 = alloc(512), whh1 = alloc(2 * 256 * 64);
        ih(B + "intra_rnn.weight_ih_l0", wih1, 512, 0);
        ih(B + "intra_rnn.weight_ih_l0_reverse", wih1, 512, 256);
        hh(B + "intra_rnn.weight_hh_l0", whh1);
        hh(B + "intra_rnn.weight_hh_l0_reverse", whh1 + 256 * 64);
        bias(B + "intra_rnn.bias_ih_l0", b1, true);
        bias(B + "intra_rnn.bias_hh_l0", b1, true);
        bias(B + "intra_rnn.bias_ih_l0_reverse", b1 + 256, true);
        bias(B + "intra_rnn.bias_hh_l0_reverse", b1 + 256, true);
        bind(&W.wih1_t, wih1); bind(&W.b1, b1); bind(&W.whh1, whh1);
        const int64_t wl1 = transposed(B + "intra_linear.weight", 64, 128, 64);
        bind(&W.wl1_t, wl1);
        bind(&W.bl1, plain(B + "intra_linear.bias", 64));
        const int64_t wih2 = alloc(64 * 256), b2 = alloc(256), whh2 = alloc(256 * 64), whh2t = alloc(64 * 256);
        ih(B + "inter_rnn.weight_ih_l0", wih2, 256, 0);
        S[B + "inter_rnn.weight_hh_l0"] = Slot{whh2, 256 * 64, [whh2, whh2t](const float* s, float* d) {
            for (int p = 0; p < 256; ++p) {
                const int r = perm_row(p);
                memcpy(d + whh2 + (int64_t)p * 64, s + r * 64, 64 * sizeof(float));
                for (int k = 0; k < 64; ++k) d[whh2t + (int64_t)k * 256 + p] = s[r * 64 + k];
            }
        }};
        bind(&W.whh2_t, whh2t);
        bias(B + "inter_rnn.bias_ih_l0", b2, true);
        bias(B + "inter_rnn.bias_hh_l0", b2, true);
        bind(&W.wih2_t, wih2); bind(&W.b2, b2); bind(&W.whh2, whh2);
        const int64_t wl2 = transposed(B + "inter_linear.weight", 64, 64, 64);
        bind(&W.wl2_t, wl2);
        bind(&W.bl2, plain(B + "inter_linear.bias", 64));
        // Q | K | V projections -> one [64][112] k-major matrix
        const int64_t wqkv = alloc(64 * NQKV), bqkv = alloc(NQKV), slopes = alloc(4);
        auto proj = [&](const std::string& mod, int rows, int col0, int slope_idx) {
            S[B + mod + ".0.weight"] = Slot{wqkv, (int64_t)rows * 64, [wqkv, rows, col0](const float* s, float* d) {
                for (int r = 0; r < rows; ++r)
                    for (int k = 0; k < 64; ++k) d[wqkv + (int64_t)k * NQKV + col0 + r] = s[r * 64 + k];
            }};
            S[B + mod + ".0.bias"] = Slot{bqkv, rows, [bqkv, rows, col0](const float* s, float* d) {
                memcpy(d + bqkv + col0, s, rows * sizeof(float));
            }};
            S[B + mod + ".1.weight"] = Slot{slopes, 1, [slopes, slope_idx](const float* s, float* d) {
                d[slopes + slope_idx] = s[0];
            }};
        };
        proj("attn_conv_Q", NHEAD * QE, 0, 0);
        proj("attn_conv_K", NHEAD * QE, NHEAD * QE, 1);
        proj("attn_conv_V", NHEAD * VD, 2 * NHEAD * QE, 2);
        bind(&W.wqkv_t, wqkv); bind(&W.bqkv, bqkv); bind(&W.slopes, slopes);
        const int64_t slope_vec = alloc(NQKV);
        bind(&W.slope_vec, slope_vec);
        const int64_t midp = alloc(MID_PACK);
        bind(&W.mid_pack, midp);
        e->mid_src.push_back({wl1, wih2, whh2t, wl2, wqkv, midp, slopes, slope_vec});
        {   // tensor-core B operands of this block (offsets into the packed fp32 buffer; planes are made at commit)
            auto reg = [&](int64_t wt, int K, int N, int ld, int col0, int64_t at) {
                e->plane_srcs.push_back({wt, K, N, ld, col0, at});
            };
            auto take = [&](int K, int N) { const int64_t o = e->planes_total; e->planes_total += ((int64_t)K * N + 63) & ~int64_t(63); return o; };
            const int64_t p_ih1 = take(64, 512), p_l1 = take(128, 64), p_ih2 = take(64, 256), p_l2 = take(64, 64), p_qkv = take(64, NQKV),
                          p_cat = take(128, 256);
            reg(wih1, 64, 512, 64, 0, p_ih1); reg(wl1, 128, 64, 128, 0, p_l1); reg(wih2, 64, 256, 64, 0, p_ih2);
            reg(wl2, 64, 64, 64, 0, p_l2); reg(wqkv, 64, NQKV, 64, 0, p_qkv);
            reg(wih2, 64, 256, 128, 0, p_cat); reg(whh2t, 64, 256, 128, 64, p_cat);      // [W_ih | W_hh]: k = [x | h]
            e->plane_stash = {p_ih1, p_l1, p_ih2, p_l2, p_qkv, p_cat};
        }
        bind(&W.lnq_g, plain(B + "attn_conv_Q.3.norm.weight", QK_DIM));
        bind(&W.lnq_b, plain(B + "attn_conv_Q.3.norm.bias", QK_DIM));
        bind(&W.lnk_g, plain(B + "attn_conv_K.3.norm.weight", QK_DIM));
        bind(&W.lnk_b, plain(B + "attn_conv_K.3.norm.bias", QK_DIM));
        bind(&W.lnv_g, plain(B + "attn_conv_V.3.norm.weight", V_DIM));
        bind(&W.lnv_b, plain(B + "attn_conv_V.3.norm.bias", V_DIM));
        const int64_t wp = transposed(B + "attn_concat_proj.0.weight", 64, 64, 64);
        bind(&W.wp_t, wp);
        {
            const int64_t p_p = e->planes_total;
            e->planes_total += 64 * 64;
            e->plane_srcs.push_back({wp, 64, 64, 64, 0, p_p});
            for (int64_t v : e->plane_stash) e->plane_of.push_back(v);
            e->plane_of.push_back(p_p);
        }
        bind(&W.bp, plain(B + "attn_concat_proj.0.bias", 64));
        S[B + "attn_concat_proj.1.weight"] = Slot{slopes, 1, [slopes](const float* s, float* d) { d[slopes + 3] = s[0]; }};
        bind(&W.lnp_g, plain(B + "attn_concat_proj.3.norm.weight", FC));
        bind(&W.lnp_b, plain(B + "attn_concat_proj.3.norm.bias", FC));
    }
    e->total = cur;
    e->host.assign(cur, 0.f);
    // stash offsets in the pointer fields; resolved to device addresses at commit
    for (auto& f : fix) *f.first = reinterpret_cast<const float*>(f.second);
}

// pointer fields hold offsets (floats) until the first commit; `base` turns them into device addresses, and a
// later re-commit on another device shifts them by (new base - old base)
static void shift_pointers(SepEngine* e, const float* new_base, const float* old_base) {
    auto fixp = [&](const float*& p) {
        if (old_base == nullptr) p = new_base + reinterpret_cast<int64_t>(p);      // offset -> address
        else p = new_base + (p - old_base);                                        // address on the old device -> new one
    };
    SepWeights& w = e->w;
    fixp(w.wat); fixp(w.ws); fixp(w.wc); fixp(w.bc); fixp(w.we); fixp(w.be); fixp(w.lne_g); fixp(w.lne_b);
    fixp(w.wd); fixp(w.bd);
    for (auto& W : e->bw) {
        fixp(W.ln1_g); fixp(W.ln1_b); fixp(W.wih1_t); fixp(W.b1); fixp(W.whh1); fixp(W.wl1_t); fixp(W.bl1);
        fixp(W.ln2_g); fixp(W.ln2_b); fixp(W.wih2_t); fixp(W.b2); fixp(W.whh2); fixp(W.whh2_t); fixp(W.mid_pack); fixp(W.wl2_t); fixp(W.bl2);
        fixp(W.wqkv_t); fixp(W.bqkv); fixp(W.slopes); fixp(W.slope_vec); fixp(W.lnq_g); fixp(W.lnq_b); fixp(W.lnk_g);
        fixp(W.lnk_b); fixp(W.lnv_g); fixp(W.lnv_b); fixp(W.wp_t); fixp(W.bp); fixp(W.lnp_g); fixp(W.lnp_b);
    }
}

// ---- workspace carve-up (floats) ---------------------------------------------------------------
struct Workspace {
    int64_t X, GX, Y, Z, Q, KALL, VALL, PRE, QKVRAW, TAPS, total;
};
// few frames in flight -> split every head's 50-row window over several CTAs
static int attn_splits(int B, int T) {
    const int ctas = B * T * NHEAD;
    return (ctas * ATT_CL <= 296) ? ATT_CL : 1;      // 8-CTA clusters while they all fit in one wave
}

static Workspace carve(int n_blocks, int B, int T, uint32_t flags) {
    Workspace ws;
    const int64_t rows = (int64_t)B * T * NF;
    int64_t cur = 0;
    auto alloc = [&](int64_t n) { int64_t o = cur; cur = (cur + n + 31) & ~int64_t(31); return o; };
    ws.X = alloc(rows * 64);
    ws.GX = alloc(rows * 512);
    ws.Y = alloc(rows * 128);
    ws.Z = alloc(rows * 64);
    ws.Q = alloc((int64_t)B * NHEAD * T * QK_LD);
    ws.KALL = alloc(T > 1 ? (int64_t)B * NHEAD * (ATT - 1 + T) * QK_LD : 0);
    ws.VALL = alloc(T > 1 ? (int64_t)B * NHEAD * (ATT - 1 + T) * V_DIM : 0);
    ws.PRE = alloc((int64_t)B * FC);
    ws.QKVRAW = alloc(rows * NQKV);
    ws.TAPS = alloc((flags & L2H_FLAG_TAPS) ? (int64_t)(1 + 3 * n_blocks) * rows * 64 : 0);
    ws.total = (cur + 511) & ~int64_t(511);      // a multiple of one GX row: pipelined hops address their slots as rows of one tensor
    return ws;
}

// grids of the mid-section kernels: persistent over (stream, row tile) items, as many CTAs per SM as their shared memory
// allows (fused 225 KB -> 1, mid_a 107 KB -> 2, mid_b 68 KB -> 3, mid_c 50 KB -> 4)
static inline int mid_items(int B) { return B * ((NF + MID_RT - 1) / MID_RT); }
static inline dim3 mid_grid_for(int B, int ctas_per_sm) { return dim3(std::min(148 * ctas_per_sm, mid_items(B))); }
// many streams: the section is throughput-bound and one fused CTA per SM (8 warps) cannot hide its own latencies; the
// three smaller kernels run 2-4 CTAs per SM (same arithmetic, +3 small global round trips)
static inline bool mid_split_for_throughput(int B) { return mid_items(B) > 148; }

// cudaFuncSetAttribute applies to the CURRENT device: keep one flag per device ordinal
static bool g_attr_done[64] = {};
static int g_tail_clusters[64] = {};     // how many 16-CTA tail_kernel clusters the device can hold at once (0: cannot be scheduled)
static int set_attrs() {
    int dev_ord = 0;
    CK(cudaGetDevice(&dev_ord));
    if (dev_ord < 0 || dev_ord >= 64) return fail(1, "device ordinal out of range");
    if (g_attr_done[dev_ord]) return 0;
    CK(cudaFuncSetAttribute(qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QKV_SMEM));
    CK(cudaFuncSetAttribute(qkv_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QKV_MANY_SMEM));
    CK(cudaFuncSetAttribute(attn_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AOUT_SMEM));
    CK(cudaFuncSetAttribute(back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BACK_SMEM));
    CK(cudaFuncSetAttribute(back_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BACK_MANY_SMEM));
    CK(cudaFuncSetAttribute(front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT_SMEM));
    CK(cudaFuncSetAttribute(front_many_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT_SMEM));
    CK(cudaFuncSetAttribute(mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM));
    CK(cudaFuncSetAttribute(mid_noproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_SMEM));
    CK(cudaFuncSetAttribute(mid_b2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_B2_SMEM));
    CK(cudaFuncSetAttribute(mid_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_A_SMEM));
    CK(cudaFuncSetAttribute(mid_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_B_SMEM));
    CK(cudaFuncSetAttribute(mid_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MID_C_SMEM));
    CK(configure_rows_gemm());
    CK(configure_lstm());
    CK(umma::configure());
    CK(configure_tc_lstm());
    CK(cudaFuncSetAttribute(front1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT1_SMEM));
    {   // tail_kernel: 16 CTAs per cluster is a non-portable size; ask whether this device can place it
        g_tail_clusters[dev_ord] = 0;
        if (cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_SMEM) == cudaSuccess &&
            cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(TAIL_CL, 1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = TAIL_SMEM;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = TAIL_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, tail_kernel, &cfg) == cudaSuccess) g_tail_clusters[dev_ord] = ncl;
        }
        cudaGetLastError();
    }
    g_attr_done[dev_ord] = true;
    return 0;
}


// ---- dense contractions on the tensor cores (csrc/umma_gemm.cuh) for calls with many rows --------------------------
constexpr int64_t TC_MIN_ROWS = 2048;      // below this the 16-row CUDA-core tiles win (one streaming frame = 97 rows)
enum { PL_IH1 = 0, PL_L1, PL_IH2, PL_L2, PL_QKV, PL_CAT, PL_P, PL_PER_BLOCK };

static umma::BPlanes tc_planes(const SepEngine* e, int blk, int which, int ld) {
    umma::BPlanes b;
    b.base = e->planes + e->plane_of[(size_t)blk * PL_PER_BLOCK + which];
    b.ld = ld; b.plane_stride = e->planes_total; b.nz = 1;
    return b;
}

// the recurrence: tensor cores when there are enough sequences to fill the GPU with 32-sequence CTAs, else lstm.cuh
static cudaError_t lstm_any(SepEngine* e, const LstmArgs& l, cudaStream_t st, bool pdl) {
    if (e->use_tc && (int64_t)l.nseq * l.ndir >= e->tcl_min_seqdirs) return launch_tc_lstm(l, e->tc_passes, st, pdl);
    return launch_lstm_rec(l, st, pdl);
}

// ... and with enough sequences the input projection moves into the recurrence kernel too (tc_lstm_x_kernel)
static bool tc_fused_lstm(const SepEngine* e, const LstmArgs& l) {
    return e->use_tc && e->fuse_ih && (int64_t)l.nseq * l.ndir >= e->tcl_min_seqdirs;
}

// C[rows][N] = epi(LN?(A[rows][lda, first K]) W^T + bias) (+ R), plain row-major rows
static int tc_rows_gemm(SepEngine* e, int blk, int which, const float* A, int64_t lda, int K, int N, const float* ln_g, const float* ln_b,
                        const float* bias, const float* prelu_vec, const float* R, float* C, int64_t ldc, int64_t rows, cudaStream_t st,
                        const float* prelu_scalar = nullptr) {
    umma::GemmDesc g;
    g.a0.base = A; g.a0.channels = K; g.a0.n_pos = rows; g.a0.pos_stride = lda;
    umma::set_plain_chunks(g, K, ln_g != nullptr);
    g.ln_g = ln_g; g.ln_b = ln_b;
    g.rows_per_seq = (int)rows; g.nseq = 1;
    g.b = tc_planes(e, blk, which, K); g.N = N; g.K = K; g.passes = e->tc_passes;
    g.bias = bias; g.prelu_vec = prelu_vec; g.prelu = prelu_scalar; g.R = R; g.C = C; g.ldc = ldc;
    g.pdl = e->cur_pdl;
    std::string why;
    const cudaError_t ce = umma::launch(g, st, &why);
    if (ce != cudaSuccess) return fail(3, std::string("umma_gemm: ") + cudaGetErrorString(ce) + " " + why);
    return 0;
}

struct Profiler {                      // per-kernel device times via CUDA events on the launching stream
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    int used = 0;
    int mark(const char* name, cudaStream_t st) {
        if (used == (int)ev.size()) { cudaEvent_t e; CK(cudaEventCreate(&e)); ev.push_back(e); }
        CK(cudaEventRecord(ev[used++], st));
        names.push_back(name);
        return 0;
    }
};

struct ChainArgs {
    const float* x; int64_t xbs, xcs; int x_len;
    const float* emb; float* state;
    float* y; int64_t ybs, ycs; int y_len;
    int B, T; float* wsp; size_t ws_bytes; uint32_t flags; int pos_rel;
    Profiler* prof = nullptr;
};

static int enqueue_chain(SepEngine* e, const ChainArgs& a, cudaStream_t st) {
    const float* x = a.x; const int64_t xbs = a.xbs, xcs = a.xcs; const int x_len = a.x_len;
    const float* emb = a.emb; float* state = a.state; float* y = a.y;
    const int64_t ybs = a.ybs, ycs = a.ycs; const int y_len = a.y_len, B = a.B, T = a.T;
    float* wsp = a.wsp; const size_t ws_bytes = a.ws_bytes; const uint32_t flags = a.flags;
    if (!e->committed) return fail(4, "weights not committed");
    if (B <= 0 || T <= 0) return fail(1, "batch and frames must be positive");
    const Workspace ws = carve(e->n_blocks, B, T, flags);
    if ((size_t)ws.total * sizeof(float) > ws_bytes) return fail(1, "workspace too small");
    if (int rc = set_attrs()) return rc;
    const int64_t ss = stream_stride(e->n_blocks);
    const int64_t rows = (int64_t)B * T * NF;
    if (rows > 0x7fffffff / 2) return fail(1, "batch*frames too large for one call; split the batch");
    float* X = wsp + ws.X; float* GX = wsp + ws.GX; float* Y = wsp + ws.Y; float* Z = wsp + ws.Z;
    float* Q = wsp + ws.Q; float* KALL = wsp + ws.KALL; float* VALL = wsp + ws.VALL; float* PRE = wsp + ws.PRE;
    float* TAPS = wsp + ws.TAPS;
    float* QKVRAW = wsp + ws.QKVRAW;
    const int nsplit = attn_splits(B, T);
    // one-frame calls: the row-local middle of every block runs as ONE fused kernel (mid_kernel.cuh).
    // (Taps want the intermediate activations of the generic chain, so they keep it.)
    const bool fused_mid = (T == 1) && e->use_mid && !(flags & L2H_FLAG_TAPS);
    // many rows (whole utterances, offline batches, many streams): the dense contractions run on the tensor cores
    const bool tc = e->use_tc && rows > TC_MIN_ROWS;
    const bool tc_mid = tc && T == 1 && !(flags & L2H_FLAG_TAPS);
    // few streams, one hop (the latency path): everything after the BiLSTM as one cluster kernel per stream
    bool fused_tail = false;
    if (fused_mid && !tc && e->use_tail && !e->fold_mid_c) {
        int dev_ord = 0;
        CK(cudaGetDevice(&dev_ord));
        fused_tail = B <= g_tail_clusters[dev_ord];      // all clusters of the launch resident at once
    }
    int tap = 0;
    auto do_tap = [&]() -> int {
        if (flags & L2H_FLAG_TAPS) {
            CK(cudaMemcpyAsync(TAPS + (int64_t)tap * rows * 64, X, rows * 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
            ++tap;
        }
        return 0;
    };
    float* sbase = state + sizeof(StateHeader) / 4;
    // programmatic dependent launch pays on the latency chain of a few rows; with many rows it is slower (256 streams: 0.939 vs
    // 0.865 ms per hop-step, profiles/r02j_b256_options.jsonl): early-launched dependents park on the SMs the big kernels need
    const bool pdl = e->use_pdl && a.prof == nullptr && !(flags & L2H_FLAG_TAPS) && !(e->use_tc && rows > TC_MIN_ROWS);
    e->cur_pdl = false;
#define MARK(name) do { if (a.prof) { if (int _rc = a.prof->mark(name, st)) return _rc; } } while (0)
    MARK("start");
    if (fused_tail) {      // the frame as 13 row tiles: spectrum of the tile's bins, conv, and block 0's input projection
        CK(launch_k(false, front1_kernel, dim3(TAIL_TILES + 1, B), dim3(256), FRONT1_SMEM, st, x, xbs, xcs, x_len, X, state, ss, e->w,
                    e->bw[0], GX, a.pos_rel, emb, PRE));
    } else if ((T > 1 || tc) && e->use_back_many) {      // many frames / streams: one CTA walks (stream, chunk) items (one CTA per SM: 150 KB of filters)
        const int per_stream = std::max(1, 148 / B);
        const int chunk = (T + per_stream - 1) / per_stream;
        const int n_chunks = (T + chunk - 1) / chunk;
        const int n_workers = std::min(148, B * n_chunks);
        CK(launch_k(false, front_many_kernel, dim3(n_workers + B, 1), dim3(256), FRONT_SMEM, st, x, xbs, xcs, x_len, X, state, ss, e->w, T,
                    a.pos_rel, emb, PRE, chunk, n_chunks, B, n_workers));
    } else {
        CK(launch_k(false, front_kernel, dim3(T + 1, B), dim3(256), FRONT_SMEM, st, x, xbs, xcs, x_len, X, state, ss, e->w, T,
                    a.pos_rel, emb, PRE, 0, 1, 0));
    }
    MARK("front");
    if (int rc = do_tap()) return rc;

    for (int b = 0; b < e->n_blocks; ++b) {
        const BlockWeights& W = e->bw[b];
        // ---- intra: LN -> W_ih (both directions) -> BiLSTM over F -> Linear -> +res ------------
        GemmArgs g{};
        LstmArgs l{};
        l.gx = GX; l.gx_ld = 512; l.out = Y; l.out_ld = 128; l.whh = W.whh1;
        l.nseq = B * T; l.L = NF; l.inner_count = 1; l.outer_stride = NF; l.inner_stride = 0; l.step_stride = 1;
        l.ndir = 2;
        if (tc_fused_lstm(e, l)) {
            // many sequences: LayerNorm, W_ih and the recurrence in ONE tensor-core kernel (no [rows x 512] projection in HBM)
            tcl::LstmXArgs xa{};
            xa.l = l; xa.x = X; xa.x_ld = 64;
            xa.wih_hi = e->planes + e->plane_of[(size_t)b * PL_PER_BLOCK + PL_IH1]; xa.wih_lo = xa.wih_hi + e->planes_total;
            xa.bias = W.b1; xa.ln_g = W.ln1_g; xa.ln_b = W.ln1_b;
            CK(launch_tc_lstm_x(xa, e->tc_passes, st, false));
            MARK("gemm_ih_intra");
        } else {
            if (fused_tail) {
                // front1_kernel (block 0) / the previous block's tail_kernel (its phase G) already wrote this block's input projection
            } else if (tc) {
                e->cur_pdl = (e->tc_pdl & 2) != 0 && a.prof == nullptr;
                if (int rc = tc_rows_gemm(e, b, PL_IH1, X, 64, 64, 512, W.ln1_g, W.ln1_b, W.b1, nullptr, nullptr, GX, 512, rows, st)) return rc;
                e->cur_pdl = false;
            } else {
                g.A = X; g.lda = 64; g.Wt = W.wih1_t; g.bias = W.b1; g.C = GX; g.ldc = 512;
                g.ln_g = W.ln1_g; g.ln_b = W.ln1_b; g.M = (int)rows; g.N = 512; g.K = 64;
                CK(launch_rows_gemm(g, st, pdl));
            }
            MARK("gemm_ih_intra");
            CK(lstm_any(e, l, st, pdl));
        }
        MARK("lstm_intra");
        if (tc_mid) {
            // many streams, one hop: the row-local middle of the block as four tensor-core GEMMs and the cell update.
            // The inter-LSTM step is ONE GEMM over the concatenated k = [LN(x) | h_prev] (h read in place from the
            // per-stream state records through a strided tensor map) against [W_ih | W_hh].
            const bool mp = (e->tc_pdl & 1) != 0 && a.prof == nullptr;      // programmatic launches inside the section (experiment, option "tc_pdl")
            e->cur_pdl = mp;
            if (int rc = tc_rows_gemm(e, b, PL_L1, Y, 128, 128, 64, nullptr, nullptr, W.bl1, nullptr, X, X, 64, rows, st)) return rc;
            {
                umma::GemmDesc q;
                q.a0.base = X; q.a0.channels = 64; q.a0.n_pos = NF; q.a0.pos_stride = 64; q.a0.n_inner = B; q.a0.inner_stride = (int64_t)NF * 64;
                q.a1.base = sbase + ST_BLK + (int64_t)b * BK_STRIDE + BK_H;
                q.a1.channels = 64; q.a1.n_pos = NF; q.a1.pos_stride = 64; q.a1.n_inner = B; q.a1.inner_stride = ss;
                q.n_chunks = 2;
                q.chunks[0].c0 = 0; q.chunks[0].dp = 0; q.chunks[0].flags = 2;      // x: LayerNorm
                q.chunks[1].c0 = 0; q.chunks[1].dp = 0; q.chunks[1].flags = 1;      // h: second source
                q.ln_g = W.ln2_g; q.ln_b = W.ln2_b;
                q.rows_per_seq = NF; q.nseq = B;
                q.b = tc_planes(e, b, PL_CAT, 128); q.N = 256; q.K = 128; q.passes = e->tc_passes;
                q.bias = W.b2; q.C = GX; q.ldc = 256; q.c_seq_stride = (int64_t)NF * 256;
                q.pdl = mp;
                std::string why;
                const cudaError_t ce = umma::launch(q, st, &why);
                if (ce != cudaSuccess) return fail(3, std::string("umma_gemm (inter step): ") + cudaGetErrorString(ce) + " " + why);
            }
            CK(launch_k(pdl || mp, lstm_cell_rows_kernel, dim3((unsigned)((rows * 64 + 255) / 256)), dim3(256), 0, st, (const float*)GX, state, ss, b, Y, (int)rows));
            if (int rc = tc_rows_gemm(e, b, PL_L2, Y, 64, 64, 64, nullptr, nullptr, W.bl2, nullptr, X, X, 64, rows, st)) return rc;
            if (int rc = tc_rows_gemm(e, b, PL_QKV, X, 64, 64, NQKV, nullptr, nullptr, W.bqkv, W.slope_vec, nullptr, QKVRAW, NQKV, rows, st)) return rc;
            e->cur_pdl = false;
            MARK("mid");
        } else if (fused_mid && mid_split_for_throughput(B) && e->mid_split_large) {
            float* GI = GX; float* HN = GX + rows * 256;         // the BiLSTM is done with GX
            CK(launch_k(pdl, mid_a_kernel, mid_grid_for(B, 2), dim3(256), MID_A_SMEM, st, (const float*)Y, X, GI, W, B, (int64_t)0, 1));
            if (e->fold_mid_c) {
                CK(launch_k(pdl, mid_b2_kernel, mid_grid_for(B, 2), dim3(256), MID_B2_SMEM, st, (const float*)GI, X, (int64_t)0, 1, state, ss, b, W, B));
            } else {
                CK(launch_k(pdl, mid_b_kernel, mid_grid_for(B, 3), dim3(256), MID_B_SMEM, st, (const float*)GI, HN, (int64_t)0, 1, state, ss, b, W, B));
                CK(launch_k(pdl, mid_c_kernel, mid_grid_for(B, 4), dim3(256), MID_C_SMEM, st, (const float*)HN, X, QKVRAW, W, B, (int64_t)0, 1));
            }
            MARK("mid");
        } else if (fused_tail) {
            NextIh nx{};
            if (b + 1 < e->n_blocks) {
                const BlockWeights& Wn = e->bw[b + 1];
                nx.ln_g = Wn.ln1_g; nx.ln_b = Wn.ln1_b; nx.wih_t = Wn.wih1_t; nx.bias = Wn.b1; nx.GX = GX;
            }
            CK(launch_cluster(pdl, dim3(TAIL_CL, 1, 1), tail_kernel, dim3(TAIL_CL, B), dim3(256), TAIL_SMEM, st, (const float*)Y, X, state, ss,
                              b, W, nx, (b == 0 && e->n_blocks > 1) ? 1 : 0, 0));
            MARK("tail");
            if (int rc = do_tap()) return rc;
            continue;
        } else if (fused_mid) {
            if (e->fold_mid_c)
                CK(launch_k(pdl, mid_noproj_kernel, mid_grid_for(B, 1), dim3(256), MID_SMEM, st, (const float*)Y, X, QKVRAW, state, ss, b, W, B));
            else
                CK(launch_k(pdl, mid_kernel, mid_grid_for(B, 1), dim3(256), MID_SMEM, st, (const float*)Y, X, QKVRAW, state, ss, b, W, B));
            MARK("mid");
        } else {
            if (tc) {
                if (int rc = tc_rows_gemm(e, b, PL_L1, Y, 128, 128, 64, nullptr, nullptr, W.bl1, nullptr, X, X, 64, rows, st)) return rc;
            } else {
                g = GemmArgs{};
                g.A = Y; g.lda = 128; g.Wt = W.wl1_t; g.bias = W.bl1; g.C = X; g.ldc = 64; g.R = X;
                g.M = (int)rows; g.N = 64; g.K = 128;
                CK(launch_rows_gemm(g, st, pdl));
            }
            MARK("gemm_lin_intra");
            if (int rc = do_tap()) return rc;
            // ---- inter: LN -> W_ih -> LSTM over T with carried (h, c) -> Linear -> +res ------------
            l = LstmArgs{};
            l.gx = GX; l.gx_ld = 256; l.out = Y; l.out_ld = 64; l.whh = W.whh2;
            l.h_state = sbase + ST_BLK + (int64_t)b * BK_STRIDE + BK_H;
            l.c_state = sbase + ST_BLK + (int64_t)b * BK_STRIDE + BK_C;
            l.hc_outer_stride = ss;
            l.nseq = B * NF; l.L = T; l.inner_count = NF; l.outer_stride = (int64_t)T * NF; l.inner_stride = 1;
            l.step_stride = NF; l.ndir = 1;
            if (tc_fused_lstm(e, l)) {
                tcl::LstmXArgs xa{};
                xa.l = l; xa.x = X; xa.x_ld = 64;
                xa.wih_hi = e->planes + e->plane_of[(size_t)b * PL_PER_BLOCK + PL_IH2]; xa.wih_lo = xa.wih_hi + e->planes_total;
                xa.bias = W.b2; xa.ln_g = W.ln2_g; xa.ln_b = W.ln2_b;
                CK(launch_tc_lstm_x(xa, e->tc_passes, st, false));
                MARK("gemm_ih_inter");
            } else {
                if (tc) {
                    if (int rc = tc_rows_gemm(e, b, PL_IH2, X, 64, 64, 256, W.ln2_g, W.ln2_b, W.b2, nullptr, nullptr, GX, 256, rows, st)) return rc;
                } else {
                    g = GemmArgs{};
                    g.A = X; g.lda = 64; g.Wt = W.wih2_t; g.bias = W.b2; g.C = GX; g.ldc = 256;
                    g.ln_g = W.ln2_g; g.ln_b = W.ln2_b; g.M = (int)rows; g.N = 256; g.K = 64;
                    CK(launch_rows_gemm(g, st, pdl));
                }
                MARK("gemm_ih_inter");
                CK(lstm_any(e, l, st, pdl));
            }
            MARK("lstm_inter");
            if (tc) {
                if (int rc = tc_rows_gemm(e, b, PL_L2, Y, 64, 64, 64, nullptr, nullptr, W.bl2, nullptr, X, X, 64, rows, st)) return rc;
            } else {
                g = GemmArgs{};
                g.A = Y; g.lda = 64; g.Wt = W.wl2_t; g.bias = W.bl2; g.C = X; g.ldc = 64; g.R = X;
                g.M = (int)rows; g.N = 64; g.K = 64;
                CK(launch_rows_gemm(g, st, pdl));
            }
            MARK("gemm_lin_inter");
            if (int rc = do_tap()) return rc;
        }
        // ---- attention --------------------------------------------------------------------------
        if (T > 1) {
            CK(launch_k(pdl, kv_gather_kernel, dim3(ATT - 1, B * NHEAD), dim3(128), 0, st, (const float*)state, ss, b, KALL,
                        VALL, T));
            MARK("kv_gather");
        }
        if (tc && !tc_mid) {      // Q|K|V projections of all rows as one tensor-core GEMM (+ bias + PReLU per column)
            if (int rc = tc_rows_gemm(e, b, PL_QKV, X, 64, 64, NQKV, nullptr, nullptr, W.bqkv, W.slope_vec, nullptr, QKVRAW, NQKV, rows, st)) return rc;
        }
        if (tc && (int64_t)B * T >= 148) {      // many frames: persistent form (LayerNorm parameters staged once per CTA, two CTAs per SM)
            CK(launch_k(pdl || ((e->tc_pdl & 4) != 0 && a.prof == nullptr), qkv_many_kernel, dim3((unsigned)std::min<int64_t>(296, (int64_t)B * T)), dim3(QKV_THREADS), QKV_MANY_SMEM, st, (const float*)QKVRAW, Q, KALL, VALL, state, ss,
                        b, W, T, B * T));
        } else {
            CK(launch_k(pdl, qkv_kernel, dim3(T, B), dim3(QKV_THREADS), QKV_SMEM, st, (const float*)X,
                        (const float*)((tc || (fused_mid && !e->fold_mid_c)) ? QKVRAW : nullptr), Q, KALL, VALL, state, ss, b, W, T, 0));
        }
        MARK("qkv");
        if (nsplit > 1) {
            CK(launch_cluster(pdl, dim3(1, ATT_CL, 1), attn_cluster_kernel, dim3(T, NHEAD * ATT_CL, B), dim3(256), 0, st,
                              (const float*)Q, (const float*)KALL, (const float*)VALL, (const float*)state, ss, b, Z, T, 0));
        } else if (T > 1 && (int64_t)B * NHEAD * ((T + ATT_TQ - 1) / ATT_TQ) >= 148) {   // enough tiles to fill the GPU: query-tiled,
            // one pass over 57 rows serves 8 queries
            CK(launch_k(pdl, attn_tile_kernel, dim3((T + ATT_TQ - 1) / ATT_TQ, NHEAD, B), dim3(256), 0, st, (const float*)Q,
                        (const float*)KALL, (const float*)VALL, Z, T));
        } else {
            CK(launch_k(pdl, attn_kernel, dim3(T, NHEAD, B), dim3(256), 0, st, (const float*)Q, (const float*)KALL,
                        (const float*)VALL, (const float*)state, ss, b, Z, T, 0));
        }
        MARK("attn");
        if (tc) {      // Linear(64->64) + PReLU of all rows on the tensor cores, then LayerNorm(6208) + residual (+ gate) per frame
            const bool p2 = (e->tc_pdl & 2) != 0 && a.prof == nullptr;
            e->cur_pdl = p2;
            if (int rc = tc_rows_gemm(e, b, PL_P, Z, 64, 64, 64, nullptr, nullptr, W.bp, nullptr, nullptr, Y, 64, rows, st, W.slopes + 3)) return rc;
            e->cur_pdl = false;
            CK(launch_k(pdl || p2, ln_frame_res_kernel, dim3(T, B), dim3(256), 0, st, (const float*)Y, X, (const float*)state, ss, W,
                        (b == 0 && e->n_blocks > 1) ? 1 : 0, T));
        } else {
            CK(launch_k(pdl, attn_out_kernel, dim3(T, B), dim3(256), AOUT_SMEM, st, (const float*)Z, X, (const float*)state, ss, W,
                        (b == 0 && e->n_blocks > 1) ? 1 : 0, T));
        }
        MARK("attn_out");
        if (int rc = do_tap()) return rc;
    }
    if ((T > 1 || tc) && e->use_back_many) {
        // many frames (or many streams): one cluster walks (stream, chunk-of-frames) items -- filters loaded once per cluster, rows
        // staged once; as many clusters as stay resident together (3 CTAs of 72 KB per SM)
        const int max_cl = 148 * 3 / BACK_CL;
        const int per_stream = std::max(1, max_cl / B);
        const int chunk = (T + per_stream - 1) / per_stream;
        const int n_chunks = (T + chunk - 1) / chunk;
        const int n_cl = std::min(max_cl, B * n_chunks);
        CK(launch_cluster(pdl, dim3(BACK_CL, 1, 1), back_many_kernel, dim3(BACK_CL * n_cl, 1), dim3(256), BACK_MANY_SMEM, st, (const float*)X, y, ybs,
                          ycs, y_len, state, ss, e->w, T, a.pos_rel, chunk, n_chunks, B));
    } else {
        CK(launch_cluster(pdl, dim3(BACK_CL, 1, 1), back_kernel, dim3(BACK_CL * T, B), dim3(256), BACK_SMEM, st, (const float*)X, y, ybs, ycs, y_len, state, ss, e->w, T,
                    a.pos_rel, 0, 1, 0, (int64_t)0));
    }
    MARK("back");
#undef MARK
    return 0;
}

// ---- wavefront pipeline over (block, frame) for one-frame calls ---------------------------------------
// Work item (block b, hop t) depends only on (b-1, t) and (b, t-1), and inside a block
// only part of the work carries state from hop to hop:
//   A   = W_ih GEMM + 97-step BiLSTM        needs X_t only            -> PIPE_LANES hops of it run side by side
//   B1  = mid_kernel (inter-LSTM step ...)   carries (h, c)            -> serial per block
//   B2a = qkv + attention                    carries the K/V rings     -> serial per block
//   B2b = attn_out                           needs Z_t, X_t only
// A graph of K consecutive one-hop chains is captured on 1 + 3*(PIPE_LANES+3) + 1 streams with event edges
// for exactly these dependencies; hops flow through the stages like a systolic wavefront and the
// steady-state cost per hop is the slowest SERIAL stage instead of the whole 250 us chain.  Every hop owns a
// workspace slot; state addressing uses pos + frame_k / parity(ncalls + frame_k); the header advances once,
// at the last hop of the graph.  The arithmetic and its order per stream are unchanged: results are
// bit-identical to running the hops one after the other (tests/test_sep_gpu.py).
constexpr int PIPE_MAX_FRAMES = 500;
constexpr int PIPE_LANES = 16;     // max hops of stage A (BiLSTM) in flight per block (engine->pipe_alanes used)
constexpr int PIPE_FLANES = 8;     // max front_kernel lanes (frames of a group do not depend on each other there)
constexpr int PIPE_BLANES = 6;     // max back_kernel lanes
constexpr int PIPE_BASE = 1 + PIPE_FLANES + PIPE_BLANES;
constexpr int PIPE_QLANES = 4;     // max qkv lanes (hops write different ring rows; RING - ATT = 6 may run ahead of the attention)
constexpr int PIPE_CLANES = 3;     // max mid_c lanes (no hop-to-hop dependency, not bound by the ring guard)
constexpr int PIPE_TLANES = 4;     // max attention lanes (attention only reads the rings)
constexpr int PIPE_OLANES = 4;     // max attn_out lanes (no hop-to-hop dependency)
constexpr int PIPE_PER_BLOCK = PIPE_LANES + 1 + PIPE_QLANES + PIPE_TLANES + PIPE_OLANES + PIPE_CLANES;   // A lanes, B1 (mid), Bq lanes (qkv), Ba lanes (attention), Bo lanes (attn_out)
constexpr int PIPE_STREAMS = PIPE_BASE + 3 * PIPE_PER_BLOCK;
constexpr int PIPE_MIDB_MAX = 8;   // max hops per mid_b launch
constexpr int PIPE_QKV_AHEAD = RING - ATT;         // qkv of hop t+3 overwrites a ring row hop t's attention still reads
static_assert(PIPE_STREAMS <= 128, "pipe_streams[]");

static int64_t pipe_slot_floats(SepEngine* e, int B) { return carve(e->n_blocks, B, 1, 0).total; }
// hops per pipelined graph: fill + drain cost one chain latency (~0.25 ms) per graph, so as many as possible -- every hop
// in flight owns a workspace slot (350 KB per stream), which is what bounds it for many streams
static int pipe_frames_for(SepEngine* e, int B) {
    if (e->pipe_frames > 0) return e->pipe_frames;
    // every hop in flight owns a workspace slot: at most 24 GB of them, and never more than half of what the device
    // has free right now (the state, the caller's buffers and other tenants need room too)
    if (e->pipe_budget == 0) {               // asked once per handle: cudaMemGetInfo costs tens of microseconds per call
        e->pipe_budget = (int64_t)24 << 30;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > 0) e->pipe_budget = std::min<int64_t>(e->pipe_budget, (int64_t)(free_b / 2));
    }
    const int64_t fit = e->pipe_budget / (pipe_slot_floats(e, B) * (int64_t)sizeof(float));
    return (int)std::max<int64_t>(2, std::min<int64_t>(PIPE_MAX_FRAMES, fit));
}
static int enqueue_pipeline(SepEngine* e, const ChainArgs& a, int K, cudaStream_t origin) {
    if (e->n_blocks != 3) return fail(1, "pipeline graph is specialised to 3 blocks");
    const int B = a.B;
    const Workspace ws = carve(e->n_blocks, B, 1, 0);
    const int64_t slot = ws.total;
    if ((size_t)slot * K * sizeof(float) > a.ws_bytes) return fail(1, "workspace too small for the pipelined stream");
    const int64_t ss = stream_stride(e->n_blocks);
    const int rows = B * NF;
    const int nsplit = attn_splits(B, 1);
    // programmatic dependent launch, per stage: the kernel's prologue (weight staging) overlaps its stream predecessor's
    // tail; every kernel reaches griddepcontrol.wait before it touches activations or state.  Worth it only on the
    // serial stage (mid_b -> mid_b of the next hop): everywhere, the parked dependents hold shared memory and CTA
    // slots the running kernels need (measured 14.0 vs 9.4 us per hop, profiles/r01f_pipeline_sweeps.jsonl)
    const int ppdl = e->pipe_pdl;         // stage bit mask (same bits as pipe_skip)
    const bool fold = e->fold_mid_c;
    const bool many = mid_split_for_throughput(B);
    const bool split_mid = many ? e->mid_split_large : e->pipe_split_mid;
    float* state = a.state;
    for (int i = 0; i < PIPE_STREAMS; ++i)
        if (!e->pipe_streams[i]) CK(cudaStreamCreateWithFlags(&e->pipe_streams[i], cudaStreamNonBlocking));
    size_t ev_used = 0;
    auto next_event = [&](cudaEvent_t* out) -> int {
        if (ev_used == e->pipe_events.size()) {
            cudaEvent_t ev;
            CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            e->pipe_events.push_back(ev);
        }
        *out = e->pipe_events[ev_used++];
        return 0;
    };
    auto edge = [&](cudaStream_t from, cudaStream_t to) -> int {     // `to` continues after everything enqueued on `from`
        if (from == to) return 0;
        cudaEvent_t ev;
        if (int rc = next_event(&ev)) return rc;
        CK(cudaEventRecord(ev, from));
        CK(cudaStreamWaitEvent(to, ev, 0));
        return 0;
    };
    // stream map: [0] capture origin (fork / join / header advance), front lanes, back lanes, then per block:
    // PIPE_LANES x A, B1, B2a, B2b
    auto sFront = [&](int k) { return e->pipe_streams[1 + k % e->pipe_flanes]; };
    auto sBackL = [&](int k) { return e->pipe_streams[1 + PIPE_FLANES + k % e->pipe_blanes]; };
    auto sA = [&](int b, int lane) { return e->pipe_streams[PIPE_BASE + b * PIPE_PER_BLOCK + lane]; };
    auto sB1 = [&](int b) { return e->pipe_streams[PIPE_BASE + b * PIPE_PER_BLOCK + PIPE_LANES]; };
    auto sBq = [&](int b, int k) { return e->pipe_streams[PIPE_BASE + b * PIPE_PER_BLOCK + PIPE_LANES + 1 + k % e->pipe_qlanes]; };
    auto sBa = [&](int b, int k) { return e->pipe_streams[PIPE_BASE + b * PIPE_PER_BLOCK + PIPE_LANES + 1 + PIPE_QLANES + k % e->pipe_tlanes]; };
    auto sBo = [&](int b, int k) {
        return e->pipe_streams[PIPE_BASE + b * PIPE_PER_BLOCK + PIPE_LANES + 1 + PIPE_QLANES + PIPE_TLANES + k % e->pipe_olanes];
    };
    auto sBc = [&](int b, int k) {
        return e->pipe_streams[PIPE_BASE + b * PIPE_PER_BLOCK + PIPE_LANES + 1 + PIPE_QLANES + PIPE_TLANES + PIPE_OLANES + k % e->pipe_clanes];
    };
    // events: stage A (BiLSTM [+ mid_a]) done, qkv done, attention done, attn_out done -- per block and hop
    std::vector<std::vector<cudaEvent_t>> a_done(3, std::vector<cudaEvent_t>(K, nullptr));
    std::vector<std::vector<cudaEvent_t>> qkv_done(3, std::vector<cudaEvent_t>(K, nullptr));
    std::vector<std::vector<cudaEvent_t>> att_done(3, std::vector<cudaEvent_t>(K, nullptr));
    std::vector<std::vector<cudaEvent_t>> out_done(3, std::vector<cudaEvent_t>(K, nullptr));
    auto record = [&](cudaEvent_t* ev, cudaStream_t s) -> int {
        if (int rc = next_event(ev)) return rc;
        CK(cudaEventRecord(*ev, s));
        return 0;
    };
    for (int i = 1; i < PIPE_STREAMS; ++i)                             // fork: bring the worker streams into the capture
        if (int rc = edge(origin, e->pipe_streams[i])) return rc;
    // The serial stage (mid_b) takes `mb` consecutive hops per launch: its ~3 us of launch overhead and its 64 KB of
    // weights are paid once per batch, h stays in shared memory and c in registers from hop to hop.  A batch waits
    // for stage A of all its hops; the upstream block is that far ahead anyway once the pipeline is full.
    const int mb = (split_mid && !many) ? std::max(1, std::min(e->pipe_midb_hops, PIPE_MIDB_MAX)) : 1;
    float* PRE = a.wsp + ws.PRE;                                       // speaker-gate scratch: front stream only
    for (int k0 = 0; k0 < K; k0 += mb) {
        const int k1 = std::min(K, k0 + mb);                           // this batch: hops [k0, k1)
        for (int b = 0; b < 3; ++b) {
            const BlockWeights& W = e->bw[b];
            // ---- stage A of the batch: ONE launch each of W_ih GEMM, BiLSTM and mid_a for its hops [k0, k1) (the hops'
            // workspace slots are `slot` floats apart: strided rows / sequences / hop index inside the kernels) ------------
            {
                const int nh = k1 - k0;
                float* wsp = a.wsp + (int64_t)k0 * slot;
                float* X = wsp + ws.X; float* GX = wsp + ws.GX; float* Y = wsp + ws.Y;
                cudaStream_t st_a = sA(b, (k0 / mb) % e->pipe_alanes);
                for (int k = k0; k < k1; ++k) {
                    if (b == 0) {
                        // x / y are the group's buffers; hop k works at sample offset k*128 (plus, with pos_rel, the clip
                        // position the device derives from the state header)
                        cudaStream_t sF = sFront(k);
                        // the speaker-gate memo CTA (blockIdx.x == 1) rides with hop 0 only: one builder of ST_GATE per group
                        if (!(e->pipe_skip & 1)) CK(launch_k((ppdl & 1) != 0, front_kernel, dim3(k == 0 ? 2 : 1, B), dim3(256), FRONT_SMEM, sF, a.x, a.xbs, a.xcs,
                                                             a.x_len, a.wsp + (int64_t)k * slot + ws.X, state, ss, e->w, 1, a.pos_rel, a.emb, PRE, k, K, k * HOP));
                        if (k == 0) {                  // ... and every attn_out lane of block 0 (the gate's only reader) waits for it once
                            cudaEvent_t gate_ev;
                            if (int rc = record(&gate_ev, sF)) return rc;
                            for (int ln = 0; ln < e->pipe_olanes; ++ln) CK(cudaStreamWaitEvent(sBo(0, ln), gate_ev, 0));
                        }
                        if (int rc = edge(sF, st_a)) return rc;
                    } else {
                        CK(cudaStreamWaitEvent(st_a, out_done[b - 1][k], 0));
                    }
                }
                GemmArgs g{};
                g.A = X; g.lda = 64; g.a_rows_per_seq = rows; g.a_seq_stride = slot;
                g.Wt = W.wih1_t; g.bias = W.b1; g.C = GX; g.ldc = 512; g.c_rows_per_seq = rows; g.c_seq_stride = slot;
                g.ln_g = W.ln1_g; g.ln_b = W.ln1_b; g.M = rows * nh; g.N = 512; g.K = 64;
                if (!(e->pipe_skip & 2)) CK(launch_rows_gemm(g, st_a, (ppdl & 2) != 0, e->pipe_gemm_shape));
                LstmArgs l{};
                l.gx = GX; l.gx_ld = 512; l.out = Y; l.out_ld = 128; l.whh = W.whh1;
                l.nseq = B * nh; l.L = NF; l.inner_count = B; l.outer_stride = slot / 512; l.inner_stride = NF; l.step_stride = 1;
                l.out_outer_stride = slot / 128; l.out_inner_stride = NF; l.out_step_stride = 1; l.ndir = 2;
                if (!(e->pipe_skip & 4)) CK(launch_lstm_rec(l, st_a, (ppdl & 4) != 0));
                // only the W_hh product + cell (mid_b) is serial per block; the rest rides on the parallel lanes.
                // GI / H' live in the hop's GX slot, which the BiLSTM has finished with.
                if (split_mid && !(e->pipe_skip & 8))
                    CK(launch_k((ppdl & 8) != 0, mid_a_kernel, mid_grid_for(B * nh, 2), dim3(256), MID_A_SMEM, st_a, (const float*)Y, X, GX, W, B, slot, nh));
                cudaEvent_t ev_a;
                if (int rc = record(&ev_a, st_a)) return rc;
                for (int k = k0; k < k1; ++k) a_done[b][k] = ev_a;
            }
            // ---- the serial stage: one launch for the batch ----------------------------------------------------
            CK(cudaStreamWaitEvent(sB1(b), a_done[b][k0], 0));
            {
                float* wsp = a.wsp + (int64_t)k0 * slot;
                float* GI = wsp + ws.GX; float* HN = GI + (int64_t)rows * 256;
                if (split_mid && fold) {
                    if (!(e->pipe_skip & 16))
                        CK(launch_k((ppdl & 16) != 0, mid_b2_kernel, mid_grid_for(B, 2), dim3(256), MID_B2_SMEM, sB1(b), (const float*)GI, wsp + ws.X,
                                    slot, k1 - k0, state, ss, b, W, B));
                } else if (split_mid) {
                    if (!(e->pipe_skip & 16))
                        CK(launch_k((ppdl & 16) != 0, mid_b_kernel, mid_grid_for(B, 3), dim3(256), MID_B_SMEM, sB1(b), (const float*)GI, HN, slot,
                                    k1 - k0, state, ss, b, W, B));
                } else if (!(e->pipe_skip & 16)) {
                    if (fold)
                        CK(launch_k((ppdl & 16) != 0, mid_noproj_kernel, mid_grid_for(B, 1), dim3(256), MID_SMEM, sB1(b), (const float*)(wsp + ws.Y),
                                    wsp + ws.X, wsp + ws.QKVRAW, state, ss, b, W, B));
                    else
                        CK(launch_k((ppdl & 16) != 0, mid_kernel, mid_grid_for(B, 1), dim3(256), MID_SMEM, sB1(b), (const float*)(wsp + ws.Y),
                                    wsp + ws.X, wsp + ws.QKVRAW, state, ss, b, W, B));
                }
            }
            cudaEvent_t midb_done;
            if (int rc = record(&midb_done, sB1(b))) return rc;
            // ---- mid_c for the whole batch (one launch), then per hop: qkv -> attention -> attn_out (lanes, ring guards) ----
            cudaEvent_t midc_done = midb_done;
            if (split_mid && !fold) {
                cudaStream_t st_c = sBc(b, k0 / mb);
                float* wsp0 = a.wsp + (int64_t)k0 * slot;
                CK(cudaStreamWaitEvent(st_c, midb_done, 0));
                if (!(e->pipe_skip & 32))
                    CK(launch_k((ppdl & 32) != 0, mid_c_kernel, mid_grid_for(B * (k1 - k0), 4), dim3(256), MID_C_SMEM, st_c,
                                (const float*)(wsp0 + ws.GX + (int64_t)rows * 256), wsp0 + ws.X, wsp0 + ws.QKVRAW, W, B, slot, k1 - k0));
                if (int rc = record(&midc_done, st_c)) return rc;
            }
            for (int k = k0; k < k1; ++k) {
                float* wsp = a.wsp + (int64_t)k * slot;
                float* X = wsp + ws.X; float* Z = wsp + ws.Z; float* Q = wsp + ws.Q; float* QKVRAW = wsp + ws.QKVRAW;
                cudaStream_t st_q = sBq(b, k);
                CK(cudaStreamWaitEvent(st_q, midc_done, 0));
                // the ring row this hop's K/V overwrite was last read by the attention of hop k-3: it and every earlier
                // attention (one per attention lane) must be done
                for (int d = 0; d < e->pipe_tlanes && k - PIPE_QKV_AHEAD - 1 - d >= 0; ++d)
                    CK(cudaStreamWaitEvent(st_q, att_done[b][k - PIPE_QKV_AHEAD - 1 - d], 0));
                if (!(e->pipe_skip & 64)) CK(launch_k((ppdl & 64) != 0, qkv_kernel, dim3(1, B), dim3(QKV_THREADS), QKV_SMEM, st_q, (const float*)X,
                                                      (const float*)(fold ? nullptr : QKVRAW), Q, (float*)nullptr, (float*)nullptr, state, ss, b, W, 1, k));
                if (int rc = record(&qkv_done[b][k], st_q)) return rc;
                // the attention reads this hop's ring row and the 49 before it: the other qkv lanes' latest hops must be in
                cudaStream_t st_t = sBa(b, k);
                for (int d = 0; d < e->pipe_qlanes && d <= k; ++d) CK(cudaStreamWaitEvent(st_t, qkv_done[b][k - d], 0));
                if (nsplit > 1) {
                    if (!(e->pipe_skip & 128)) CK(launch_cluster((ppdl & 128) != 0, dim3(1, ATT_CL, 1), attn_cluster_kernel, dim3(1, NHEAD * ATT_CL, B),
                                                                 dim3(256), 0, st_t, (const float*)Q, (const float*)nullptr, (const float*)nullptr,
                                                                 (const float*)state, ss, b, Z, 1, k));
                } else {
                    if (!(e->pipe_skip & 128)) CK(launch_k((ppdl & 128) != 0, attn_kernel, dim3(1, NHEAD, B), dim3(256), 0, st_t, (const float*)Q,
                                                           (const float*)nullptr, (const float*)nullptr, (const float*)state, ss, b, Z, 1, k));
                }
                if (int rc = record(&att_done[b][k], st_t)) return rc;
                CK(cudaStreamWaitEvent(sBo(b, k), att_done[b][k], 0));
                if (!(e->pipe_skip & 256)) CK(launch_k((ppdl & 256) != 0, attn_out_kernel, dim3(1, B), dim3(256), AOUT_SMEM, sBo(b, k), (const float*)Z, X,
                                                       (const float*)state, ss, W, b == 0 ? 1 : 0, 1));
                if (int rc = record(&out_done[b][k], sBo(b, k))) return rc;
            }
        }
        for (int k = k0; k < k1; ++k) {
            float* X = a.wsp + (int64_t)k * slot + ws.X;
            cudaStream_t sBack = sBackL(k);
            // this hop's output and the three before it (deconv / overlap-add context) sit on different attn_out lanes
            for (int d = 0; d <= 3 && d <= k; ++d) CK(cudaStreamWaitEvent(sBack, out_done[2][k - d], 0));
            if (!(e->pipe_skip & 512)) CK(launch_cluster((ppdl & 512) != 0, dim3(BACK_CL, 1, 1), back_kernel, dim3(BACK_CL, B), dim3(256), BACK_SMEM, sBack,
                                                         (const float*)X, a.y, a.ybs, a.ycs, a.y_len, state, ss, e->w, 1, a.pos_rel, k, K, k * HOP, slot));
        }
    }
    for (int i = 1; i < PIPE_STREAMS; ++i)                             // join
        if (int rc = edge(e->pipe_streams[i], origin)) return rc;
    advance_header_kernel<<<1, 1, 0, origin>>>(state, K);                // pos += K, ncalls += 1: after every hop of the group
    CK(cudaGetLastError());
    return 0;
}

// kernel nodes of a captured graph (for the launch counter)
static int count_kernel_nodes(cudaGraph_t graph) {
    size_t n = 0;
    if (cudaGraphGetNodes(graph, nullptr, &n) != cudaSuccess || n == 0) return 0;
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(graph, nodes.data(), &n) != cudaSuccess) return 0;
    int k = 0;
    for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) ++k;
    }
    return k;
}

// K chained one-frame calls as one pipelined graph (cached on the argument set + K)
static int run_pipeline(SepEngine* e, const ChainArgs& a, int K, cudaStream_t st) {
    std::vector<int64_t> key = {(int64_t)a.x, a.xbs, a.xcs, a.x_len, (int64_t)a.emb, (int64_t)a.state, (int64_t)a.y,
                                a.ybs, a.ycs, a.y_len, a.B, -K, (int64_t)a.wsp, (int64_t)a.flags, a.pos_rel};
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        if (!e->committed) return fail(4, "weights not committed");
        if (int rc = set_attrs()) return rc;
        if (!e->pipe_streams[0]) CK(cudaStreamCreateWithFlags(&e->pipe_streams[0], cudaStreamNonBlocking));
        if (e->graphs.size() >= 32) {
            for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
            e->graphs.clear();
            e->graph_kernels.clear();
        }
        cudaStream_t origin = e->pipe_streams[0];
        CK(cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_pipeline(e, a, K, origin);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(origin, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(3, std::string("cudaStreamEndCapture (pipeline): ") + cudaGetErrorString(ce));
        if (e->graph_stats) {                       // debug (option "graph_stats"): node / edge census of the captured pipeline graph
            size_t n_nodes = 0, n_edges = 0;
            cudaGraphGetNodes(graph, nullptr, &n_nodes);
            cudaGraphGetEdges_v2(graph, nullptr, nullptr, nullptr, &n_edges);
            std::vector<cudaGraphNode_t> from(n_edges), to(n_edges);
            std::vector<cudaGraphEdgeData> ed(n_edges);
            cudaGraphGetEdges_v2(graph, from.data(), to.data(), ed.data(), &n_edges);
            size_t prog = 0;
            for (const auto& d : ed) prog += (d.type == cudaGraphDependencyTypeProgrammatic) ? 1 : 0;
            fprintf(stderr, "[l2h] pipeline graph: hops %d, nodes %zu, edges %zu, programmatic edges %zu\n", K, n_nodes, n_edges, prog);
        }
        cudaGraphExec_t exec = nullptr;
        const int nk = count_kernel_nodes(graph);
        CK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        e->graph_kernels[exec] = nk;
        it = e->graphs.emplace(key, exec).first;
    }
    CK(cudaGraphLaunch(it->second, st));
    e->launch_count += e->graph_kernels[it->second];
    return 0;
}

// Launch the chain directly, or replay it from a cached CUDA graph (captured on a private
// stream the first time this exact argument set is seen; graph launches go to the caller's stream).
static int run_chain(SepEngine* e, const ChainArgs& a, cudaStream_t st, bool use_graph) {
    if (!use_graph || (a.flags & L2H_FLAG_TAPS)) {
        const long long before = g_launches;
        const int rc = enqueue_chain(e, a, st);
        e->launch_count += g_launches - before;
        return rc;
    }
    std::vector<int64_t> key = {(int64_t)a.x, a.xbs, a.xcs, a.x_len, (int64_t)a.emb, (int64_t)a.state, (int64_t)a.y,
                                a.ybs, a.ycs, a.y_len, a.B, a.T, (int64_t)a.wsp, (int64_t)a.flags, a.pos_rel};
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        if (!e->committed) return fail(4, "weights not committed");
        if (int rc = set_attrs()) return rc;
        if (!e->cap_stream) CK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
        if (e->graphs.size() >= 32) {
            for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
            e->graphs.clear();
            e->graph_kernels.clear();
        }
        CK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_chain(e, a, e->cap_stream);
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(3, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
        cudaGraphExec_t exec = nullptr;
        const int nk = count_kernel_nodes(graph);
        CK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        e->graph_kernels[exec] = nk;
        it = e->graphs.emplace(key, exec).first;
    }
    CK(cudaGraphLaunch(it->second, st));
    e->launch_count += e->graph_kernels[it->second];
    return 0;
}

}  // namespace l2h

using namespace l2h;

extern "C" {

int l2h_abi_version(void) { return L2H_ABI_VERSION; }
const char* l2h_last_error(void) { return g_err.c_str(); }

int l2h_sep_create(const l2h_sep_config* c, void** handle) {
    if (!c || !handle) return fail(1, "null argument");
    if (c->stft_chunk_size != HOP || c->stft_pad_size != LOOKAHEAD || c->embed_dim != SPK || c->num_ch != NMIC ||
        c->D != CH || c->L != NHEAD || c->I != 1 || c->J != 1 || c->H != HID || c->local_atten_len != ATT ||
        !c->use_attn || !c->lookahead || !c->chunk_causal || c->num_src != NSRC || c->B < 1 || c->B > 16)
        return fail(1, "unsupported configuration: the kernels are specialised to configs/tsh.json "
                       "(chunk 128, pad 64, embed 256, 2 ch, D 64, H 64, 4 heads, I=J=1, window 50, 2 src)");
    SepEngine* e = new SepEngine();
    e->cfg = *c;
    e->n_blocks = c->B;
    build_layout(e);
    *handle = e;
    return 0;
}

// everything the handle owns on its device: cached graphs, capture / pipeline streams, events, the weight buffer
static void release_device_resources(SepEngine* e) {
    int cur = -1;
    const bool sw = e->device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != e->device;
    if (sw) cudaSetDevice(e->device);
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
    e->graphs.clear();
    e->graph_kernels.clear();
    if (e->cap_stream) { cudaStreamDestroy(e->cap_stream); e->cap_stream = nullptr; }
    for (auto& ps : e->pipe_streams) if (ps) { cudaStreamDestroy(ps); ps = nullptr; }
    for (auto& ev : e->pipe_events) cudaEventDestroy(ev);
    e->pipe_events.clear();
    if (e->trace_dev) { cudaFree(e->trace_dev); e->trace_dev = nullptr; e->trace_cap = 0; }
    if (e->dev) { cudaFree(e->dev); e->dev = nullptr; }
    if (e->planes) { cudaFree(e->planes); e->planes = nullptr; }
    if (sw) cudaSetDevice(cur);
}

// a handle is bound to the device that was current at its last commit (one handle per device)
static int check_device(SepEngine* e) {
    int cur = -1;
    CK(cudaGetDevice(&cur));
    if (e->device >= 0 && cur != e->device)
        return fail(1, "this handle's weights live on device " + std::to_string(e->device) + " but device " + std::to_string(cur) +
                       " is current: commit the weights again with the new device current (Net.to(device) does), or use one handle per device");
    return 0;
}

int l2h_sep_destroy(void* handle) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return 0;
    release_device_resources(e);
    delete e;
    return 0;
}

int l2h_sep_load_weight(void* handle, const char* name, const float* data, int64_t numel) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !name || !data) return fail(1, "null argument");
    auto it = e->slots.find(name);
    if (it == e->slots.end()) return fail(2, std::string("unknown weight name: ") + name);
    Slot& s = it->second;
    if (numel != s.numel) return fail(1, std::string("wrong element count for ") + name);
    if (s.accumulate) s.raw.assign(data, data + numel);
    else s.repack(data, e->host.data());
    s.loaded = true;
    e->committed = false;
    return 0;
}

int l2h_sep_weights_expected(void* handle, int32_t* n_expected, int32_t* n_loaded) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return fail(1, "null handle");
    int n = 0;
    for (auto& kv : e->slots) n += kv.second.loaded ? 1 : 0;
    if (n_expected) *n_expected = (int)e->slots.size();
    if (n_loaded) *n_loaded = n;
    return 0;
}

int l2h_sep_weight_info(void* handle, int32_t index, const char** name, int64_t* numel) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return fail(1, "null handle");
    if (index < 0 || index >= (int32_t)e->slots.size()) return fail(1, "weight index out of range");
    auto it = e->slots.begin();
    std::advance(it, index);
    if (name) *name = it->first.c_str();             // owned by the handle, valid until l2h_sep_destroy
    if (numel) *numel = it->second.numel;
    return 0;
}

int l2h_sep_commit_weights(void* handle, void* stream) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return fail(1, "null handle");
    for (auto& kv : e->slots)
        if (!kv.second.loaded) return fail(4, "weight not loaded: " + kv.first);
    for (auto& kv : e->slots)
        if (kv.second.accumulate) std::fill(e->host.begin() + kv.second.off, e->host.begin() + kv.second.off + 256, 0.f);
    for (auto& kv : e->slots)
        if (kv.second.accumulate) kv.second.repack(kv.second.raw.data(), e->host.data());
    for (const auto& m : e->mid_src)        // per-column PReLU slopes of the fused Q|K|V projection
        for (int n = 0; n < NQKV; ++n) e->host[m.slope_vec + n] = e->host[m.slopes + (n < NHEAD * QE ? 0 : (n < 2 * NHEAD * QE ? 1 : 2))];
    for (const auto& m : e->mid_src) {      // k-sliced, bank-padded copies for mid_kernel (layout: mid_kernel.cuh)
        float* h = e->host.data();
        std::fill(h + m.dst, h + m.dst + MID_PACK, 0.f);
        for (int k = 0; k < M1_K; ++k)
            for (int n = 0; n < M1_N; ++n) h[m.dst + MID_W1 + mid_widx(M1_KS, M1_N, k, n)] = h[m.wl1 + (int64_t)k * 64 + n];
        for (int k = 0; k < M3_K; ++k)
            for (int n = 0; n < M3_N; ++n) {
                h[m.dst + MID_W3A + mid_widx(M3_KS, M3_N, k, n)] = h[m.wih2 + (int64_t)k * 256 + n];
                h[m.dst + MID_W3B + mid_widx(M3_KS, M3_N, k, n)] = h[m.whh2t + (int64_t)k * 256 + n];
            }
        for (int k = 0; k < M5_K; ++k)
            for (int n = 0; n < M5_N; ++n) h[m.dst + MID_W5 + mid_widx(M5_KS, M5_N, k, n)] = h[m.wl2 + (int64_t)k * 64 + n];
        for (int k = 0; k < M6_K; ++k)
            for (int n = 0; n < M6_N; ++n) h[m.dst + MID_W6 + mid_widx(M6_KS, M6_N, k, n)] = h[m.wqkv + (int64_t)k * NQKV + n];
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int cur = -1;
    CK(cudaGetDevice(&cur));
    const float* old_base = e->dev;              // null before the first commit (pointer fields then hold offsets)
    if (e->dev != nullptr && e->device != cur) {  // the module moved to another GPU: everything device-side is rebuilt there
        release_device_resources(e);
    }
    if (e->dev == nullptr) {
        CK(cudaMalloc(&e->dev, e->total * sizeof(float)));
        e->device = cur;
        shift_pointers(e, e->dev, old_base);
        CK(cudaMalloc(&e->planes, 2 * e->planes_total * sizeof(__nv_bfloat16)));
    }
    CK(cudaMemcpyAsync(e->dev, e->host.data(), e->total * sizeof(float), cudaMemcpyHostToDevice, st));
    for (const auto& ps : e->plane_srcs)        // k-major fp32 [K][N] -> bf16 hi/lo planes [N][ld] (tensor-core B operands)
        CK(umma::split_planes(e->dev + ps.wt_off, 1, ps.N, ps.N, ps.K, ps.ld, e->planes + ps.plane_off + ps.col0,
                              e->planes + e->planes_total + ps.plane_off + ps.col0, st));
    CK(cudaStreamSynchronize(st));
    e->committed = true;
    e->w.gen = (int)(++e->weight_gen & 0x7fffff) + 1;      // never 0 (= a freshly initialised state)
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);     // cached graphs carry the old generation in their kernel arguments
    e->graphs.clear();
    e->graph_kernels.clear();
    return 0;
}

int l2h_sep_state_offsets(void* handle, int64_t* out, int32_t n) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !out) return fail(1, "null argument");
    const int64_t v[L2H_STATE_OFFSETS] = {RING, QK_LD, QK_DIM, V_DIM, ATT, ST_EMB, ST_GATE, ST_CONV, ST_DECONV, ST_ISTFT, ST_BLK,
                                          BK_K, BK_V, BK_H, BK_C, BK_STRIDE};
    if (n < L2H_STATE_OFFSETS) return fail(1, "need room for L2H_STATE_OFFSETS values");
    for (int i = 0; i < L2H_STATE_OFFSETS; ++i) out[i] = v[i];
    return 0;
}

int l2h_sep_state_layout(void* handle, int64_t* header_bytes, int64_t* stride_floats) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return fail(1, "null handle");
    if (header_bytes) *header_bytes = sizeof(StateHeader);
    if (stride_floats) *stride_floats = stream_stride(e->n_blocks);
    return 0;
}

int l2h_sep_state_bytes(void* handle, int32_t batch, size_t* bytes) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !bytes || batch <= 0) return fail(1, "bad argument");
    *bytes = sizeof(StateHeader) + (size_t)batch * stream_stride(e->n_blocks) * sizeof(float);
    return 0;
}

int l2h_sep_state_init(void* handle, void* state, int32_t batch, void* stream) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !state || batch <= 0) return fail(1, "bad argument");
    const int64_t ss = stream_stride(e->n_blocks);
    const int64_t total = sizeof(StateHeader) / 4 + (int64_t)batch * ss;
    state_init_kernel<<<592, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(state), total, ss, batch);
    CK(cudaGetLastError());
    e->launch_count += 1;
    return 0;
}

int l2h_sep_workspace_bytes(void* handle, int32_t batch, int32_t frames, uint32_t flags, size_t* bytes) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !bytes || batch <= 0 || frames <= 0) return fail(1, "bad argument");
    *bytes = (size_t)carve(e->n_blocks, batch, frames, flags).total * sizeof(float);
    return 0;
}

int l2h_sep_tap_info(void* handle, int32_t batch, int32_t frames, int64_t* off, int32_t* n_stages) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return fail(1, "null handle");
    if (off) *off = carve(e->n_blocks, batch, frames, L2H_FLAG_TAPS).TAPS;
    if (n_stages) *n_stages = 1 + 3 * e->n_blocks;
    return 0;
}

int l2h_sep_launches_per_forward(void* handle, int32_t frames, int32_t* n) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !n) return fail(1, "bad argument");
    // one l2h_sep_forward: a one-hop call is 6 kernels per block (gemm, bilstm, mid, qkv, attn, attn_out; 8 with many
    // streams, where the mid section runs as three kernels), a multi-hop call 10.  Streams of one-hop calls go through
    // the pipelined graph instead -- l2h_sep_launch_count has the exact figure for everything this handle launched.
    const int one_hop = e->use_mid ? 6 : 9;
    int dev_ord = 0;
    cudaGetDevice(&dev_ord);
    if (frames == 1 && e->use_mid && e->use_tail && !e->fold_mid_c && dev_ord >= 0 && dev_ord < 64 && g_tail_clusters[dev_ord] > 0)
        *n = 1 + e->n_blocks * 2 + 1;       // front1, (BiLSTM, tail_kernel) per block, back
    else
        *n = 1 + e->n_blocks * (frames == 1 ? one_hop : 10) + 1;
    return 0;
}

int l2h_sep_trace_start(void* handle, int32_t capacity) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || capacity < 0) return fail(1, "bad argument");
    CK(cudaDeviceSynchronize());
    TraceRec* none = nullptr;
    unsigned int zero = 0;
    CK(cudaMemcpyToSymbol(g_trace, &none, sizeof(none)));
    if (e->trace_dev) { cudaFree(e->trace_dev); e->trace_dev = nullptr; e->trace_cap = 0; }
    if (capacity == 0) return 0;                       // tracing off
    CK(cudaMalloc(&e->trace_dev, (size_t)capacity * sizeof(TraceRec)));
    CK(cudaMemset(e->trace_dev, 0, (size_t)capacity * sizeof(TraceRec)));
    e->trace_cap = capacity;
    const unsigned int cap = (unsigned int)capacity;
    CK(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)));
    CK(cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap)));
    CK(cudaMemcpyToSymbol(g_trace, &e->trace_dev, sizeof(e->trace_dev)));
    return 0;
}

int l2h_sep_trace_read(void* handle, void* records_host, int32_t max_records, int32_t* n_records) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !n_records) return fail(1, "bad argument");
    CK(cudaDeviceSynchronize());
    unsigned int n = 0;
    CK(cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)));
    n = std::min<unsigned int>(n, (unsigned int)e->trace_cap);
    *n_records = (int32_t)n;
    const unsigned int take = std::min<unsigned int>(n, (unsigned int)std::max(0, max_records));
    if (records_host && take) CK(cudaMemcpy(records_host, e->trace_dev, (size_t)take * sizeof(TraceRec), cudaMemcpyDeviceToHost));
    unsigned int zero = 0;
    CK(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)));   // the next run starts a fresh trace
    return 0;
}

int l2h_sep_launch_count(void* handle, int64_t* kernels, int32_t reset) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e) return fail(1, "null handle");
    if (kernels) *kernels = e->launch_count;
    if (reset) e->launch_count = 0;
    return 0;
}

int l2h_sep_forward(void* handle, const float* x, int64_t xbs, int64_t xcs, int32_t x_len, const float* emb,
                    void* state, float* y, int64_t ybs, int64_t ycs, int32_t y_len, int32_t batch, int32_t frames,
                    void* ws, size_t ws_bytes, uint32_t flags, void* stream) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (e) { if (int rc_dev = check_device(e)) return rc_dev; }
    if (!e || !x || !emb || !state || !y || !ws) return fail(1, "null argument");
    ChainArgs a{x, xbs, xcs, x_len, emb, static_cast<float*>(state), y, ybs, ycs, y_len, batch, frames,
                static_cast<float*>(ws), ws_bytes, flags & ~L2H_FLAG_GRAPH, 0};
    return run_chain(e, a, static_cast<cudaStream_t>(stream), (flags & L2H_FLAG_GRAPH) != 0);
}

// One-hop calls are pipelined over hops (wavefront graph) only for FEW streams: with many streams every kernel of a hop
// already fills the GPU, the dense stages run on the tensor cores (enqueue_chain) and the hops replay one chain graph.
static bool pipeline_applies(const SepEngine* e, int batch, int cpc, int n_calls) {
    if (!(cpc == 1 && e->use_pipe && e->use_mid && e->n_blocks == 3 && n_calls > 1)) return false;
    return !(e->use_tc && (int64_t)batch * NF > TC_MIN_ROWS);
}

int l2h_sep_stream_host(void* handle, const float* x_host, int32_t x_len, const float* emb, void* state,
                        float* y_host, int32_t y_len, int32_t batch, int32_t n_calls, int32_t cpc,
                        float* x_stage, float* y_stage, void* ws, size_t ws_bytes, void* stream) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (e) { if (int rc_dev = check_device(e)) return rc_dev; }
    if (!e || !x_host || !emb || !state || !y_host || !x_stage || !y_stage || !ws) return fail(1, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // hops moved per host<->device round: one call's worth, or (one-hop calls, pipelining on) a group of up to
    // PIPE_MAX_FRAMES hops that then run as ONE wavefront-pipelined graph of one-hop chains
    const bool pipe = pipeline_applies(e, batch, cpc, n_calls);
    int group = cpc;
    if (pipe) {
        const int64_t slot = pipe_slot_floats(e, batch);
        group = (int)std::min<int64_t>(std::min(pipe_frames_for(e, batch), n_calls), (int64_t)(ws_bytes / sizeof(float)) / slot);
        if (group < 2) group = 1;
    }
    const int hops_total = n_calls * cpc;
    const int in_len = HOP * group + LOOKAHEAD, out_len = HOP * group;      // staging strides
    for (int h0 = 0; h0 < hops_total; h0 += group) {
        const int hops = std::min(group, hops_total - h0);
        const int s0 = h0 * HOP;
        int n_in = std::min(x_len - s0, HOP * hops + LOOKAHEAD);
        if (n_in <= 0) return fail(1, "x_host shorter than n_calls * chunks_per_call * 128 samples");
        CK(cudaMemcpy2DAsync(x_stage, in_len * sizeof(float), x_host + s0, (size_t)x_len * sizeof(float),
                             (size_t)n_in * sizeof(float), (size_t)batch * NMIC, cudaMemcpyHostToDevice, st));
        // (a short last round changes the sizes and therefore the graph key: at most two graphs)
        ChainArgs a{x_stage, (int64_t)NMIC * in_len, in_len, n_in, emb, static_cast<float*>(state), y_stage,
                    (int64_t)NSRC * out_len, out_len, HOP * hops, batch, pipe ? 1 : hops, static_cast<float*>(ws), ws_bytes, 0, 0};
        int rc = (pipe && hops > 1) ? run_pipeline(e, a, hops, st) : run_chain(e, a, st, true);
        if (rc) return rc;
        const int n_out = std::min(y_len - s0, HOP * hops);
        if (n_out > 0)
            CK(cudaMemcpy2DAsync(y_host + s0, (size_t)y_len * sizeof(float), y_stage, out_len * sizeof(float),
                                 (size_t)n_out * sizeof(float), (size_t)batch * NSRC, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return 0;
}

int l2h_sep_stream_dev(void* handle, const float* x_dev, int32_t x_len, const float* emb, void* state,
                       float* y_dev, int32_t y_len, int32_t batch, int32_t n_calls, int32_t cpc, void* ws,
                       size_t ws_bytes, void* stream) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (e) { if (int rc_dev = check_device(e)) return rc_dev; }
    if (!e || !x_dev || !emb || !state || !y_dev || !ws) return fail(1, "null argument");
    if (n_calls <= 0 || cpc <= 0) return fail(1, "n_calls and chunks_per_call must be positive");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    set_clip_base_kernel<<<1, 1, 0, st>>>(static_cast<float*>(state));
    CK(cudaGetLastError());
    e->launch_count += 1;
    ChainArgs a{x_dev, (int64_t)NMIC * x_len, x_len, x_len, emb, static_cast<float*>(state), y_dev,
                (int64_t)NSRC * y_len, y_len, y_len, batch, cpc, static_cast<float*>(ws), ws_bytes, 0, 1};
    if (pipeline_applies(e, batch, cpc, n_calls)) {
        // groups of up to PIPE_MAX_FRAMES one-frame calls, each group one wavefront-pipelined graph
        const int64_t slot = pipe_slot_floats(e, batch);
        int kmax = (int)std::min<int64_t>(pipe_frames_for(e, batch), (int64_t)(ws_bytes / sizeof(float)) / slot);
        if (kmax >= 2) {
            int done = 0;
            while (done < n_calls) {
                const int K = std::min(kmax, n_calls - done);
                if (K == 1) { if (int rc = run_chain(e, a, st, true)) return rc; }
                else if (int rc = run_pipeline(e, a, K, st)) return rc;
                done += K;
            }
            return 0;
        }
    }
    for (int i = 0; i < n_calls; ++i)
        if (int rc = run_chain(e, a, st, true)) return rc;
    return 0;
}

int l2h_sep_stream_workspace_bytes(void* handle, int32_t batch, int32_t chunks_per_call, size_t* bytes) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !bytes || batch <= 0 || chunks_per_call <= 0) return fail(1, "bad argument");
    if (pipeline_applies(e, batch, chunks_per_call, 2))
        *bytes = (size_t)pipe_slot_floats(e, batch) * pipe_frames_for(e, batch) * sizeof(float);
    else
        *bytes = (size_t)carve(e->n_blocks, batch, chunks_per_call, 0).total * sizeof(float);
    return 0;
}

int l2h_sep_set_option(void* handle, const char* name, int32_t value) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !name) return fail(1, "bad argument");
    const std::string n(name);
    if (n == "defaults") {                      // lane counts and mid split back to the built-in defaults
        const SepEngine d{};
        e->pipe_alanes = d.pipe_alanes; e->pipe_qlanes = d.pipe_qlanes; e->pipe_clanes = d.pipe_clanes; e->pipe_tlanes = d.pipe_tlanes;
        e->pipe_olanes = d.pipe_olanes; e->pipe_flanes = d.pipe_flanes; e->pipe_blanes = d.pipe_blanes;
        e->pipe_split_mid = d.pipe_split_mid; e->pipe_frames = d.pipe_frames; e->pipe_skip = 0; e->pipe_pdl = d.pipe_pdl; e->pipe_midb_hops = d.pipe_midb_hops;
    }
    else if (n == "pipeline") e->use_pipe = value != 0;
    else if (n == "pipeline_frames") e->pipe_frames = value <= 0 ? 0 : std::max(2, std::min(PIPE_MAX_FRAMES, (int)value));
    else if (n == "pipeline_lanes") e->pipe_alanes = std::max(1, std::min(PIPE_LANES, (int)value));
    else if (n == "pipeline_debug_skip") e->pipe_skip = value;
    else if (n == "pipeline_pdl") e->pipe_pdl = value;
    else if (n == "pipeline_gemm_shape") e->pipe_gemm_shape = std::max(0, std::min(2, (int)value));
    else if (n == "pipeline_midb_hops") e->pipe_midb_hops = std::max(1, std::min(PIPE_MIDB_MAX, (int)value));
    else if (n == "pipeline_split_mid") e->pipe_split_mid = value != 0;
    else if (n == "pipeline_qkv_lanes") e->pipe_qlanes = std::max(1, std::min(PIPE_QLANES, (int)value));
    else if (n == "pipeline_midc_lanes") e->pipe_clanes = std::max(1, std::min(PIPE_CLANES, (int)value));
    else if (n == "pipeline_attn_lanes") e->pipe_tlanes = std::max(1, std::min(PIPE_TLANES, (int)value));
    else if (n == "pipeline_out_lanes") e->pipe_olanes = std::max(1, std::min(PIPE_OLANES, (int)value));
    else if (n == "pipeline_front_lanes") e->pipe_flanes = std::max(1, std::min(PIPE_FLANES, (int)value));
    else if (n == "pipeline_back_lanes") e->pipe_blanes = std::max(1, std::min(PIPE_BLANES, (int)value));
    else if (n == "pdl") e->use_pdl = value != 0;
    else if (n == "fused_mid") e->use_mid = value != 0;
    else if (n == "fused_tail") e->use_tail = value != 0;
    else if (n == "back_many") e->use_back_many = value != 0;
    else if (n == "tc_pdl") e->tc_pdl = (int)value;
    else if (n == "tc_lstm_min") e->tcl_min_seqdirs = std::max(1, (int)value);
    else if (n == "mid_split_large") e->mid_split_large = value != 0;
    else if (n == "fold_mid_c") e->fold_mid_c = value != 0;
    else if (n == "tensor_cores") e->use_tc = value != 0;
    else if (n == "fuse_ih") e->fuse_ih = value != 0;
    else if (n == "bf16") e->tc_passes = value == 0 ? 3 : (value == 2 ? 1 : 2);   // 1: bf16 weights x split activations; 2: plain bf16 both
    else if (n == "graph_stats") e->graph_stats = value != 0;
    else return fail(2, "unknown option: " + n);
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);     // cached graphs were built with the old setting
    e->graphs.clear();
    e->graph_kernels.clear();
    return 0;
}

int l2h_sep_pipeline_frames(void* handle, int32_t* frames) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (!e || !frames) return fail(1, "bad argument");
    *frames = (e->use_pipe && e->use_mid && e->n_blocks == 3) ? pipe_frames_for(e, 1) : 1;
    return 0;
}

int l2h_sep_profile(void* handle, const float* x_dev, int32_t x_len, const float* emb, void* state, float* y_dev,
                    int32_t batch, int32_t frames, void* ws, size_t ws_bytes, int32_t iters, const char** names,
                    float* ms_total, int32_t* counts, int32_t* n_names, void* stream) {
    SepEngine* e = static_cast<SepEngine*>(handle);
    if (e) { if (int rc_dev = check_device(e)) return rc_dev; }
    if (!e || !x_dev || !emb || !state || !y_dev || !ws || !names || !ms_total || !counts || !n_names)
        return fail(1, "null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Profiler prof;
    std::vector<std::string> order;
    std::map<std::string, std::pair<double, int>> acc;
    const int out_len = HOP * frames;
    for (int it = -2; it < iters; ++it) {           // two untimed warm-up chains
        prof.used = 0;
        prof.names.clear();
        ChainArgs a{x_dev, (int64_t)NMIC * x_len, x_len, x_len, emb, static_cast<float*>(state), y_dev,
                    (int64_t)NSRC * out_len, out_len, out_len, batch, frames, static_cast<float*>(ws), ws_bytes, 0, 0};
        a.prof = &prof;
        if (int rc = enqueue_chain(e, a, st)) return rc;
        CK(cudaStreamSynchronize(st));
        if (it < 0) continue;
        for (int i = 1; i < prof.used; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, prof.ev[i - 1], prof.ev[i]));
            const std::string nm = prof.names[i];
            if (!acc.count(nm)) order.push_back(nm);
            acc[nm].first += ms;
            acc[nm].second += 1;
        }
    }
    for (auto ev : prof.ev) cudaEventDestroy(ev);
    static std::vector<std::string> keep;            // storage for the returned C strings
    keep = order;
    int n = 0;
    for (auto& nm : keep) {
        if (n >= 64) break;
        names[n] = nm.c_str();
        ms_total[n] = (float)acc[nm].first;
        counts[n] = acc[nm].second;
        ++n;
    }
    *n_names = n;
    return 0;
}

}  // extern "C"
