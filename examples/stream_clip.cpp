// examples/stream_clip.cpp -- a C++ host driving the separator through the C ABI alone (no Python, no torch):
// create the engine, feed it the weights tensor by tensor (here: deterministic pseudo-random values, the way a
// checkpoint converter would walk l2h_sep_weight_info), then stream a 4 s binaural clip from pinned host memory in
// 8 ms hops with l2h_sep_stream_host and report the rate.  The reference does the same with
// `for chunk in clip: model.predict(chunk, embed, state, pad=False)` (net.py:54-66).
//
//   g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include examples/stream_clip.cpp \
//       -L lookoncetohear_b200/lib -llookonce_b200 -L /usr/local/cuda/lib64 -lcudart \
//       -Wl,-rpath,$PWD/lookoncetohear_b200/lib -o stream_clip
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "lookonce_b200.h"

#define L2H(call)                                                                     \
    do {                                                                              \
        if ((call) != 0) { std::fprintf(stderr, "%s: %s\n", #call, l2h_last_error()); return 1; } \
    } while (0)
#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } \
    } while (0)

int main() {
    const int hops = 500, hop = 128, lookahead = 64, n = hops * hop;     // 4 s at 16 kHz
    l2h_sep_config cfg = {128, 64, 256, 2, 64, 4, 1, 1, 3, 64, 50, 1, 1, 1, 2};   // configs/tsh.json model_params
    void* h = nullptr;
    L2H(l2h_sep_create(&cfg, &h));

    // weights: walk the expected tensors by index (a converter would look each name up in the checkpoint)
    int32_t n_expected = 0;
    L2H(l2h_sep_weights_expected(h, &n_expected, nullptr));
    uint32_t seed = 12345u;
    for (int32_t i = 0; i < n_expected; ++i) {
        const char* name = nullptr;
        int64_t numel = 0;
        L2H(l2h_sep_weight_info(h, i, &name, &numel));
        std::vector<float> w((size_t)numel);
        for (auto& v : w) { seed = seed * 1664525u + 1013904223u; v = ((seed >> 8) * (1.f / 16777216.f) - 0.5f) * 0.2f; }
        L2H(l2h_sep_load_weight(h, name, w.data(), numel));
    }
    L2H(l2h_sep_commit_weights(h, nullptr));

    // buffers: pinned host clip in/out, device state, staging and workspace as the header asks for
    float *x_host = nullptr, *y_host = nullptr, *emb = nullptr, *x_stage = nullptr, *y_stage = nullptr;
    void *state = nullptr, *work = nullptr;
    CU(cudaMallocHost(&x_host, sizeof(float) * 2 * n));
    CU(cudaMallocHost(&y_host, sizeof(float) * 2 * n));
    for (int i = 0; i < 2 * n; ++i) x_host[i] = 0.1f * std::sin(0.01f * (float)i);
    std::vector<float> emb_h(256);
    for (int i = 0; i < 256; ++i) emb_h[i] = std::cos(0.37f * (float)i) / 16.f;
    CU(cudaMalloc(&emb, sizeof(float) * 256));
    CU(cudaMemcpy(emb, emb_h.data(), sizeof(float) * 256, cudaMemcpyHostToDevice));
    size_t state_bytes = 0, work_bytes = 0;
    int32_t group = 1;
    L2H(l2h_sep_state_bytes(h, 1, &state_bytes));
    L2H(l2h_sep_stream_workspace_bytes(h, 1, 1, &work_bytes));
    L2H(l2h_sep_pipeline_frames(h, &group));          // one-hop calls per pipelined graph
    CU(cudaMalloc(&state, state_bytes));
    CU(cudaMalloc(&work, work_bytes));
    CU(cudaMalloc(&x_stage, sizeof(float) * 2 * (hop * group + lookahead)));
    CU(cudaMalloc(&y_stage, sizeof(float) * 2 * hop * group));

    double best_ms = 1e30;
    for (int it = 0; it < 5; ++it) {                  // first pass captures the graph
        L2H(l2h_sep_state_init(h, state, 1, nullptr));
        CU(cudaDeviceSynchronize());
        const auto t0 = std::chrono::steady_clock::now();
        L2H(l2h_sep_stream_host(h, x_host, n, emb, state, y_host, n, 1, hops, 1, x_stage, y_stage, work, work_bytes, nullptr));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (it > 0 && ms < best_ms) best_ms = ms;
    }
    int64_t kernels = 0;
    L2H(l2h_sep_launch_count(h, &kernels, 0));
    double energy = 0.0;
    for (int i = 0; i < 2 * n; ++i) energy += (double)y_host[i] * y_host[i];
    std::printf("%d hops in %.3f ms = %.0f frames/s (%.0fx real time), %lld kernels launched in total, output energy %.6g\n", hops,
                best_ms, hops / (best_ms * 1e-3), hops / (best_ms * 1e-3) / 125.0, (long long)kernels, energy);

    cudaFree(y_stage); cudaFree(x_stage); cudaFree(work); cudaFree(state); cudaFree(emb);
    cudaFreeHost(y_host); cudaFreeHost(x_host);
    L2H(l2h_sep_destroy(h));
    return 0;
}
