// On-GPU evaluation epilogue: the three per-mixture figures the reference's evaluation
// driver computes on the CPU after `outputs.cpu()` -- /root/reference/src/ts_hear_test.py:139-146 --
//   output_sisnr  = mean over ears of SI-SNR(estimate, target)
//   si_snr_i      = mean over ears of SI-SNR(estimate, target) - SI-SNR(mixture, target)
//   embedding_sim = cosine_similarity(embedding, embedding_gt)
// so that the device->host traffic of an evaluation step shrinks from the separated audio to three floats per mixture.
// SI-SNR as torchmetrics' scale_invariant_signal_noise_ratio (zero-mean; alpha = (<p,t>+eps)/(<t,t>+eps);
// 10 log10((|alpha t|^2+eps)/(|alpha t - p|^2+eps)), eps = float32 eps).  All sums in double.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/lookonce_b200.h"

namespace l2h {
int fail(int code, const std::string& msg);

__device__ __forceinline__ double blk_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

__device__ __forceinline__ double si_snr_from_sums(double n, double sp, double st, double spp, double stt, double spt) {
    const double eps = 1.1920928955078125e-07;          // torch.finfo(torch.float32).eps
    const double pt = spt - sp * st / n, tt = stt - st * st / n, pp = spp - sp * sp / n;
    const double alpha = (pt + eps) / (tt + eps);
    const double sig = alpha * alpha * tt;
    const double noise = sig - 2.0 * alpha * pt + pp;
    return 10.0 * log10((sig + eps) / (fmax(noise, 0.0) + eps));
}

// one CTA per mixture; 256 threads
__global__ void __launch_bounds__(256)
eval_metrics_kernel(const float* __restrict__ est, const float* __restrict__ tgt, const float* __restrict__ mix, int ch, int n,
                    const float* __restrict__ emb, const float* __restrict__ emb_gt, int dim, float* __restrict__ out) {
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    double acc_s = 0.0, acc_i = 0.0;
    for (int c = 0; c < ch; ++c) {
        const int64_t off = ((int64_t)b * ch + c) * n;
        double sp = 0, st = 0, sm = 0, spp = 0, stt = 0, smm = 0, spt = 0, smt = 0;
        for (int i = tid; i < n; i += 256) {
            const double p = est[off + i], t = tgt[off + i];
            sp += p; st += t; spp += p * p; stt += t * t; spt += p * t;
            if (mix) { const double m = mix[off + i]; sm += m; smm += m * m; smt += m * t; }
        }
        sp = blk_sum(sp, red); st = blk_sum(st, red); spp = blk_sum(spp, red); stt = blk_sum(stt, red); spt = blk_sum(spt, red);
        const double s_est = si_snr_from_sums((double)n, sp, st, spp, stt, spt);
        acc_s += s_est;
        if (mix) {
            sm = blk_sum(sm, red); smm = blk_sum(smm, red); smt = blk_sum(smt, red);
            acc_i += s_est - si_snr_from_sums((double)n, sm, st, smm, stt, smt);
        }
    }
    double cs = 0.0;
    if (emb && emb_gt) {
        double xy = 0, xx = 0, yy = 0;
        for (int i = tid; i < dim; i += 256) {
            const double x = emb[(int64_t)b * dim + i], y = emb_gt[(int64_t)b * dim + i];
            xy += x * y; xx += x * x; yy += y * y;
        }
        xy = blk_sum(xy, red); xx = blk_sum(xx, red); yy = blk_sum(yy, red);
        cs = xy / (fmax(sqrt(xx), 1e-8) * fmax(sqrt(yy), 1e-8));       // F.cosine_similarity, eps = 1e-8
    }
    if (tid == 0) {
        out[b * 3 + 0] = (float)(acc_s / ch);
        out[b * 3 + 1] = (float)(acc_i / ch);
        out[b * 3 + 2] = (float)cs;
    }
}
}  // namespace l2h

extern "C" int l2h_eval_metrics(const float* est_dev, const float* target_dev, const float* mixture_dev, int32_t batch, int32_t channels,
                                int32_t n_samples, const float* emb_dev, const float* emb_gt_dev, int32_t emb_dim, float* out_dev,
                                void* stream) {
    using namespace l2h;
    if (!est_dev || !target_dev || !out_dev || batch <= 0 || channels <= 0 || n_samples <= 1)
        return fail(1, "l2h_eval_metrics: bad argument");
    eval_metrics_kernel<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(est_dev, target_dev, mixture_dev, channels, n_samples,
                                                                                emb_dev, emb_gt_dev, emb_dim, out_dev);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(3, std::string("eval_metrics_kernel: ") + cudaGetErrorString(e));
    return 0;
}
