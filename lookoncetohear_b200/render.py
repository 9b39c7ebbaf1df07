"""GPU binaural renderer: mono events + per-ear impulse responses -> binaural events, mixture and target, as the
reference's simulators and dataset do on the CPU (src/datasets/multi_ch_simulator.py:40-61 `SOFASimulator._convolve`;
src/datasets/MixLibriSpeechNoisyEnrollNorm.py:179-202 noise scaling / peak normalisation / mixture).  The arithmetic is
`l2h_render_binaural` (hand-written CUDA); no CPU fallback."""
import torch

from . import _cabi


def render_binaural(srcs, rirs, noise=None, noise_scale=None):
    """srcs [B, S, N] mono events, rirs [B, S, 2, L] impulse responses (at the sampling rate of `srcs`), noise [B, 2, N]
    binaural background or None, noise_scale [B] or None.  CUDA tensors.
    Returns (events [B, S, 2, N], mixture [B, 2, N], norm [B]); the target of a sample is events[:, tgt_idx]."""
    if not srcs.is_cuda:
        raise RuntimeError("lookoncetohear_b200.render.render_binaural needs CUDA tensors (no CPU fallback)")
    dev = srcs.device
    src = srcs.contiguous().float()
    rir = rirs.to(dev, torch.float32).contiguous()
    B, S, N = src.shape
    if rir.shape[:3] != (B, S, 2):
        raise ValueError(f"rirs must be [B, S, 2, L], got {tuple(rir.shape)}")
    L = rir.shape[-1]
    nz = noise.to(dev, torch.float32).contiguous() if noise is not None else None
    ns = noise_scale.to(dev, torch.float32).contiguous() if noise_scale is not None else None
    events = torch.empty(B, S, 2, N, dtype=torch.float32, device=dev)
    mixture = torch.empty(B, 2, N, dtype=torch.float32, device=dev)
    norm = torch.empty(B, dtype=torch.float32, device=dev)
    scratch = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().l2h_render_binaural(
            src.data_ptr(), rir.data_ptr(), nz.data_ptr() if nz is not None else None,
            ns.data_ptr() if ns is not None else None, B, S, N, L, events.data_ptr(), mixture.data_ptr(), norm.data_ptr(),
            scratch.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return events, mixture, norm
