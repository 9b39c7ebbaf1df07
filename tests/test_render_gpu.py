"""GPU binaural renderer (l2h_render_binaural) against the reference's data-side arithmetic restated with the same
library calls: scipy.signal.convolve(src, rir[ear])[:len(src)] per ear (src/datasets/multi_ch_simulator.py:56-58) and
the noise scaling / peak normalisation / mixture of src/datasets/MixLibriSpeechNoisyEnrollNorm.py:179-202."""
import numpy as np
import pytest
import torch
from scipy.signal import convolve

from lookoncetohear_b200.render import render_binaural

pytestmark = pytest.mark.gpu


def _reference(srcs, rirs, noise, noise_scale):
    B, S, N = srcs.shape
    ev = np.zeros((B, S, 2, N), np.float64)
    for b in range(B):
        for s in range(S):
            for ear in range(2):
                ev[b, s, ear] = convolve(srcs[b, s].astype(np.float64), rirs[b, s, ear].astype(np.float64))[:N]
    nz = noise.astype(np.float64) * noise_scale[:, None, None]
    norm = np.abs(ev.sum(1) + nz).max(axis=(1, 2))
    norm = np.where(norm > 1.0, norm, 1.0)
    ev = ev / norm[:, None, None, None]
    nz = nz / norm[:, None, None]
    return ev, ev.sum(1) + nz, norm


@pytest.mark.parametrize("B,S,N,L,gain", [(2, 3, 16000, 73, 1.0), (1, 2, 5001, 1, 0.2), (2, 1, 4096, 1500, 6.0), (3, 4, 80000, 200, 3.0)])
def test_render_matches_reference_arithmetic(B, S, N, L, gain):
    rng = np.random.default_rng(7 + L)
    srcs = (gain * 0.2 * rng.standard_normal((B, S, N))).astype(np.float32)
    rirs = (rng.standard_normal((B, S, 2, L)) * np.exp(-np.arange(L) / max(L / 6, 1.0))).astype(np.float32)
    noise = (0.05 * rng.standard_normal((B, 2, N))).astype(np.float32)
    nscale = rng.uniform(0.5, 2.0, B).astype(np.float32)
    ev, mix, norm = render_binaural(torch.from_numpy(srcs).cuda(), torch.from_numpy(rirs).cuda(), torch.from_numpy(noise).cuda(),
                                    torch.from_numpy(nscale).cuda())
    ev_r, mix_r, norm_r = _reference(srcs, rirs, noise, nscale)
    rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    assert rel(ev.cpu().numpy().astype(np.float64), ev_r) < 1e-5
    assert rel(mix.cpu().numpy().astype(np.float64), mix_r) < 1e-5
    assert np.allclose(norm.cpu().numpy(), norm_r, rtol=1e-5)
    assert float(mix.abs().max()) <= 1.0 + 1e-5
    if gain >= 3.0:
        assert (norm_r > 1.0).any()                       # the normalisation branch was exercised


def test_render_without_noise():
    rng = np.random.default_rng(3)
    srcs = (0.1 * rng.standard_normal((1, 2, 3000))).astype(np.float32)
    rirs = rng.standard_normal((1, 2, 2, 50)).astype(np.float32) * 0.1
    ev, mix, norm = render_binaural(torch.from_numpy(srcs).cuda(), torch.from_numpy(rirs).cuda())
    ev_r, mix_r, _ = _reference(srcs, rirs, np.zeros((1, 2, 3000), np.float32), np.ones(1, np.float32))
    assert np.allclose(mix.cpu().numpy(), mix_r, atol=1e-6)
    assert float(norm[0]) == 1.0
