"""Pin the oracle (oracle/restate.py): against the reference's own code where the checkout exists
(build container), and against the committed fixtures generated from the reference (everywhere).
The reference repo has no golden vectors of its own."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lookoncetohear_b200 import Net, synth
from oracle import ref_loader as rl
from oracle import restate as rs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not rl.available(), reason="reference checkout not present on this box")


def _wsum(sd):
    return np.array([float(sum(v.double().abs().sum() for v in sd.values())),
                     float(sum((v.double() ** 2).sum() for v in sd.values()))])


def _seeded_sd(tsh_params, seed):
    torch.manual_seed(seed)
    return {k: v.detach().clone() for k, v in Net(**tsh_params).state_dict().items()}


@needs_ref
def test_seeded_init_matches_reference(tsh_params):
    ref = rl.reference_net(0).state_dict()
    mine = _seeded_sd(tsh_params, 0)
    assert set(ref) == set(mine)
    for k in ref:
        assert torch.allclose(ref[k], mine[k], atol=1e-7, rtol=0), k


@needs_ref
def test_param_counts():
    assert sum(p.numel() for p in rl.reference_net(0).parameters()) == 2_037_960
    assert sum(p.numel() for p in rl.reference_embed_net(0).parameters()) == 2_368_681


@needs_ref
def test_restatement_equals_reference_forward_and_state():
    net = rl.reference_net(3)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x, _ = synth.mixture(2, 128 * 9 + 77, seed0=50)
    e = synth.embedding(2, seed0=60)
    with torch.no_grad():
        y_ref = net(x, e)
        st_ref = net.init_buffers(2, "cpu")
        _, st_ref = net.predict(x, e[:, 0], st_ref)
    st = rs.sep_init_state(sd, 2)
    y, st = rs.sep_predict(sd, x, e[:, 0], st)
    assert rs.rel_l2(y, y_ref) < 5e-6
    for k in ("conv_buf", "deconv_buf", "istft_buf"):
        assert rs.rel_l2(st[k], st_ref[k]) < 5e-6, k
    for i in range(3):
        for k in ("K_buf", "V_buf", "h0", "c0"):
            assert rs.rel_l2(st["gridnet_bufs"][f"buf{i}"][k], st_ref["gridnet_bufs"][f"buf{i}"][k]) < 5e-6


@needs_ref
def test_restatement_fp64_floor():
    net = rl.reference_net(1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x, _ = synth.mixture(1, 128 * 8)
    e = synth.embedding(1)
    with torch.no_grad():
        y_ref = net(x, e)
    y64 = rs.sep_forward(rs.cast_sd(sd, torch.float64), x.double(), e.double())
    assert rs.rel_l2(y64, y_ref) < 5e-6


@needs_ref
def test_embed_restatement_equals_reference():
    en = rl.reference_embed_net(2)
    sd = {k: v.clone() for k, v in en.state_dict().items()}
    x = synth.enrollment(2, 5000)
    with torch.no_grad():
        r = en(x)
    o = rs.embed_forward(sd, x)
    assert rs.rel_l2(o, r) < 5e-5
    assert float(F.cosine_similarity(o, r).min()) > 0.99999


def test_restatement_streaming_equals_whole(tsh_params):
    sd = _seeded_sd(tsh_params, 5)
    x, _ = synth.mixture(1, 128 * 7)
    e = synth.embedding(1)
    y = rs.sep_forward(sd, x, e)
    st = rs.sep_init_state(sd, 1)
    xp = F.pad(x, (0, 64))
    ys = torch.cat([rs.sep_predict(sd, xp[..., 128 * i:128 * i + 192], e[:, 0], st, pad=False)[0]
                    for i in range(7)], -1)
    assert rs.rel_l2(ys, y) < 5e-6


def test_fast_timing_mode_equals_explicit_loop(tsh_params):
    sd = _seeded_sd(tsh_params, 6)
    x, _ = synth.mixture(1, 128 * 5)
    e = synth.embedding(1)
    y = rs.sep_forward(sd, x, e)
    rs.set_fast(True)
    try:
        yf = rs.sep_forward(sd, x, e)
    finally:
        rs.set_fast(False)
    assert rs.rel_l2(yf, y) < 5e-6


def test_golden_sep(tsh_params):
    g = np.load(os.path.join(GOLD, "sep_golden.npz"))
    sd = _seeded_sd(tsh_params, int(g["seed"]))
    assert np.allclose(_wsum(sd), g["wsum"], rtol=1e-9), "seeded init differs from the build that made the fixture"
    B, N = int(g["B"]), int(g["N"])
    x, _ = synth.mixture(B, N)
    e = synth.embedding(B)
    st = rs.sep_init_state(sd, B)
    y, st = rs.sep_predict(sd, x, e[:, 0], st)
    assert rs.rel_l2(y, torch.from_numpy(g["y"])) < 5e-6
    assert rs.rel_l2(st["gridnet_bufs"]["buf2"]["h0"], torch.from_numpy(g["h0_buf2"])) < 5e-6


def test_golden_embed(embed_params):
    g = np.load(os.path.join(GOLD, "embed_golden.npz"))
    from lookoncetohear_b200.embed import EmbedTFGridNet
    torch.manual_seed(int(g["seed"]))
    sd = {k: v.detach().clone() for k, v in EmbedTFGridNet(**embed_params).state_dict().items()}
    assert np.allclose(_wsum(sd), g["wsum"], rtol=1e-9)
    o = rs.embed_forward(sd, synth.enrollment(2, int(g["n"])))
    assert rs.rel_l2(o, torch.from_numpy(g["emb"])) < 5e-5


def test_si_sdr_known_answer():
    t = torch.sin(torch.arange(1000.) * 0.1)[None]
    n = torch.cos(torch.arange(1000.) * 0.37)[None]
    n = n - (n * t).sum() / (t * t).sum() * t          # orthogonal noise
    p = 3.0 * t + 0.3 * n * (t.norm() / n.norm()) * 3.0
    assert abs(float(rs.si_sdr(p, t)) - 20 * np.log10(1 / 0.3)) < 1e-3


@needs_ref
def test_stft_shim_equals_the_stft_the_reference_vendors():
    """The one piece of espnet2 arithmetic the reference DOES carry -- Stft.forward, src/models/tfgridnet_orig/
    stft.py:68-195, identical to what espnet2's STFTEncoder calls -- pins the shim the enrollment oracle uses
    (oracle/shims/espnet2/enh/encoder/stft_encoder.py): same frames, same bins, same values, same output lengths."""
    import importlib
    rl._prepare()
    vendored = importlib.import_module("src.models.tfgridnet_orig.stft").Stft
    from espnet2.enh.encoder.stft_encoder import STFTEncoder
    for n_fft, hop, n in ((128, 64, 5000), (128, 64, 4999), (192, 128, 3001)):
        x = synth.enrollment(3, n).transpose(1, 2).contiguous()          # [B, N, M] as EmbedTFGridNet.forward passes it
        ilens = torch.tensor([n, n, n])
        ref, rl_out = vendored(n_fft=n_fft, win_length=n_fft, hop_length=hop, window="hann")(x, ilens)     # [B,T,M,F,2]
        got, gl_out = STFTEncoder(n_fft, n_fft, hop, window="hann")(x, ilens)                             # complex [B,T,M,F]
        assert got.shape == ref.shape[:-1] and got.shape[1] == 1 + n // hop
        assert torch.equal(rl_out, gl_out)
        assert rs.rel_l2(torch.view_as_real(got), ref) < 1e-6
