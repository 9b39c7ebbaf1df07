"""Enrollment network parity: CUDA path (through the reference-shaped API -> C ABI) vs the CPU
oracle and the fixture generated from the reference.  Gates: rel-L2 <= 1e-3, cosine >= 0.9999."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lookoncetohear_b200 import EmbedTFGridNet, synth
from oracle import restate as rs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def model(embed_params, dev):
    torch.manual_seed(0)
    net = EmbedTFGridNet(**embed_params).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return net.to(dev), sd


def _check(o, r):
    o = o.float().cpu()
    assert o.shape == r.shape and torch.isfinite(o).all()
    assert rs.rel_l2(o, r) <= 1e-3, rs.rel_l2(o, r)
    assert float(F.cosine_similarity(o, r).min()) >= 0.9999


@pytest.mark.parametrize("B,N", [(2, 4800), (1, 1000), (3, 64 * 9 + 5)])
def test_embed_vs_oracle(model, dev, B, N):
    net, sd = model
    x = synth.enrollment(B, N, seed0=2000 + N)
    with torch.no_grad():
        o = net(x.to(dev))
    _check(o, rs.embed_forward(sd, x))


def test_embed_golden(model, dev):
    g = np.load(os.path.join(GOLD, "embed_golden.npz"))
    net, _ = model
    assert int(g["seed"]) == 0
    with torch.no_grad():
        o = net(synth.enrollment(2, int(g["n"])).to(dev))
    _check(o, torch.from_numpy(g["emb"]))


def test_embed_scale_invariance(model, dev):
    """The net divides by the utterance std (tfgridnet.py:109-110): scaling the input is a no-op.
    A size-independent property, checked at the full 5 s length."""
    net, _ = model
    x = synth.enrollment(1, 80000)
    with torch.no_grad():
        a = net(x.to(dev))
        b = net((7.5 * x).to(dev))
    assert rs.rel_l2(a.cpu(), b.cpu()) < 1e-4


def test_embed_batch_independence_and_chunking(model, dev):
    net, sd = model
    x = synth.enrollment(3, 3200)
    with torch.no_grad():
        full = net(x.to(dev)).cpu()
        one = torch.cat([net(x[i:i + 1].to(dev)).cpu() for i in range(3)])
    assert rs.rel_l2(full, one) < 1e-5


def test_embedding_feeds_separation(model, dev, tsh_params):
    """The evaluation path of ts_hear_test.py:134-138: enrollment -> embedding -> separation."""
    from lookoncetohear_b200 import Net
    net, sd = model
    torch.manual_seed(1)
    sep = Net(**tsh_params).eval()
    sd_sep = {k: v.detach().clone() for k, v in sep.state_dict().items()}
    sep = sep.to(dev)
    enr = synth.enrollment(2, 3200)
    mix, _ = synth.mixture(2, 128 * 8)
    with torch.no_grad():
        emb = net(enr.to(dev)).unsqueeze(1)
        y = sep(mix.to(dev), emb)
    emb_ref = rs.embed_forward(sd, enr).unsqueeze(1)
    y_ref = rs.sep_forward(sd_sep, mix, emb_ref)
    assert rs.rel_l2(y.cpu(), y_ref) <= 1e-3
