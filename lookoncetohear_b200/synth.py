"""Seeded synthetic inputs of the shapes the reference's evaluation feeds the two networks
(there is no dataset in the image).  CPU generators so that the same
numbers are produced on the build box and the GPU box."""
import torch


def mixture(batch, n_samples, seed0=1000):
    """target + interferer, each 0.07*N(0,1), [B,2,N] fp32 -> (mixture, target)."""
    mix, tgt = [], []
    for b in range(batch):
        g = torch.Generator().manual_seed(seed0 + b)
        t = 0.07 * torch.randn(2, n_samples, generator=g)
        i = 0.07 * torch.randn(2, n_samples, generator=g)
        mix.append(t + i)
        tgt.append(t)
    return torch.stack(mix), torch.stack(tgt)


def embedding(batch, dim=256, seed0=3000):
    """|N(0,1)| L2-normalised, [B,1,dim] (d-vectors are ReLU'd and unit norm)."""
    out = []
    for b in range(batch):
        g = torch.Generator().manual_seed(seed0 + b)
        e = torch.randn(dim, generator=g).abs()
        out.append(e / e.norm())
    return torch.stack(out).unsqueeze(1)


def enrollment(batch, n_samples, seed0=2000):
    """0.1*N(0,1) binaural 'look' recordings [B,2,N]."""
    out = []
    for b in range(batch):
        g = torch.Generator().manual_seed(seed0 + b)
        out.append(0.1 * torch.randn(2, n_samples, generator=g))
    return torch.stack(out)


def seeded_state_dict(module_ctor, seed=0):
    """Weights = PyTorch default init of the parameter containers under a seed (identical to the
    reference modules constructed under the same seed; tests/test_oracle.py checks that)."""
    torch.manual_seed(seed)
    return module_ctor()
