"""ORACLE SHIM."""
import torch


def make_pad_mask(lengths, xs=None, length_dim=-1):
    lengths = torch.as_tensor(lengths)
    n = xs.size(length_dim)
    idx = torch.arange(n, device=xs.device)
    mask = idx[None, :] >= lengths[:, None].to(xs.device)
    shape = [1] * xs.dim()
    shape[0] = mask.shape[0]
    shape[length_dim] = n
    return mask.view(shape).expand_as(xs)
