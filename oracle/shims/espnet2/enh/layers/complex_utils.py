"""ORACLE SHIM: the few helpers the reference imports (unused on the hot path)."""
import torch


def is_torch_complex_tensor(c):
    return torch.is_tensor(c) and torch.is_complex(c)


def is_complex(c):
    return is_torch_complex_tensor(c)


def new_complex_like(ref, real_imag):
    return torch.complex(*real_imag)
