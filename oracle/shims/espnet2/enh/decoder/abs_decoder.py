"""ORACLE SHIM."""
import torch


class AbsDecoder(torch.nn.Module):
    pass
