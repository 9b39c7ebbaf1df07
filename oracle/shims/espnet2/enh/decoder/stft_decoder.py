"""ORACLE SHIM: importable placeholder; the enrollment forward constructs but never calls it."""
import torch


class STFTDecoder(torch.nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
