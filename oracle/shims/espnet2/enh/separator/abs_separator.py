"""ORACLE SHIM: espnet2 AbsSeparator = nn.Module + ABC."""
from abc import ABC, abstractmethod
import torch


class AbsSeparator(torch.nn.Module, ABC):
    @abstractmethod
    def forward(self, *args, **kwargs):
        raise NotImplementedError

    @property
    @abstractmethod
    def num_spk(self):
        raise NotImplementedError
