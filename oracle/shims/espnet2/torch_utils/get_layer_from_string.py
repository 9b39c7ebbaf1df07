"""ORACLE SHIM: espnet2 get_layer = case-insensitive lookup of a torch.nn class."""
import torch


def get_layer(l_name, library=torch.nn):
    names = [x for x in dir(library) if x.lower() == l_name.lower()]
    if len(names) != 1:
        raise NotImplementedError(l_name)
    return getattr(library, names[0])
