"""ORACLE (test infrastructure): import the reference's own modules, unmodified, by path.

Works only where the reference checkout exists (/root/reference in the build container, or a
driver-placed copy under baseline/_ref); the GPU box has neither, there the checker is
oracle/restate.py + the committed fixtures under tests/golden/.

The reference's two third-party packages that are absent from this image (asteroid_filterbanks,
espnet2 -- SURVEY.md section 8c) are satisfied by the restatements in oracle/shims/.
Nothing is copied from the reference: its files are executed where they lie.
"""
import importlib
import json
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")
_CANDIDATES = ["/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def reference_root():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "src", "models", "tfgridnet_realtime", "net.py")):
            return c
    return None


def available():
    return reference_root() is not None


def _prepare():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference checkout not present (expected on the GPU box)")
    for p in (root, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:  # typeguard >= 3 dropped check_argument_types (used by the vendored stft.py:8,45)
        import typeguard
        if not hasattr(typeguard, "check_argument_types"):
            typeguard.check_argument_types = lambda *a, **k: True
    except ImportError:
        pass
    return root


def load_config(name):
    """name: 'tsh' or 'embed' -> model_params dict of configs/<name>.json."""
    root = _prepare()
    with open(os.path.join(root, "configs", f"{name}.json")) as f:
        return json.load(f)["pl_module_args"]["model_params"]


def reference_net(seed=0, **overrides):
    """The reference separation model (src.models.tfgridnet_realtime.net.Net), eval mode."""
    import torch
    _prepare()
    mod = importlib.import_module("src.models.tfgridnet_realtime.net")
    params = dict(load_config("tsh"))
    params.update(overrides)
    torch.manual_seed(seed)
    return mod.Net(**params).eval()


def reference_embed_net(seed=0, **overrides):
    """The reference enrollment model (src.models.tfgridnet_orig.tfgridnet.EmbedTFGridNet)."""
    import torch
    _prepare()
    mod = importlib.import_module("src.models.tfgridnet_orig.tfgridnet")
    params = dict(load_config("embed"))
    params.update(overrides)
    torch.manual_seed(seed)
    return mod.EmbedTFGridNet(**params).eval()
