"""Generate the committed fixtures from the REFERENCE ITSELF (run in the build container, where
/root/reference exists):  python tests/golden/make_golden.py

Weights are not stored (8 MB): they are the PyTorch default init under torch.manual_seed(seed),
which the reference modules and the engine's parameter containers reproduce identically
(tests/test_oracle.py::test_seeded_init_matches_reference); a checksum of the weights is stored
so that an RNG drift between torch builds is detected rather than misread as a parity failure.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lookoncetohear_b200 import synth  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def weight_checksum(sd):
    return np.array([float(sum(v.double().abs().sum() for v in sd.values())),
                     float(sum((v.double() ** 2).sum() for v in sd.values()))])


def main():
    torch.set_num_threads(8)
    # ---- separation: whole utterance (ragged length) + chunked streaming, B=2 ----------------
    seed = 0
    net = rl.reference_net(seed)
    B, N = 2, 128 * 14 - 51
    x, tgt = synth.mixture(B, N)
    e = synth.embedding(B)
    with torch.no_grad():
        y = net(x, e)
        st = net.init_buffers(B, "cpu")
        xp = torch.nn.functional.pad(x, (0, 128 * 14 - N + 64))
        ys = torch.cat([net.predict(xp[..., 128 * i:128 * i + 192], e[:, 0], st, pad=False)[0] for i in range(14)], -1)
    np.savez_compressed(os.path.join(HERE, "sep_golden.npz"), seed=seed, B=B, N=N, y=y.numpy(),
                        y_stream=ys.numpy(), h0_buf2=st["gridnet_bufs"]["buf2"]["h0"].numpy(),
                        istft_buf=st["istft_buf"].numpy(), wsum=weight_checksum(net.state_dict()))
    # ---- separation: longer than the attention window (T = 70 > 50), B=1, keep only a digest ---
    N2 = 128 * 70
    x2, _ = synth.mixture(1, N2, seed0=1100)
    e2 = synth.embedding(1, seed0=3100)
    with torch.no_grad():
        y2 = net(x2, e2)
    np.savez_compressed(os.path.join(HERE, "sep_golden_long.npz"), seed=seed, N=N2, y_tail=y2[..., -1024:].numpy(),
                        y_rms=float(y2.pow(2).mean().sqrt()), y_sum=float(y2.double().sum()))
    # ---- enrollment ------------------------------------------------------------------------------
    en = rl.reference_embed_net(seed)
    xe = synth.enrollment(2, 4800)
    with torch.no_grad():
        emb = en(xe)
    np.savez_compressed(os.path.join(HERE, "embed_golden.npz"), seed=seed, n=4800, emb=emb.numpy(),
                        wsum=weight_checksum(en.state_dict()))
    # ---- state_dict key names and shapes of both reference modules (checkpoint compatibility) ----
    import json
    keys = {"sep": {k: list(v.shape) for k, v in net.state_dict().items()},
            "embed": {k: list(v.shape) for k, v in en.state_dict().items()}}
    with open(os.path.join(HERE, "ckpt_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)
    print("written", os.listdir(HERE))


if __name__ == "__main__":
    main()
