"""Enrollment network (BASELINE configs[3] shape: 5 s binaural utterances): utterances/s on one GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lookoncetohear_b200 import EmbedTFGridNet, synth
from lookoncetohear_b200.configs import EMBED_PARAMS

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = EmbedTFGridNet(**EMBED_PARAMS).eval().to(dev)
batches = [int(a) for a in sys.argv[1:]] or [1, 8, 32]
for B in batches:
    x = synth.enrollment(B, 80000).to(dev)
    with torch.no_grad():
        net(x); torch.cuda.synchronize()
        best = None
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); net(x); b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b); best = ms if best is None else min(best, ms)
    print(json.dumps({"batch": B, "ms": round(best, 2), "utt_per_s": round(B / (best * 1e-3), 2),
                      "tflops_fp32_algorithmic": round(B * 255e9 / (best * 1e-3) / 1e12, 2)}))
