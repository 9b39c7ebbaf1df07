"""Per-kernel device times (l2h_sep_profile: CUDA events between launches, no graph) for several
call sizes.  Usage: python tools/profile_kernels.py [cpc ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from lookoncetohear_b200 import Net, synth
from lookoncetohear_b200.configs import TSH_PARAMS

args = [a for a in sys.argv[1:] if not a.startswith("B=")]
batch = int(([a[2:] for a in sys.argv[1:] if a.startswith("B=")] or ["1"])[0])
cpcs = [int(a) for a in args] or [1, 20, 500]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**TSH_PARAMS).eval().to(dev)
x, _ = synth.mixture(1, 64000 + 64)
x = x.to(dev)
emb = synth.embedding(1)[:, 0].to(dev)
for c in cpcs:
    prof = bench.profile_chain(net, x, emb, dev, c, iters=10, batch=batch)
    tot = sum(v["ms_total"] for v in prof.values())
    print(json.dumps({"batch": batch, "cpc": c, "chain_us": round(1e3 * tot, 1), "frames_per_s_unpipelined": round(batch * c / (tot * 1e-3)),
                      "kernel_us_per_chain": {k: round(1e3 * v["ms_total"], 1) for k, v in prof.items()}}))
