"""Which stage bounds the one-hop pipeline?  Re-times the pipelined stream with single stages NOT launched
(engine option pipeline_debug_skip; outputs are garbage, only the schedule is of interest): the stage whose
removal buys the most time is the bottleneck.   python tools/pipe_skip_experiment.py [A:Q:T:O:F:B]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lookoncetohear_b200 import Net, synth
from lookoncetohear_b200.configs import TSH_PARAMS

STAGES = ["front", "gemm_ih", "lstm", "mid_a", "mid_b", "mid_c", "qkv", "attn", "attn_out", "back"]
names = ("pipeline_lanes", "pipeline_qkv_lanes", "pipeline_attn_lanes", "pipeline_out_lanes", "pipeline_front_lanes",
         "pipeline_back_lanes")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**TSH_PARAMS).eval().to(dev)
x, _ = synth.mixture(1, 64000)
x = x.to(dev)
emb = synth.embedding(1)[:, 0].to(dev)
y = torch.empty(1, 2, 64000, device=dev)
combo = sys.argv[1] if len(sys.argv) > 1 else "8:3:2:4:4:4"
for n, v in zip(names, [int(v) for v in combo.split(":")]):
    net.set_option(n, v)


def run(mask):
    net.set_option("pipeline_debug_skip", mask)
    best = None
    for it in range(4):
        st = net.init_buffers(1, dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        net.stream_dev(x, emb, chunks_per_call=1, state=st, n_calls=500, out=y)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if it > 0 and (best is None or ms < best):
            best = ms
    return 1e3 * best / 500


ALL = 1023
bits = {n: 1 << i for i, n in enumerate(STAGES)}
base = run(0)
print(json.dumps({"lanes": combo, "launched": "all", "us_per_hop": round(base, 2)}), flush=True)
for i, s in enumerate(STAGES):
    t = run(1 << i)
    print(json.dumps({"skipped": s, "us_per_hop": round(t, 2), "gain_us": round(base - t, 2)}), flush=True)
groups = [("front",), ("gemm_ih", "lstm", "mid_a"), ("mid_b",), ("mid_c", "qkv"), ("attn",), ("attn_out",), ("back",)]
for g in groups:                                   # one stage group alone: its own floor
    keep = sum(bits[n] for n in g)
    print(json.dumps({"only": "+".join(g), "us_per_hop": round(run(ALL & ~keep), 2)}), flush=True)
keep = 0
for g in groups:                                   # cumulative: where does the time appear?
    keep |= sum(bits[n] for n in g)
    print(json.dumps({"cumulative_up_to": "+".join(g), "us_per_hop": round(run(ALL & ~keep), 2)}), flush=True)
net.set_option("pipeline_debug_skip", 0)
