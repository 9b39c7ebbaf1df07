"""Timeline of ONE one-hop chain (the latency path: a single 8 ms chunk, B = 1) from the device-side trace: per kernel its
entry and exit (first thread of CTA 0, %globaltimer) relative to the first kernel's entry.   python tools/hop_trace.py [repeats]"""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lookoncetohear_b200 import Net, synth, _cabi
from lookoncetohear_b200.configs import TSH_PARAMS

NAMES = ["front", "gemm_ih", "lstm", "mid_a", "mid_b", "mid_c", "qkv", "attn", "attn_out", "back", "mid", "tail"]
REC = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("ptr", "<u8"), ("kernel", "<u4"), ("sm", "<u4")])
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 9
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = Net(**TSH_PARAMS).eval().to(dev)
L = _cabi.lib()
HOPS = 80
x, _ = synth.mixture(1, 128 * HOPS)
x = x.to(dev)
emb = synth.embedding(1)[:, 0].to(dev)
y = torch.empty(1, 2, 128 * HOPS, device=dev)
st = net.init_buffers(1, dev)
for h in range(60):                         # warm: rings full, graphs instantiated
    net.stream_dev(x[..., 128 * h:], emb, chunks_per_call=1, state=st, n_calls=1, out=y[..., 128 * h:])
torch.cuda.synchronize()
# untraced: back-to-back one-hop calls (the host runs ahead of the device), device time per hop = chain + the gap between two graphs
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
per_hop = []
for rep in range(5):
    st2 = net.init_buffers(1, dev)
    for h in range(10):
        net.stream_dev(x[..., 128 * h:], emb, chunks_per_call=1, state=st2, n_calls=1, out=y[..., 128 * h:])
    torch.cuda.synchronize()
    e0.record()
    for h in range(10, 70):
        net.stream_dev(x[..., 128 * h:], emb, chunks_per_call=1, state=st2, n_calls=1, out=y[..., 128 * h:])
    e1.record()
    torch.cuda.synchronize()
    per_hop.append(e0.elapsed_time(e1) * 1e3 / 60)
print(json.dumps({"untraced_back_to_back_us_per_hop": [round(v, 1) for v in sorted(per_hop)]}))
runs = []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(reps):
    h = 60 + r
    _cabi.check(L.l2h_sep_trace_start(net._engine(), 256))
    torch.cuda.synchronize()
    ev0.record()
    net.stream_dev(x[..., 128 * h:], emb, chunks_per_call=1, state=st, n_calls=1, out=y[..., 128 * h:])
    ev1.record()
    torch.cuda.synchronize()
    buf = np.zeros(256, dtype=REC)
    n = ctypes.c_int32()
    _cabi.check(L.l2h_sep_trace_read(net._engine(), buf.ctypes.data_as(ctypes.c_void_p), 256, ctypes.byref(n)))
    rec = buf[:n.value]
    rec = rec[rec["t1"] > 0]
    rec = rec[np.argsort(rec["t0"])]
    org = int(rec["t0"].min())
    runs.append((float(rec["t1"].max() - org) / 1e3, ev0.elapsed_time(ev1) * 1e3, rec, org))
_cabi.check(L.l2h_sep_trace_start(net._engine(), 0))
runs.sort(key=lambda t: t[0])
span, ev_us, rec, org = runs[len(runs) // 2]
print(json.dumps({"kernels": int(len(rec)), "span_us_median": round(span, 1), "event_us_same_run": round(ev_us, 1),
                  "span_us_all": [round(r[0], 1) for r in runs]}))
print("| # | kernel | entry us | exit us | in-kernel us | exit - previous exit us |")
print("|---|---|---|---|---|---|")
prev = 0.0
prev_pt = 0.0
for i, r in enumerate(rec):
    a, b = (int(r["t0"]) - org) / 1e3, (int(r["t1"]) - org) / 1e3
    k = int(r["kernel"])
    if k >= 100:        # a time stamp inside the kernel (TraceScope::mark)
        print("| %d | . %s point %d | %.1f | | | %.1f |" % (i, NAMES[(k - 100) // 20], (k - 100) % 20, a, a - prev_pt))
        prev_pt = a
        continue
    print("| %d | %s | %.1f | %.1f | %.1f | %.1f |" % (i, NAMES[k], a, b, b - a, b - prev))
    prev = b
