"""GPU debugging aid: run the separation engine with taps and print the rel-L2 error of every
stage against the CPU oracle (oracle/restate.py).  Usage: python tools/debug_sep.py [T] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lookoncetohear_b200 import Net, synth
from oracle import restate as rs
from lookoncetohear_b200.configs import TSH_PARAMS as TSH


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.manual_seed(0)
    net = Net(**TSH).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x, _ = synth.mixture(B, 128 * T - 37)
    e = synth.embedding(B)
    taps_ref = {}
    st_ref = rs.sep_init_state(sd, B)
    y_ref, st_ref = rs.sep_predict(sd, x, e[:, 0], st_ref, pad=True, taps=taps_ref)
    net = net.cuda()
    with torch.no_grad():
        y, taps, st = net.forward_with_taps(x.cuda(), e.cuda())
    torch.cuda.synchronize()
    names = ["enc"] + [f"b{b}.{s}" for b in range(3) for s in ("intra", "inter", "out")]
    ref_list = [taps_ref["enc"]] + [taps_ref[f"block{b}{s}"] for b in range(3) for s in ("_intra", "_inter", "")]
    ref_list[3] = ref_list[3] * taps_ref["gate"]        # the engine folds the speaker gate into block 0's epilogue
    for n, t, r in zip(names, taps, ref_list):
        print(f"{n:10s} rel_l2 {rs.rel_l2(t.cpu(), r):.3e}   |ref| {r.abs().mean():.4f}  |out| {t.abs().mean().item():.4f}")
    print("gate      rel_l2 %.3e" % rs.rel_l2(st._rec()[:, 256:256 + 6208].cpu().view(B, 97, 64), taps_ref["gate"][:, 0]))
    print("y         rel_l2 %.3e" % rs.rel_l2(y.cpu(), y_ref), tuple(y.shape), tuple(y_ref.shape))
    sr = st.to_reference()
    for k in ("conv_buf", "deconv_buf", "istft_buf"):
        print(f"state {k:10s} rel_l2 {rs.rel_l2(sr[k].cpu(), st_ref[k]):.3e}")
    for k in ("K_buf", "V_buf", "h0", "c0"):
        print(f"state buf2.{k:6s} rel_l2 {rs.rel_l2(sr['gridnet_bufs']['buf2'][k].cpu(), st_ref['gridnet_bufs']['buf2'][k]):.3e}")
    print("header", st.header())


if __name__ == "__main__":
    main()
