import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lookoncetohear_b200 import Net, synth
from lookoncetohear_b200.configs import TSH_PARAMS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
net = Net(**TSH_PARAMS).eval().cuda()
x, _ = synth.mixture(B, 128 * 3)
e = synth.embedding(B)[:, 0].cuda()
xp = torch.nn.functional.pad(x, (0, 64)).cuda()
st = net.init_buffers(B, "cuda")
with torch.no_grad():
    for i in range(3):
        y, st = net.predict(xp[..., 128 * i:128 * i + 192], e, st, pad=False)
torch.cuda.synchronize()
print("ok", y.shape, float(y.abs().mean()))
