"""Dev: step time against the share of frames whose weight-gradient slices run beside the BPTT kernel (in-process
A/B, interleaved so that clock / thermal drift hits every setting alike)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import ops, _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
crit = svb.GE2ELoss("cuda")
x = torch.tensor(I.logmel(640, 160, seed=1)).cuda()
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")

def step():
    net.zero_grad(set_to_none=True)
    crit(net(x).reshape(64, 10, -1)).backward()

def measure(n=6):
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)

for _ in range(5): step()
settings = [0, 15, 20, 25, 30, 35, 40, 50]
acc = {k: [] for k in settings}
for rnd in range(4):
    for k in settings:
        ops.set_wgrad_overlap(k > 0)
        if k: L.svb_set_wgrad_late_pct(k)
        step()
        acc[k].append(measure())
for k in settings:
    print(f"late {k:2d}%: " + " ".join(f"{v:.3f}" for v in acc[k]) + f"   mean {sum(acc[k]) / len(acc[k]):.3f} ms")
