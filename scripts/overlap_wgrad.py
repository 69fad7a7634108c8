"""Dev: does the late-frame weight-gradient work really run beside the BPTT kernel?  Event times on both streams
(SVB_WGRAD_DEBUG=1) and the step time with the overlap on / off."""
import ctypes, os, sys
os.environ["SVB_WGRAD_DEBUG"] = "1"
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
    loss = crit(net(x).reshape(64, 10, -1))
    loss.backward()

for mode in (1, 0, 1, 0):
    ops.set_wgrad_overlap(mode)
    for _ in range(3): step()
    ts = []
    for _ in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = (ctypes.c_float * 4)()
    msg = ""
    if mode and L.svb_wgrad_overlap_timing(out) == 0:
        msg = "  since fork: gate open %.2f ms, side done %.2f, BPTT done %.2f, backward done %.2f" % tuple(out)
    print(f"overlap={mode}: step {sorted(ts)[len(ts)//2]:.3f} ms (min {min(ts):.3f})" + msg)
