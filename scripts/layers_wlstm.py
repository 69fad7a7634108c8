"""Dev: forward time of the persistent kernel for 1, 2, 3 layers (isolates cross-layer coupling)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import hparam as H
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
for L in (1, 2, 3):
    H.configure(num_layer=L)
    torch.manual_seed(0)
    net = svb.SpeechEmbedder().cuda()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        net(x); net(x); t0.record()
        for _ in range(3): net(x)
        t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    print(f"L={L}: forward {ms:.3f} ms  -> {ms * 1e-3 * 1.85e9 / (160 * 10):.0f} cycles per (frame, tile) at 1.85 GHz", flush=True)
H.configure(num_layer=3)
