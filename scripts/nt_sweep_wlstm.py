"""Dev: frame time of the persistent forward kernel as a function of the number of 64-row tiles per frame."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
T = 160
for B in (64, 128, 256, 384, 512, 640, 960, 1280, 2560):
    x = torch.tensor(I.logmel(B, T, seed=1)).cuda()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        net(x); net(x); t0.record()
        for _ in range(3): net(x)
        t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 3
    nt = (B + 63) // 64
    print(f"B={B:5d} nt={nt:3d}: forward {ms:.3f} ms  frame {ms * 1e3 / (T + 2):.2f} us  per tile {ms * 1e3 / (T + 2) / nt:.2f} us", flush=True)
