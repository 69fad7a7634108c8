"""Dev: forward kernel time per 64-row tile-step for several (B, T, training) shapes."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
def run(B, T, train):
    x = torch.tensor(I.logmel(min(B, 4096), T, seed=B + T)).cuda()
    if B > 4096:
        x = x.repeat((B + 4095) // 4096, 1, 1)[:B].contiguous()
    def f():
        if train:
            return net(x)
        with torch.no_grad():
            return net(x)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        del y
    ms = min(ts)
    nt = (B + 63) // 64
    print(f"B={B:6d} T={T:4d} train={int(train)}: {ms:8.3f} ms  -> {ms * 1e3 / (T * nt):6.3f} us per tile-step, "
          f"{B * T * 23838720 * 2 / 2 / ms / 1e9:7.1f} TFLOP/s")
for B, T, tr in ((640, 160, True), (640, 160, False), (640, 24, False), (2560, 24, False), (6400, 24, False),
                 (6400, 160, False), (25600, 24, False), (46797, 24, False), (46797, 24, True)):
    run(B, T, tr)
