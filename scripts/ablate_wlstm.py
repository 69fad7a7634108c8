"""Dev: time the persistent forward kernel with parts switched off (results are garbage in those runs)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
for mask in (0, 7, 8, 15, 9, 10, 12):
    L.svb_set_ablate(mask)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        net(x); t0.record()
        for _ in range(3): net(x)
        t1.record(); torch.cuda.synchronize()
    print(f"ablate={mask} (1 no MMA, 2 no epilogue math, 4 no operand loads, 8 no cross-CTA deps): {t0.elapsed_time(t1)/3:.3f} ms", flush=True)
L.svb_set_ablate(0)
