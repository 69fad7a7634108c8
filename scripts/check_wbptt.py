"""Persistent BPTT kernel vs the per-frame BPTT kernels: same parameter gradients; timing of the whole backward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import ops

torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
shapes = [(20, 6), (150, 40), (640, 160)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (B, T) in shapes:
    x = torch.tensor(I.logmel(B, T, seed=5)).cuda()
    out = {}
    for mode in (False, True):
        ops.set_persistent_bwd(mode)
        for rep in range(2):
            net.zero_grad()
            e = net(x)
            loss = e.square().sum().mul(0.5).add(e.sum())
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record(); loss.backward(); t1.record(); torch.cuda.synchronize()
        out[mode] = {k: p.grad.clone() for k, p in net.named_parameters()}
        print(f"B={B} T={T} persistent_bwd={mode}: backward {t0.elapsed_time(t1):.3f} ms", flush=True)
    for k in out[True]:
        a, b = out[True][k], out[False][k]
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        print(f"   {k:32s} rel-L2 diff {rel:.3e}  nan={int(torch.isnan(a).sum())}", flush=True)
ops.set_persistent_bwd(True)
