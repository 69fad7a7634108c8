"""Persistent wavefront forward kernel vs the per-frame kernels: same embeddings, same BPTT stash; timing."""
import os, sys, time
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
        ops.set_persistent(mode)
        with torch.no_grad():
            e = net(x)
        torch.cuda.synchronize()
        net.zero_grad()
        e2 = net(x)
        e2.square().sum().mul(0.5).add(e2.sum()).backward()
        torch.cuda.synchronize()
        out[mode] = (e, e2.detach(), net.LSTM_stack.weight_hh_l0.grad.clone(), net.LSTM_stack.weight_ih_l2.grad.clone(),
                     net.LSTM_stack.bias_ih_l1.grad.clone())
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            net(x); t0.record()
            for _ in range(3): net(x)
            t1.record(); torch.cuda.synchronize()
        print(f"B={B} T={T} persistent={mode}: inference forward {t0.elapsed_time(t1)/3:.3f} ms", flush=True)
    for k, (a, b) in enumerate(zip(out[True], out[False])):
        print(f"   out[{k}] max|diff| {(a-b).abs().max().item():.3e}  max|ref| {b.abs().max().item():.3e}", flush=True)
ops.set_persistent(True)
