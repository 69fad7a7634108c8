"""Dev: the persistent kernel must be run-to-run identical and independent of the batch composition."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
import os
from pytorch_speaker_verification_b200 import _lib
_lib.lib().svb_set_ablate(int(os.environ.get("ABLATE", "0")))
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
for (B, T) in ((35, 24), (50, 40), (100, 40), (200, 30)):
    x = torch.tensor(I.logmel(B, T, seed=9)).cuda()
    with torch.no_grad():
        for rep in range(60):
            bad = int(torch.isnan(net(x)).sum())
            if bad: print("   rep", rep, "NaNs", bad)
        a = net(x); b = net(x)
        sub = net(x[:B // 3])
        perm = torch.randperm(B, device="cuda")
        pm = net(x[perm])
    print("   NaNs:", [int(torch.isnan(v).sum()) for v in (a, b, sub, pm)])
    print(f"B={B} T={T}: rerun max|diff| {(a-b).abs().max().item():.3e}  sub-batch {(sub - a[:B//3]).abs().max().item():.3e} "
          f"perm {(pm - a[perm]).abs().max().item():.3e}", flush=True)
    net.zero_grad()
    e = net(x); e.square().sum().backward(); g1 = net.LSTM_stack.weight_hh_l1.grad.clone()
    net.zero_grad()
    e = net(x); e.square().sum().backward(); g2 = net.LSTM_stack.weight_hh_l1.grad.clone()
    print(f"   training rerun: grad max|diff| {(g1-g2).abs().max().item():.3e} (max {g1.abs().max().item():.3e})", flush=True)
