import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import ops
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
for (B, T) in ((16, 30), (20, 6), (64, 30), (65, 30), (128, 12), (640, 160)):
    x = torch.tensor(I.logmel(B, T, seed=9)).cuda()
    gs = []
    for rep in range(4):
        net.zero_grad()
        e = net(x); (e.square().sum() + e.sum()).backward()
        gs.append({k: p.grad.clone() for k, p in net.named_parameters()})
    ops.set_persistent_bwd(False)
    net.zero_grad(); e = net(x); (e.square().sum() + e.sum()).backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters()}
    ops.set_persistent_bwd(True)
    worst = max(float((gs[0][k] - gs[r][k]).abs().max()) for r in range(1, 4) for k in gs[0])
    rel = max(float((gs[0][k] - ref[k]).norm() / ref[k].norm()) for k in ref)
    print(f"B={B} T={T}: rerun max|diff| {worst:.3e}; vs per-frame path max rel-L2 {rel:.3e}", flush=True)
