"""Warm-up + measured C2 train steps (fwd + GE2E + bwd + fused clip/SGD tail) for ncu captures."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
crit = svb.GE2ELoss("cuda")
opt = svb.FusedClipSGD([{"params": net.parameters(), "max_norm": 3.0}, {"params": crit.parameters(), "max_norm": 1.0}],
                       lr=0.01)
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
for i in range(steps):
    opt.zero_grad()
    loss = crit(net(x).reshape(64, 10, 256))
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
print("loss", loss.item())
