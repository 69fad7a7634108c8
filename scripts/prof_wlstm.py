"""Dev: a few inference forwards of the persistent kernel at B=640 (short T) for ncu."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
torch.manual_seed(0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 24
net = svb.SpeechEmbedder().cuda()
x = torch.tensor(I.logmel(640, T, seed=1234)).cuda()
for i in range(3):
    with torch.no_grad():
        e = net(x)
    torch.cuda.synchronize()
print("ok", float(e.sum()))
