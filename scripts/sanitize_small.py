"""Dev: small shapes of every new kernel for compute-sanitizer (memcheck): persistent forward / BPTT at B = 20, T = 6
and B = 70, T = 5, the tensor-core GE2E path at 320 x 7, the EER sweep."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
crit = svb.GE2ELoss("cuda")
for (N, M, T) in ((4, 5, 6), (7, 10, 5)):
    x = torch.tensor(I.logmel(N * M, T, seed=3)).cuda()
    net.zero_grad()
    loss = crit(net(x).reshape(N, M, -1)); loss.backward()
    torch.cuda.synchronize()
    print("train step", N, M, T, float(loss), flush=True)
r = np.random.RandomState(7)
E = torch.tensor((r.randn(320, 1, 256) + 0.8 * r.randn(320, 7, 256)).astype(np.float32), device="cuda", requires_grad=True)
l = crit(E); l.backward(); torch.cuda.synchronize()
print("ge2e tc", float(l), flush=True)
enr, ver = I.eer_embeddings(64, 6, 0.06, 0.5, 11)
tup, sim = svb.compute_eer(torch.tensor(enr).cuda(), torch.tensor(ver).cuda())
print("eer", float(tup[0]), flush=True)
