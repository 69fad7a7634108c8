"""Warm-up + measured C2 train steps (fwd + GE2E + bwd + fused clip/SGD tail), then one launch each of the persistent
input-projection GEMM, the N = 512 row-sharded GE2E, the C5 EER sweep: the command ncu captures for profiles/."""
import ctypes, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib, eer as E
from pytorch_speaker_verification_b200._lib import ptr

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
L = _lib.lib()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
M, N, K = 640 * 160, 3072, 768
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
C = torch.empty(M, N, device="cuda")
i64 = ctypes.c_int64
assert L.svb_gemm_persistent(ptr(A), ptr(B), ptr(C), None, M, N, K, i64(K), i64(K), i64(N), 0, st) == 0
torch.cuda.synchronize()
Eg = torch.tensor(I.ge2e_embeddings(512, 10, 256, "unit")).cuda()
Cc = svb.get_centroids(Eg)
red, dE = torch.ops.svb200.ge2e_rows(Eg[:64].contiguous(), Cc, crit.w.detach(), crit.b.detach(), 0)
enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
tup, sim = svb.compute_eer(torch.tensor(enr).cuda(), torch.tensor(ver).cuda())
torch.cuda.synchronize()
print("eer", float(tup[0]))
