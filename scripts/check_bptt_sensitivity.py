import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import ops
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
crit = svb.GE2ELoss("cuda")
N, M, T = 4, 4, 30
x = torch.tensor(I.logmel(N * M, T, seed=42)).cuda()
def grads(noise, persistent=True):
    ops.set_persistent_bwd(persistent)
    net.zero_grad(); crit.zero_grad()
    emb = net(x)
    loss = crit(emb.reshape(N, M, -1))
    g = torch.autograd.grad(loss, emb, retain_graph=False)[0]
    torch.manual_seed(1)
    g = g * (1 + noise * torch.randn_like(g))
    net.zero_grad()
    emb = net(x)
    emb.backward(g)
    return {k: p.grad.clone() for k, p in net.named_parameters()}
for persistent in (True, False):
    a = grads(0.0, persistent); b = grads(1e-6, persistent); c = grads(1e-4, persistent)
    print("persistent" if persistent else "per-frame")
    for k in a:
        print(f"   {k:30s} |g| {float(a[k].norm()):.3e}  rel diff at 1e-6 noise {float((a[k]-b[k]).norm()/a[k].norm()):.2e}   at 1e-4 noise {float((a[k]-c[k]).norm()/a[k].norm()):.2e}")
ops.set_persistent_bwd(True)
