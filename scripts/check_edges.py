"""Dev: edge shapes of the persistent kernels (tiny T, ragged / single tiles) against the per-frame kernels."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import ops
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
ops.set_poison_workspace(True)
for (B, T) in ((1, 1), (1, 5), (65, 2), (129, 3), (640, 1), (64, 7), (200, 4), (2, 160)):
    x = torch.tensor(I.logmel(B, T, seed=B * 7 + T)).cuda()
    res = {}
    for mode in (True, False):
        ops.set_persistent(mode); ops.set_persistent_bwd(mode)
        net.zero_grad()
        e = net(x)
        e.square().sum().mul(0.5).add(e.sum()).backward()
        torch.cuda.synchronize()
        res[mode] = (e.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()})
    de = ((res[True][0] - res[False][0]).norm(dim=1) / res[False][0].norm(dim=1)).max().item()
    dg = max(((res[True][1][k] - res[False][1][k]).norm() / res[False][1][k].norm().clamp_min(1e-30)).item() for k in res[True][1])
    nan = int(torch.isnan(res[True][0]).sum()) + sum(int(torch.isnan(v).sum()) for v in res[True][1].values())
    print(f"B={B:4d} T={T:3d}: embedding rel diff {de:.2e}  worst grad rel-L2 diff {dg:.2e}  NaNs {nan}", flush=True)
ops.set_persistent(True); ops.set_persistent_bwd(True); ops.set_poison_workspace(False)
