import os, sys, torch, torch.distributed as dist, warnings
warnings.simplefilter("always")
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import dist as sd
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda(); crit = svb.GE2ELoss("cuda")
x = torch.tensor(I.logmel(16, 30, seed=42)).cuda()
emb = net(x); sd.GlobalGE2ELoss(crit)(emb.reshape(4, 4, -1)).backward()
g = [p.grad for p in net.parameters()]
base = sd._flat_view(g)
print(rank, "flat view", None if base is None else (base.numel(), base.is_contiguous()), sum(t.numel() for t in g), flush=True)
local = [t.clone() for t in g]
import time
sd.allreduce_gradients(list(net.parameters()), peer=True)
print(rank, "cache", {k: (v is not None) for k, v in sd._PEER_ALLREDUCE.items()}, flush=True)
for p, l in zip(net.parameters(), local):
    r = l.clone(); dist.all_reduce(r)
    assert torch.allclose(p.grad, r, rtol=1e-6, atol=1e-9), float((p.grad - r).abs().max())
# timing
flat = base
torch.cuda.synchronize()
for fn, name in ((lambda: sd._PEER_ALLREDUCE[next(iter(sd._PEER_ALLREDUCE))](flat), "peer"), (lambda: dist.all_reduce(flat), "nccl")):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(name, "all-reduce of", flat.numel() * 4 / 1e6, "MB:", e0.elapsed_time(e1) / 20 * 1e3, "us", flush=True)
dist.destroy_process_group()
