"""Dev (torchrun, N GPUs): time of the global GE2E forward (loss + saved gradients) per step for the NCCL exchange
(mode="rows") and the peer-memory exchange (mode="peer"), 64 speakers x 10 utterances per rank."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200.dist import GlobalGE2ELoss
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
crit = svb.GE2ELoss("cuda")
E = torch.tensor(I.ge2e_embeddings(64 * world, 10, 256, "unit"))[rank * 64:(rank + 1) * 64].cuda().requires_grad_(True)
res = {}
for mode in ("rows", "peer", "rows", "peer"):
    gl = GlobalGE2ELoss(crit, mode=mode)
    for _ in range(5):
        E.grad = None; gl(E).backward()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        E.grad = None
        loss = gl(E); loss.backward()
    e1.record(); torch.cuda.synchronize()
    res.setdefault(mode, []).append(e0.elapsed_time(e1) / 50 * 1e3)
    val = loss.item()
    if rank == 0: print(f"{mode:5s}: {res[mode][-1]:7.1f} us per fwd+bwd of the global GE2E ({world} ranks), loss {val:.4f}", flush=True)
dist.destroy_process_group()
