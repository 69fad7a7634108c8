import ctypes, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "scripts")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
Eg = torch.tensor(I.ge2e_embeddings(512, 10, 256, "unit")).cuda().requires_grad_(True)
crit = svb.GE2ELoss("cuda")
for _ in range(3):
    Eg.grad = None
    crit(Eg).backward()
torch.cuda.synchronize()
