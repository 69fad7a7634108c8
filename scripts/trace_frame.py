"""Dev: per-CTA globaltimer stamps of one forward and one backward frame kernel (layer 1, t = T/2)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda(); crit = svb.GE2ELoss("cuda")
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
for i in range(3):
    if i == 2: L.svb_set_trace(ctypes.c_void_p(buf.data_ptr()))
    for p in net.parameters(): p.grad = None
    crit(net(x).reshape(64, 10, 256)).backward()
    torch.cuda.synchronize()
L.svb_set_trace(None)
names = ["start", "prod_begin", "prod_done", "mma_first", "mma_last", "epi_in_ready", "epi_acc_ready", "epi_math_done",
         "epi_staged", "stores_done", "exit"]
for tag, off in (("FWD frame", 0), ("BWD frame", 4096)):
    t = buf[off:off + 120 * 16].cpu().numpy().reshape(120, 16).astype(np.float64)
    t0 = t[:, 0].min()
    print(tag, "(us after first CTA start; mean / min / max over 120 CTAs)")
    for i, n in enumerate(names):
        v = (t[:, i] - t0) / 1e3
        print(f"  {n:14s} {v.mean():7.2f} {v.min():7.2f} {v.max():7.2f}")
