"""Dev: clock64 stamps of R(1,0,0) and X(1,0,0) of the persistent BPTT kernel at frame T/2 (cycles, per tile)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
B, T = 640, 160
x = torch.tensor(I.logmel(B, T, seed=1234)).cuda()
nt = (B + 63) // 64
buf = torch.zeros(2 * nt * 16 + 256, dtype=torch.int64, device="cuda")
import os
ABL = int(os.environ.get("SVB_ABLATE", "0"))     # (accounting build only: parts of the kernel switched off)
for i in range(3):
    if i == 2: L.svb_set_trace_bwd(ctypes.c_void_p(buf.data_ptr()))
    net.zero_grad()
    L.svb_set_ablate(0)
    e = net(x); loss = e.square().sum()
    L.svb_set_ablate(ABL)
    loss.backward()
    torch.cuda.synchronize()
L.svb_set_ablate(0)
L.svb_set_trace_bwd(None)
dur = buf[2 * nt * 16:].cpu().numpy()
t = buf[:2 * nt * 16].cpu().numpy().reshape(2, nt, 16).astype(np.float64)
names = ["mma_start", "snd_bufs_ok", "snd_staged", "mma_fullL", "xch_start", "xch_bufs_ok", "xch_accfull", "xch_staged",
         "math_start", "math_own_ok", "math_reduced", "math_in_full", "push_ready", "push_issued", "math_done", "store_done"]
import os
TL = os.environ.get("SVB_TRACE_LAYER", "1")
for r, tag in enumerate((f"R({TL},0,0)", f"X({TL},0,0)")):
    t0 = t[r][t[r] > 0].min()
    print(tag)
    print("      " + " ".join(f"{n[:11]:>11s}" for n in names))
    for j in range(nt):
        print(f"  j={j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:11.0f}" for v in t[r, j]))
print("per-CTA cycles per tile:", {f"R{l}": round(float(dur[l*24:(l+1)*24].mean()) / (T * nt)) for l in range(3)},
      {f"X{l}": round(float(dur[72+(l-1)*24:72+l*24].mean()) / (T * nt)) for l in (1, 2)})
