"""Dev: clock64 stamps of R(1,0) and P(1,0) of the persistent forward kernel at frame T/2 (cycles, per tile)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
training = len(sys.argv) > 1 and sys.argv[1] == "train"
import os
L.svb_set_ablate(int(os.environ.get("ABLATE", "0")))
net = svb.SpeechEmbedder().cuda()
B, T = 640, 160
x = torch.tensor(I.logmel(B, T, seed=1234)).cuda()
nt = (B + 63) // 64
buf = torch.zeros(4 * nt * 16 + 256, dtype=torch.int64, device="cuda")
for i in range(3):
    if i == 2: L.svb_set_trace(ctypes.c_void_p(buf.data_ptr()))
    if training:
        e = net(x)
    else:
        with torch.no_grad(): e = net(x)
    torch.cuda.synchronize()
L.svb_set_trace(None)
dur = buf[4 * nt * 16:].cpu().numpy()
t = buf[:4 * nt * 16].cpu().numpy().reshape(4, nt, 16).astype(np.float64)
names = ["prod_start", "prod_flags_ok", "prod_issued", "mma_start", "mma_acc_free", "mma_first_full", "mma_last_full",
         "epi_start", "epi_gin_ok", "epi_acc_full", "epi_act_done", "epi_stg_free", "epi_cin_ok", "epi_done",
         "st_begin", "st_done"]
for r, tag in enumerate(("R(1,0)", "P(1,0)")):
    t0 = t[r][t[r] > 0].min()
    print(tag, "cycles after the first stamp of the frame; rows = tiles")
    print("      " + " ".join(f"{n[:12]:>12s}" for n in names))
    for j in range(nt):
        print(f"  j={j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:12.0f}" for v in t[r, j]))

print("store warp detail R(1,0): got_full, committed, read_done, prev_complete, proxy_fence, threadfence, red")
t0 = t[0][t[0] > 0].min()
for j in range(nt):
    print(f"  j={j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:9.0f}" for v in t[2, j, :7]))

print("MMA detail R(1,0): (before_wait, after_wait) x 6 groups, issued_all | producer issue time of groups 0,1,2")
for j in range(nt):
    print(f"  j={j:2d} " + " ".join(f"{(v - t0) if v > 0 else -1:7.0f}" for v in t[3, j, :16]))

print("per-CTA kernel-loop cycles per tile (mean over the 24 slices): ",
      {f"R{l}": round(float(dur[l*24:(l+1)*24].mean()) / (T * nt)) for l in range(3)},
      {f"P{l}": round(float(dur[72+l*24:72+(l+1)*24].mean()) / (T * nt)) for l in range(3)})
