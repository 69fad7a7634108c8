"""Dev: wait accounting of the persistent forward kernel (cycles per tile each role thread spends in its waits)."""
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
buf = torch.zeros(8192 + 256, dtype=torch.int64, device="cuda")
L.svb_set_trace_mode(2)
for i in range(3):
    if i == 2: L.svb_set_trace(ctypes.c_void_p(buf.data_ptr()))
    with torch.no_grad(): net(x)
    torch.cuda.synchronize()
L.svb_set_trace(None); L.svb_set_trace_mode(1)
a = buf[8192:].cpu().numpy().reshape(4, 8, 8).astype(np.float64) / (T * nt)
roles = {0: ("poller", ["dep_free wait", "poll"]), 1: ("c-loader", ["dep_ready", "stg_full(it-2)"]), 2: ("producer", ["dep_ready", "empty"]),
         3: ("mma", ["acc_empty", "full"]), 4: ("store/signal", ["stg_full|gin_done", "wait_group", "release"]),
         5: ("epilogue grp0", ["dep_ready", "acc_full", "stg_free", "cin_full"]), 6: ("epilogue grp1", ["dep_ready", "acc_full", "stg_free", "cin_full"])}
for r, tag in enumerate(("R(0,0)", "R(1,0)", "P(1,0)")):
    print(tag, "-- cycles per tile (group threads: per tile of the whole sequence, i.e. half of their per-own-tile cost)")
    for k, (name, ws) in roles.items():
        if a[r, k, 4] > 0:
            print(f"   {name:14s} total {a[r, k, 4]:6.0f}  " + "  ".join(f"{w}={a[r, k, i]:.0f}" for i, w in enumerate(ws)))
