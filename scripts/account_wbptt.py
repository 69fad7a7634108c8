"""Dev: per-role wait accounting of the persistent BPTT kernel (build: make -C pytorch_speaker_verification_b200/csrc
ALT=1 ALTFLAGS=-DSVB_WB_ACCOUNT; run with SVB_LIB_PATH=pytorch_speaker_verification_b200/libsvb200_alt.so).  Cycles per
tile that each role thread spends in each of its waits, averaged over the CTAs of a kind, plus the spread over CTAs."""
import ctypes, os, sys
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
base = 2 * nt * 16 + 256
buf = torch.zeros(base + 32 * 120, dtype=torch.int64, device="cuda")
ABL = int(os.environ.get("SVB_ABLATE", "0"))     # parts of the kernel switched off (see scripts/ablate_wbptt.py)
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
a = buf[base:].cpu().numpy().reshape(120, 32).astype(np.float64) / (T * nt)
dur = buf[2 * nt * 16:2 * nt * 16 + 120].cpu().numpy().astype(np.float64) / (T * nt)
names = ["poll:dep_free", "poll:counters", "load:dep", "load:stg_free", "load:c_buf", "tma:dep", "tma:ring_empty", "mma:acc_empty",
         "mma:ring_full", "xch0:buf_free", "xch0:acc_full", "push:send_ready", "xch1:buf_free", "xch1:acc_full", "math:dep",
         "math:own", "math:recv", "math:sync", "math:in_full/x_taken", "store:stg_full", "store:read_done", "store:group"]
kinds = {"R2": range(48, 72), "R1": range(24, 48), "R0": range(0, 24), "X2": range(96, 120), "X1": range(72, 96)}
print("cycles per tile (mean over the 24 CTAs of a kind [min .. max])")
print(f"{'':22s}" + "".join(f"{k:>22s}" for k in kinds))
print(f"{'tile period':22s}" + "".join(f"{dur[list(r)].mean():10.0f} [{dur[list(r)].min():4.0f}..{dur[list(r)].max():4.0f}]" for r in kinds.values()))
for i, n in enumerate(names):
    print(f"{n:22s}" + "".join(f"{a[list(r), i].mean():10.0f} [{a[list(r), i].min():4.0f}..{a[list(r), i].max():4.0f}]" for r in kinds.values()))
smid = buf[base:].cpu().numpy().reshape(120, 32)[:, 31]
print("SM ids of the CTAs:", smid.tolist())
np.set_printoptions(linewidth=250)
for kind in sys.argv[1:]:
    r = list(kinds[kind])
    print(f"per-CTA table of {kind} (rows = slots, columns = the 24 CTAs in grid order: cluster u = 0..5, rank s = 0..3)")
    print(f"{'tile period':22s}" + "".join(f"{v:6.0f}" for v in dur[r]))
    for i, n in enumerate(names):
        print(f"{n:22s}" + "".join(f"{v:6.0f}" for v in a[r, i]))
