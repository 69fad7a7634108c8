"""Dev: mean per-phase device time of the C2 train step (library CUDA events), weight-gradient overlap off so that
the BPTT kernel is timed alone.  A/B a build variant with SVB_LIB_PATH=pytorch_speaker_verification_b200/libsvb200_alt.so."""
import ctypes, os, sys
os.environ.setdefault("SVB_WGRAD_OVERLAP", "0")
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
crit = svb.GE2ELoss("cuda")
x = torch.tensor(I.logmel(640, 160, seed=1)).cuda()
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
def step():
    net.zero_grad(set_to_none=True)
    crit(net(x).reshape(64, 10, -1)).backward()
for _ in range(4): step()
names = ["prep", "input_gemm", "recurrent_fwd", "projection", "projection_bwd", "recurrent_bwd", "weight_grads", "bias", "dx"]
acc = [0.0] * 9
buf = (ctypes.c_float * 16)()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for _ in range(n):
    flush.fill_(1)
    L.svb_profile_enable(1)
    step(); torch.cuda.synchronize()
    L.svb_profile_read(buf, 16)
    for i in range(9): acc[i] += buf[i]
L.svb_profile_enable(0)
print(os.path.basename(_lib.LIB_PATH), " ".join(f"{names[i]}={acc[i]/n:.3f}" for i in (2, 5, 6)), f"sum={sum(acc)/n:.3f} ms")
