"""Dev: persistent forward kernel in TRAINING mode (stash written) with parts switched off; phase 2 = recurrent fwd."""
import ctypes, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
buf = (ctypes.c_float * 16)()
for mask in (0, 1, 2, 4, 3, 5, 6, 7, 8, 15):
    res = []
    for rep in range(3):
        L.svb_set_ablate(mask)
        L.svb_profile_enable(1)
        e = net(x)
        torch.cuda.synchronize()
        L.svb_profile_read(buf, 16)
        res.append(buf[2])
        del e
    print(f"ablate={mask:2d} (1 no MMA, 2 no epilogue math/stores, 4 no operand loads, 8 no deps): fwd kernel {min(res):.3f} ms", flush=True)
L.svb_set_ablate(0); L.svb_profile_enable(0)
