"""Dev: time the persistent BPTT kernel with parts switched off (results are garbage in those runs).  Needs the
accounting build: make -C pytorch_speaker_verification_b200/csrc ALT=1 ALTFLAGS=-DSVB_WB_ACCOUNT, SVB_LIB_PATH=.../libsvb200_alt.so."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
import ctypes
buf = (ctypes.c_float * 16)()
# (the wait-skipping bits >= 512 can fault when combined with live pushes: use them one at a time on top of 447 + 64)
for mask in (0, 1, 2, 4, 8, 16, 32, 128, 256, 2 + 256, 1 + 4, 2 + 32 + 128 + 256, 1 + 4 + 16, 447 - 8, 447, 511, 511 + 1024 + 2048 + 4096, 511 + 16384 + 32768 + 65536):
    res = []
    for rep in range(3):
        net.zero_grad()
        L.svb_set_ablate(0)
        e = net(x); loss = e.square().sum()
        L.svb_set_ablate(mask)
        L.svb_profile_enable(1)
        loss.backward()
        torch.cuda.synchronize()
        L.svb_profile_read(buf, 16)
        res.append(buf[5])
    print(f"ablate={mask:3d} (1 MMA, 2 math, 4 operand loads, 8 deps, 16 staging stores, 32 input loads, 64 push, 128 stores, 256 math fence; skeleton waits: 512 xch buffers, 1024 in_full, 2048 loader, 4096 publish, 8192 own, 16384 acc_empty, 32768 ring, 65536 acc_full, 131072 math bar): BPTT kernel {min(res):.3f} ms", flush=True)
L.svb_set_ablate(0); L.svb_profile_enable(0)
