"""Dev: A/B of library builds on ONE box: each variant (SVB_LIB_PATH) in its own process, interleaved, phase times of
the C2 train step (ms): recurrent_fwd, recurrent_bwd, weight_grads, whole value step."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes, sys, os, json
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import torch, _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda(); crit = svb.GE2ELoss("cuda")
x = torch.tensor(I.logmel(640, 160, seed=1234)).cuda()
params = list(net.parameters()) + [crit.w, crit.b]
def step():
    for p in params: p.grad = None
    crit(net(x).reshape(64, 10, 256)).backward()
for _ in range(4): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(12): step()
e1.record(); torch.cuda.synchronize()
buf = (ctypes.c_float * 16)(); acc = [0.0] * 9
for _ in range(4):
    L.svb_profile_enable(1); step(); torch.cuda.synchronize(); L.svb_profile_read(buf, 16)
    for i in range(9): acc[i] += buf[i] / 4
L.svb_profile_enable(0)
print(json.dumps({"step": e0.elapsed_time(e1) / 12, "fwd": acc[2], "bwd": acc[5], "wgrad": acc[6]}))
''' % (ROOT, ROOT)
variants = sys.argv[1:]
res = {v: [] for v in variants}
for rep in range(3):
    for v in variants:
        env = dict(os.environ, SVB_LIB_PATH=os.path.join(ROOT, "pytorch_speaker_verification_b200", v))
        r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=300)
        try:
            res[v].append(json.loads(r.stdout.strip().splitlines()[-1]))
        except Exception:
            print(v, "failed:", r.stderr[-1500:])
for v in variants:
    for k in ("step", "fwd", "bwd", "wgrad"):
        vals = [round(x[k], 3) for x in res[v]]
        print(f"{v:24s} {k:6s} {vals}")
