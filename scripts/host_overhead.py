"""Where does the train step's time go between the kernels?  (1) host time per step without synchronising (is the host
ahead of the GPU?), (2) device time per step in a back-to-back loop, (3) a torch.profiler (kineto/CUPTI) trace of three
steps written to gpurun_out/ and summarised: GPU busy time, idle gaps > 5 us and the kernels on either side."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import _inputs as I
import pytorch_speaker_verification_b200 as svb

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = svb.SpeechEmbedder().to(dev)
crit = svb.GE2ELoss(dev)
params = list(net.parameters()) + [crit.w, crit.b]
x = torch.tensor(I.logmel(640, 160, seed=1234)).to(dev)


def step():
    for p in params:
        p.grad = None
    loss = crit(net(x).reshape(64, 10, 256))
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
host = []
e0.record()
for _ in range(n):
    t0 = time.perf_counter()
    step()
    host.append(time.perf_counter() - t0)
e1.record()
torch.cuda.synchronize()
print(f"device ms/step (back to back): {e0.elapsed_time(e1) / n:.3f}; host ms/step: median {sorted(host)[n // 2] * 1e3:.3f} "
      f"min {min(host) * 1e3:.3f} max {max(host) * 1e3:.3f}")

out = os.path.join(ROOT, "gpurun_out", "r2_step_trace.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
kern = sorted([e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")], key=lambda e: e["ts"])
print("gpu activities:", len(kern))
t_first, t_last = kern[0]["ts"], max(e["ts"] + e["dur"] for e in kern)
busy_end = kern[0]["ts"]
gaps = []
busy = 0.0
for e in kern:
    if e["ts"] > busy_end:
        gaps.append((e["ts"] - busy_end, prev["name"][:50], e["name"][:50]))
        busy_end = e["ts"]
    if e["ts"] + e["dur"] > busy_end:
        busy += e["ts"] + e["dur"] - max(busy_end, e["ts"])
        busy_end = e["ts"] + e["dur"]
        prev = e
print(f"span {(t_last - t_first) / 3e3:.3f} ms/step, busy {busy / 3e3:.3f} ms/step, idle {(t_last - t_first - busy) / 3e3:.3f} ms/step")
for g in sorted(gaps, reverse=True)[:25]:
    print(f"  gap {g[0]:8.1f} us  after {g[1]:50s} before {g[2]}")
# host-side cost of the custom ops / library calls
cpu = [e for e in ev if e.get("cat") in ("cpu_op", "user_annotation", "python_function") and e.get("dur", 0) > 20]
agg = {}
for e in cpu:
    agg.setdefault(e["name"][:70], []).append(e["dur"])
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:25]:
    print(f"  cpu {k:70s} n={len(v):3d} total {sum(v) / 3e3:8.3f} ms/step")
