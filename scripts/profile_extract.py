"""Dev: where does the end-to-end extraction time go?  Wall time, host enqueue time (until the final synchronize
starts) and the device-resident LSTM time for the same windows, for several chunk sizes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import dvector as D

torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda().eval()
r = np.random.RandomState(4321)
Ts = r.randint(100, 501, size=2000)
base = [np.log10(I.power_spec(int(T), seed=int(T) + 7 * i) + 1e-6).astype(np.float32) for i, T in enumerate(Ts[:64])]
specs = [base[i % 64][:, :int(T)] if base[i % 64].shape[1] >= T else np.tile(base[i % 64], (1, 8))[:, :int(T)]
         for i, T in enumerate(Ts)]
specs = [np.ascontiguousarray(s) for s in specs]
nwin = int(sum(max(0, -(-(int(T) - 24) // 12)) for T in Ts))

orig_sync = torch.cuda.Event.synchronize
marks = {}
def patched(self):
    marks.setdefault("enq", time.perf_counter())          # first blocking wait = everything is queued
    return orig_sync(self)
torch.cuda.Event.synchronize = patched

for cf in (1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 30):
    svb.extract_dvectors(net, specs, chunk_frames=cf)
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        marks.clear()
        t0 = time.perf_counter()
        out = svb.extract_dvectors(net, specs, chunk_frames=cf)
        t1 = time.perf_counter()
        cur = (t1 - t0, marks.get("enq", t1) - t0)
        best = cur if best is None or cur[0] < best[0] else best
    print(f"chunk_frames {cf:>10d}: wall {best[0]*1e3:6.1f} ms (host enqueue {best[1]*1e3:6.1f} ms) -> {nwin / best[0] / 1e3:6.0f} k windows/s")

# pieces
t0 = time.perf_counter()
for _ in range(3):
    buf = np.empty((40, int(Ts.sum())), dtype=np.float32)
    np.concatenate(specs, axis=1, out=buf)
print(f"np.concatenate of all specs (pageable): {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
t0 = time.perf_counter()
for _ in range(3):
    D._chunk_plan(Ts.astype(np.int64))
print(f"index math for all utterances: {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
x = torch.tensor(I.logmel(nwin, 24, seed=1)).cuda()
with torch.no_grad():
    net(x); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net(x); e1.record(); torch.cuda.synchronize()
print(f"LSTM forward of {nwin} windows in one launch, device resident: {e0.elapsed_time(e1):.1f} ms")
