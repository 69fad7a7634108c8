"""Dev: device time of the C5 EER sweep (N = 1024, M = 6) on a typical verification set and on an adversarial one (every
similarity above the first threshold), CUDA-graph replay of 20 launches and a plain call loop."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import _lib, eer as E
from pytorch_speaker_verification_b200._lib import ptr, stream_ptr
L = _lib.lib()
dev = torch.device("cuda")
enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
enr, ver = torch.tensor(enr).to(dev), torch.tensor(ver).to(dev)
with torch.no_grad():
    sim = svb.get_cossim(ver, svb.get_centroids(enr)).contiguous()
N, Mv, T = 1024, 3, 50
thr = E._thresholds_f32(sim.device, E.THRESHOLDS)
ca = torch.empty(N, T, dtype=torch.int32, device=dev); cd = torch.empty_like(ca)
scratch = torch.zeros(1 + 16 * T, dtype=torch.int64, device=dev)
res = torch.empty(4 + 2 * T, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, s in (("typical", sim), ("all above thr[0]", (sim * 0.2 + 0.75).contiguous())):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        st = stream_ptr()
        call = lambda: L.svb_eer_sweep(ptr(s), N, Mv, ptr(thr), T, ptr(ca), ptr(cd), ptr(scratch), ptr(res), st)
        for _ in range(3): call()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(20): call()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay(); side.synchronize()
        a.record(side)
        for _ in range(10): g.replay()
        b.record(side); side.synchronize()
        us_graph = a.elapsed_time(b) * 1e3 / 200
        # cold: L2 flushed before every launch
        tot = 0.0
        for _ in range(10):
            flush.fill_(1)
            a.record(side); call(); b.record(side); side.synchronize()
            tot += a.elapsed_time(b) * 1e3
    print(f"{name:18s}: {us_graph:6.2f} us per launch (graph replay, L2-warm: {s.numel() * 4 / us_graph / 1e3:.0f} GB/s), "
          f"{tot / 10:6.2f} us cold (L2 flushed: {s.numel() * 4 / (tot / 10) / 1e3:.0f} GB/s)   EER {float(res[0]):.6f}")
