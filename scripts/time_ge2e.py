"""Dev: device time of the GE2E and EER kernels through the raw C ABI with preallocated buffers (CUDA events around
a loop of back-to-back launches; the CPU launch path is shorter than the kernels, so this is GPU time per call)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from pytorch_speaker_verification_b200 import ops, eer as E, _lib
from pytorch_speaker_verification_b200._lib import ptr, stream_ptr
L = _lib.lib()

def timeit(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

def graph_timeit(fn, reps=20, n=10):
    """Device time per call: `reps` calls captured in a CUDA graph, replayed n times."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): fn()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps): fn()
        g.replay(); side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(n): g.replay()
        e1.record(side); side.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3

def ge2e_raw(N, M, D, need_grad, fused):
    Eg = torch.tensor(I.ge2e_embeddings(N, M, D, "unit")).cuda()
    w = torch.tensor(10.0, device="cuda"); b = torch.tensor(-5.0, device="cuda")
    nb = ctypes.c_size_t(0); L.svb_ge2e_workspace_bytes(N, M, D, N, ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    loss = torch.empty((), device="cuda"); per = torch.empty(N, M, device="cuda")
    dE = torch.empty_like(Eg); dw = torch.empty((), device="cuda"); db = torch.empty((), device="cuda")
    def call():
        st = stream_ptr()
        rc = L.svb_ge2e(ptr(Eg), None, N, M, D, N, ptr(w), ptr(b), None, None, None, ptr(per), ptr(loss),
                        ptr(dE) if need_grad else None, None, ptr(dw) if need_grad else None,
                        ptr(db) if need_grad else None, ptr(ws), ctypes.c_size_t(nb.value), int(fused), st)
        assert rc == 0
    return call

if __name__ == "__main__":
  for (N, M) in ((64, 10), (512, 10)):
      for fused in (True, False):
          print(f"GE2E fwd+bwd N={N} M={M} fused={fused}: {timeit(ge2e_raw(N, M, 256, True, fused)):.1f} us")
      print(f"GE2E fwd only N={N}: {timeit(ge2e_raw(N, M, 256, False, True)):.1f} us")
      print(f"GE2E fwd+bwd N={N} M={M}: {graph_timeit(ge2e_raw(N, M, 256, True, True)):.1f} us by graph replay (device time)")

  enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
  enr, ver = torch.tensor(enr).cuda(), torch.tensor(ver).cuda()
  sim = svb.get_cossim(ver, svb.get_centroids(enr))
  thr = E._thresholds_f32(sim.device, E.THRESHOLDS)
  N, Mv, T = 1024, 3, 50
  ca = torch.empty(N, T, dtype=torch.int32, device="cuda"); cd = torch.empty_like(ca)
  scratch = torch.zeros(1 + 16 * T, dtype=torch.int64, device="cuda"); out = torch.empty(4 + 2 * T, device="cuda")
  st = stream_ptr()
  def sweep():
      assert L.svb_eer_sweep(ptr(sim), N, Mv, ptr(thr), T, ptr(ca), ptr(cd), ptr(scratch), ptr(out), st) == 0
  us = timeit(sweep)
  print(f"EER fused sweep N=1024 Mv=3 (memset + 1 kernel): {us:.1f} us -> {sim.numel()*4/us/1e3:.0f} GB/s of {sim.numel()*4/1e6:.1f} MB")
  def counts():
      assert L.svb_eer_counts(ptr(sim), N, Mv, N, 0, ptr(thr), T, ptr(ca), ptr(cd), st) == 0
  print(f"EER counts only (old kernel): {timeit(counts):.1f} us")
  fo = torch.empty(4 + 2 * T, device="cuda")
  print(f"EER finish (sequential): {timeit(lambda: L.svb_eer_finish(ptr(ca), ptr(cd), N, Mv, T, ptr(fo), st)):.1f} us")
  print("fused == sequential:", torch.equal(out[:4].cpu(), fo[:4].cpu()), out[:4].tolist())
