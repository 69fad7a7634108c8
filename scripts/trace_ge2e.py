"""Dev: per-phase clock64 stamps of the one-CTA-per-speaker GE2E kernel (SVB_GE2E_TRACE=1 must be set before the
library is loaded).  Prints, per phase boundary, the mean / max over CTAs of the cycles since the CTA's first stamp."""
import ctypes, os, sys
os.environ["SVB_GE2E_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
from pytorch_speaker_verification_b200 import _lib
from pytorch_speaker_verification_b200._lib import ptr, stream_ptr
L = _lib.lib()
NAMES = ["start", "phase1 done", "barrier1 passed", "centroids in smem", "cos done", "softmax done", "R/P done",
         "barrier2 passed", "slabs summed", "dE done"]
for (N, M, D, grad) in ((64, 10, 256, True), (64, 10, 256, False), (128, 10, 256, True)):
    Eg = torch.tensor(I.ge2e_embeddings(N, M, D, "unit")).cuda()
    w = torch.tensor(10.0, device="cuda"); b = torch.tensor(-5.0, device="cuda")
    nb = ctypes.c_size_t(0); L.svb_ge2e_workspace_bytes(N, M, D, N, ctypes.byref(nb))
    off = ctypes.c_size_t(0); L.svb_ge2e_trace_offset(N, M, D, N, ctypes.byref(off))
    ws = torch.zeros(nb.value, dtype=torch.uint8, device="cuda")
    loss = torch.empty((), device="cuda"); dE = torch.empty_like(Eg)
    dw = torch.empty((), device="cuda"); db = torch.empty((), device="cuda")
    for _ in range(5):
        rc = L.svb_ge2e(ptr(Eg), None, N, M, D, N, ptr(w), ptr(b), None, None, None, None, ptr(loss),
                        ptr(dE) if grad else None, None, ptr(dw) if grad else None, ptr(db) if grad else None,
                        ptr(ws), ctypes.c_size_t(nb.value), 1, stream_ptr())
        assert rc == 0
    torch.cuda.synchronize()
    tr = ws[off.value:off.value + 148 * 16 * 8].view(torch.int64).view(148, 16).cpu().numpy()
    tr = tr[tr[:, 0] != 0]
    print(f"{len(tr)} CTAs")
    rel = tr[:, :10] - tr[:, :1]
    rel[rel < 0] = 0
    print(f"N={N} M={M} D={D} grad={grad}  (cycles since the CTA's first stamp; mean / max over CTAs)")
    for i, n in enumerate(NAMES):
        print(f"  {n:20s} {rel[:, i].mean():9.0f} {rel[:, i].max():9.0f}")
