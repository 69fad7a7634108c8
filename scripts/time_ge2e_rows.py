"""Dev: device time (CUDA-graph replay) of the row-sharded GE2E (svb_ge2e_rows: 64 x 10 rows against 512 centroids, the
per-rank work of the 8-GPU step) and of get_cossim at the C5 shape (1024 x 3 rows against 1024 centroids)."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "scripts")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
from time_ge2e import graph_timeit
Eg = torch.tensor(I.ge2e_embeddings(512, 10, 256, "unit")).cuda()
Cc = svb.get_centroids(Eg)
w = torch.tensor(10.0, device="cuda"); b = torch.tensor(-5.0, device="cuda")
El = Eg[:64].contiguous()
print(f"ge2e_rows 640 x 512: {graph_timeit(lambda: torch.ops.svb200.ge2e_rows(El, Cc, w, b, 0)):.1f} us")
enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
enr, ver = torch.tensor(enr).cuda(), torch.tensor(ver).cuda()
C = svb.get_centroids(enr)
with torch.no_grad():
    print(f"get_cossim 3072 x 1024: {graph_timeit(lambda: svb.get_cossim(ver, C)):.1f} us")
