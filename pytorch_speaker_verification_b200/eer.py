"""TI-SV equal-error-rate sweep: host mirror of train_speech_embedder.py:127-149."""
import torch

from . import ops
from .utils import get_centroids, get_cossim

THRESHOLDS = [0.01 * i + 0.5 for i in range(50)]          # train_speech_embedder.py:134
_thr_cache = {}


def _thresholds_f32(device, thresholds):
    key = (str(device), tuple(thresholds))
    if key not in _thr_cache:
        t = torch.tensor(list(thresholds), dtype=torch.float64).to(torch.float32)   # torch rounds the double to fp32
        if not bool((t[1:] >= t[:-1]).all()):
            raise ValueError("thresholds must be ascending")
        _thr_cache[key] = t.to(device)
    return _thr_cache[key]


def eer_sweep(sim_matrix, thresholds=None):
    """sim_matrix (N, M/2, N) -> (EER, EER_thresh, EER_FAR, EER_FRR) exactly as the loop at
    train_speech_embedder.py:132-149 leaves them: float32 0-dim tensors on sim_matrix's device and the Python
    double threshold, or the integers 0 when no threshold ever satisfied ``diff > |FAR-FRR|``."""
    thresholds = THRESHOLDS if thresholds is None else list(thresholds)
    out_device = sim_matrix.device
    dev = ops._target_device(sim_matrix)
    with torch.cuda.device(dev):
        sim = ops._stage(sim_matrix.detach(), torch.float32, dev)
        thr = _thresholds_f32(sim.device, thresholds)
        if sim.shape[0] == sim.shape[2] and sim.shape[0] >= 2 and len(thresholds) < 64 and sim.shape[1] * sim.shape[2] < (1 << 23):
            out_d, ca, cd = ops.eer_sweep_fused(sim, thr)
            out = out_d.cpu()
            if int(out[1]) == -2:                       # totals beyond 2^24: sequential float32 emulation
                out = ops.eer_finish(ca, cd, sim.shape[1]).cpu()
        else:
            ca, cd = ops.eer_counts(sim, thr)
            out = ops.eer_finish(ca, cd, sim.shape[1]).cpu()
    sel = int(out[1])
    if sel < 0:
        return 0, 0, 0, 0
    return (out[0].clone().to(out_device), thresholds[sel], out[2].clone().to(out_device),
            out[3].clone().to(out_device))


def compute_eer(enrollment_embeddings, verification_embeddings):
    """train_speech_embedder.py:127-149 from the two (N, M/2, D) embedding blocks."""
    enrollment_centroids = get_centroids(enrollment_embeddings)
    sim_matrix = get_cossim(verification_embeddings, enrollment_centroids)
    return eer_sweep(sim_matrix), sim_matrix
