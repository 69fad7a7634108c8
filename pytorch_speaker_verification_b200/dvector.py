"""d-vector extraction: host mirror of dvector_create.py:48-52, 55-73 and 98-101."""
import numpy as np
import torch

from . import ops

WIN = 24      # dvector_create.py:49-50
HOP = 12      # int(.12 / hp.data.hop), dvector_create.py:48


def window_starts(T, win=WIN, hop=HOP):
    """Frames j = 0, hop, 2 hop, ... while j + win < T (strict; dvector_create.py:48-52)."""
    if T <= win:
        return np.zeros(0, dtype=np.int32)
    n = -(-(T - win) // hop)
    return (np.arange(n, dtype=np.int32) * hop).astype(np.int32)


def partition_offsets(W):
    """Segment boundaries of align_embeddings (dvector_create.py:56-68) as offsets [0, ..., W]."""
    if W == 0:
        return np.zeros(1, dtype=np.int32)
    offs = [0]
    j = 1
    end = 0
    for i in range(W):
        if (i * .12) + .24 < j * .401:
            end += 1
        else:
            offs.append(end)
            end += 1
            j += 1
    offs.append(end)
    return np.asarray(offs, dtype=np.int32)


def get_windows(S):
    """S (nmels, T) log-mel -> (W, 24, nmels) float32 tensor on S's device (CPU input is staged through the GPU)."""
    S = torch.as_tensor(S)
    out_device = S.device
    with torch.cuda.device(ops._dev()):
        Sg = ops._stage(S, torch.float32)
        starts = torch.from_numpy(window_starts(int(S.shape[1]))).to(Sg.device)
        out = ops.dvector_windows(Sg, starts, WIN)
    return out if out_device.type == "cuda" else out.to(out_device)


def align_embeddings(embeddings):
    """(W, D) window embeddings -> (P, D) float64 numpy array of partition means (dvector_create.py:55-73)."""
    emb = torch.as_tensor(embeddings)
    with torch.cuda.device(ops._dev()):
        eg = ops._stage(emb.detach(), torch.float32)
        offs = torch.from_numpy(partition_offsets(int(eg.shape[0]))).to(eg.device)
        out = ops.segment_mean(eg, offs)
    return out.cpu().numpy()


_part_cache = {}


def _cached_partition_offsets(W):
    po = _part_cache.get(W)
    if po is None:
        po = partition_offsets(W)
        _part_cache[W] = po
    return po


@torch.no_grad()
def extract_dvectors(embedder_net, specs, max_windows=65536):
    """Batched extraction for many utterances: specs = list of (nmels, T_u) log-mel arrays.
    Returns a list of (P_u, D) float64 arrays (empty (0, D) for utterances with no window).  One window-gather
    launch and one segment-mean launch for the whole batch; the LSTM runs on chunks of <= max_windows windows."""
    dev = ops._dev()
    with torch.cuda.device(dev):
        Ts = np.asarray([int(s.shape[1]) for s in specs], dtype=np.int64)
        cat = torch.from_numpy(np.concatenate([np.asarray(s, dtype=np.float32) for s in specs], axis=1))
        cat = cat.pin_memory().to(dev, non_blocking=True)
        nw = np.where(Ts > WIN, -(-(Ts - WIN) // HOP), 0)                 # windows per utterance (strict j+24 < T)
        base = np.concatenate([[0], np.cumsum(Ts)[:-1]])                  # first frame of each utterance
        W = int(nw.sum())
        D = embedder_net.projection.out_features
        if W == 0:
            return [np.zeros((0, D)) for _ in specs]
        wfirst = np.concatenate([[0], np.cumsum(nw)[:-1]])               # first window index of each utterance
        within = np.arange(W, dtype=np.int64) - np.repeat(wfirst, nw)
        starts = (np.repeat(base, nw) + HOP * within).astype(np.int32)
        seg, counts = [np.zeros(1, dtype=np.int64)], []
        for n, w0 in zip(nw.tolist(), wfirst.tolist()):
            if n:
                po = _cached_partition_offsets(n)
                seg.append(po[1:].astype(np.int64) + w0)
                counts.append(len(po) - 1)
            else:
                counts.append(0)
        seg = np.concatenate(seg).astype(np.int32)
        frames = ops.dvector_windows(cat, torch.from_numpy(starts).to(dev), WIN)
        emb = torch.cat([embedder_net(frames[i:i + max_windows]) for i in range(0, W, max_windows)], dim=0)
        out = ops.segment_mean(emb, torch.from_numpy(seg).to(dev)).cpu().numpy()
    res, o = [], 0
    for c in counts:
        res.append(out[o:o + c])
        o += c
    return res
