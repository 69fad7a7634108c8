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
    dev = ops._target_device(S)
    with torch.cuda.device(dev):
        Sg = ops._stage(S, torch.float32, dev)
        starts = torch.from_numpy(window_starts(int(S.shape[1]))).to(Sg.device)
        out = ops.dvector_windows(Sg, starts, WIN)
    return out if out_device.type == "cuda" else out.to(out_device)


def align_embeddings(embeddings):
    """(W, D) window embeddings -> (P, D) float64 numpy array of partition means (dvector_create.py:55-73)."""
    emb = torch.as_tensor(embeddings)
    dev = ops._target_device(emb)
    with torch.cuda.device(dev):
        eg = ops._stage(emb.detach(), torch.float32, dev)
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


class _PinnedPool:
    """Reusable page-locked staging buffers (cudaHostAlloc costs milliseconds; extraction is called per file list)."""

    def __init__(self):
        self.free = []

    def take(self, nbytes):
        best = None
        for i, t in enumerate(self.free):
            if t.numel() >= nbytes and (best is None or t.numel() < self.free[best].numel()):
                best = i
        if best is not None:
            return self.free.pop(best)
        size = 1 << 20                                   # power-of-two buckets: any later chunk of similar size fits
        while size < nbytes:
            size <<= 1
        return torch.empty(size, dtype=torch.uint8).pin_memory()

    def give(self, t):
        self.free.append(t)


_pool = _PinnedPool()


def _chunk_plan(Ts):
    """Index math of one chunk of utterances with frame counts Ts: window start frames (in the concatenated chunk),
    partition offsets, partitions per utterance."""
    nw = np.where(Ts > WIN, -(-(Ts - WIN) // HOP), 0)                     # windows per utterance (strict j+24 < T)
    base = np.concatenate([[0], np.cumsum(Ts)[:-1]])                      # first frame of each utterance
    W = int(nw.sum())
    wfirst = np.concatenate([[0], np.cumsum(nw)[:-1]])                    # first window index of each utterance
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
    return W, starts, np.concatenate(seg).astype(np.int32), counts


_stage_workers = None


def _workers():
    global _stage_workers
    if _stage_workers is None:
        from concurrent.futures import ThreadPoolExecutor
        _stage_workers = ThreadPoolExecutor(max_workers=4, thread_name_prefix="svb-stage")
    return _stage_workers


def _fill_specs(dst, specs):
    """dst (nmels, F) float32 view of a pinned buffer <- the utterances' log-mels side by side.  600 MB of host
    copies for 12.5 k utterances: split over the staging workers (numpy copies release the GIL)."""
    n = len(specs)
    if n < 64:
        np.concatenate([np.asarray(s, dtype=np.float32) for s in specs], axis=1, out=dst)
        return
    offs = np.concatenate([[0], np.cumsum([int(s.shape[1]) for s in specs])])
    parts = 4
    bounds = [n * i // parts for i in range(parts + 1)]

    def job(i):
        lo, hi = bounds[i], bounds[i + 1]
        if hi > lo:
            np.concatenate([np.asarray(s, dtype=np.float32) for s in specs[lo:hi]], axis=1,
                           out=dst[:, int(offs[lo]):int(offs[hi])])

    list(_workers().map(job, range(parts)))


@torch.no_grad()
def extract_dvectors(embedder_net, specs, max_windows=65536, chunk_frames=1 << 16):
    """Batched extraction for many utterances: specs = list of (nmels, T_u) log-mel arrays.
    Returns a list of (P_u, D) float64 arrays (empty (0, D) for utterances with no window).

    The utterances are processed in chunks of about ``chunk_frames`` frames: per chunk one pinned staging copy, one
    asynchronous H2D copy, one window-gather launch, the LSTM on <= max_windows windows at a time, one segment-mean
    launch and an asynchronous D2H copy into pinned memory.  The staging of the chunks (index math + the host copies
    into page-locked memory, the largest host cost) runs on a producer thread up to three chunks ahead of the thread
    that queues the GPU work; nothing synchronises until the last chunk is queued."""
    import queue
    import threading
    dev = ops._target_device(*embedder_net.parameters())        # the module's device, else the current one
    D = embedder_net.projection.out_features
    n = len(specs)
    if n == 0:
        return []
    nmels = int(specs[0].shape[0])
    Ts_all = np.asarray([int(s.shape[1]) for s in specs], dtype=np.int64)
    # chunk boundaries by cumulative frames; the first two chunks are smaller so that the GPU starts early
    bounds, acc, target = [0], 0, max(1, chunk_frames // 4)
    for i, T in enumerate(Ts_all.tolist()):
        acc += T
        if acc >= target:
            bounds.append(i + 1)
            acc = 0
            target = min(chunk_frames, target * 2)
    if bounds[-1] != n:
        bounds.append(n)
    pending = []          # per chunk: (pinned result view, D2H event, buffers) until harvested
    counts_of = {}
    res_chunks = [None] * (len(bounds) - 1)
    pool_lock = threading.Lock()

    def take(nbytes):
        with pool_lock:
            return _pool.take(nbytes)

    def give(t):
        with pool_lock:
            _pool.give(t)

    def harvest(block):
        """Copy finished chunks out of their page-locked buffers (one memcpy per chunk) and recycle the buffers;
        non-blocking while later chunks are still being queued."""
        for k, item in enumerate(pending):
            if item is None:
                continue
            hv, ev, bufs = item
            if hv is not None:
                if not block and not ev.query():
                    continue
                if block:
                    ev.synchronize()
                res_chunks[k] = np.array(hv.numpy())
                for t in bufs:
                    give(t)
            pending[k] = None

    staged = queue.Queue(maxsize=3)

    def producer():
        try:
            for k, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
                Ts = Ts_all[lo:hi]
                F = int(Ts.sum())
                W, starts, seg, counts = _chunk_plan(Ts)
                if W == 0:
                    staged.put((k, None, counts, 0, 0, 0, 0))
                    continue
                # one staging buffer: [log-mel (nmels, F) float32 | starts int32 | seg int32]
                nb_spec, nb_st, nb_seg = nmels * F * 4, starts.size * 4, seg.size * 4
                stage = take(nb_spec + nb_st + nb_seg)
                sv = stage.numpy()
                _fill_specs(sv[:nb_spec].view(np.float32).reshape(nmels, F), specs[lo:hi])
                sv[nb_spec:nb_spec + nb_st].view(np.int32)[:] = starts
                sv[nb_spec + nb_st:nb_spec + nb_st + nb_seg].view(np.int32)[:] = seg
                staged.put((k, stage, counts, W, (nb_spec, nb_st, nb_seg), F, int(seg.size) - 1))
        except BaseException as exc:                    # surfaced on the consumer side
            staged.put(exc)

    th = threading.Thread(target=producer, name="svb-extract-stage", daemon=True)
    th.start()
    with torch.cuda.device(dev):
        for _ in range(len(bounds) - 1):
            item = staged.get()
            if isinstance(item, BaseException):
                raise item
            k, stage, counts, W, sizes, F, P = item
            res_chunks[k] = counts                       # replaced by the data in harvest() when the chunk has windows
            if stage is None:
                pending.append((None, None, ()))
                continue
            nb_spec, nb_st, nb_seg = sizes
            g = stage[:nb_spec + nb_st + nb_seg].to(dev, non_blocking=True)
            cat = g[:nb_spec].view(torch.float32).view(nmels, F)
            st_d = g[nb_spec:nb_spec + nb_st].view(torch.int32)
            seg_d = g[nb_spec + nb_st:].view(torch.int32)
            frames = ops.dvector_windows(cat, st_d, WIN)
            if W <= max_windows:
                emb = embedder_net(frames)
            else:
                emb = torch.cat([embedder_net(frames[i:i + max_windows]) for i in range(0, W, max_windows)], dim=0)
            out = ops.segment_mean(emb, seg_d)                             # (P, D) float64
            host = take(P * D * 8)
            hv = host[:P * D * 8].view(torch.float64).view(P, D)
            hv.copy_(out, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((hv, ev, (stage, host)))
            counts_of[k] = counts
            harvest(False)
        harvest(True)
    th.join()
    res = []
    for k in range(len(res_chunks)):
        counts = counts_of.get(k)
        if counts is None:                                # chunk without any window
            res.extend(np.zeros((0, D)) for _ in res_chunks[k])
            continue
        arr, o = res_chunks[k], 0
        for c in counts:
            res.append(arr[o:o + c] if c else np.zeros((0, D)))            # views of the chunk's array
            o += c
    return res
