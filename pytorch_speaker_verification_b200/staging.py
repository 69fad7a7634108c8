"""Batch staging for the training loop (train_speech_embedder.py:44-47).

The reference moves every batch with a blocking ``mel_db_batch.to(device)`` (:45) from pageable memory, reshapes it
to (N*M, T, nmels) (:47) and permutes the rows (:48-52) only to un-permute the embeddings again (:57): the row order
does not matter to a batch-independent LSTM, so the permutation is dropped here.  ``prefetch`` wraps any iterable of
host batches (the reference's DataLoader), copies batch i+1 through page-locked staging buffers on a copy stream
while batch i trains, and yields device tensors already reshaped.  torch streams/events only: no kernels.
"""
import torch

from . import ops


def prefetch(batches, device=None, depth=2, flatten=True):
    """Yield CUDA float32 tensors for the host tensors in ``batches``; (N, M, T, F) batches become (N*M, T, F) when
    ``flatten``.  ``depth`` page-locked buffers are cycled; the H2D copy of the next batch overlaps the caller's
    work on the current one."""
    dev = torch.device(device) if device is not None else ops._dev()
    if dev.type != "cuda":
        raise ops._lib.SvbError("prefetch needs a CUDA device (no CPU fallback)")
    copy_stream = torch.cuda.Stream(dev)
    pinned = [None] * depth
    copied = [None] * depth          # event: the H2D copy out of pinned[i] has finished (the buffer may be refilled)
    queue = []

    def issue(i, b):
        b = torch.as_tensor(b)
        if b.dtype != torch.float32:
            b = b.float()                                              # speech_embedder_net.py:28
        slot = i % depth
        if b.is_pinned() and b.is_contiguous():
            host = b                                                   # already page-locked (DataLoader pin_memory)
        else:
            if copied[slot] is not None:
                copied[slot].synchronize()
            if pinned[slot] is None or pinned[slot].numel() < b.numel():
                pinned[slot] = torch.empty(b.numel(), dtype=torch.float32).pin_memory()
            host = pinned[slot][:b.numel()].view(b.shape)
            host.copy_(b)
        with torch.cuda.stream(copy_stream):
            d = host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        copied[slot] = ev
        if flatten and d.dim() == 4:
            d = d.reshape(d.shape[0] * d.shape[1], d.shape[2], d.shape[3])   # train_speech_embedder.py:47
        queue.append((d, ev))

    def pop():
        d, ev = queue.pop(0)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        d.record_stream(cur)
        return d

    i = 0
    for b in batches:
        issue(i, b)
        i += 1
        if len(queue) > 1:
            yield pop()
    while queue:
        yield pop()
