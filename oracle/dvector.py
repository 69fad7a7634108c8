"""Oracle (test infrastructure): d-vector window extraction and alignment.

Follows /root/reference/dvector_create.py:
  :48-52  windows S[:, j:j+24] for j = 0,12,24,... while j+24 < T (strict)
  :98-99  np.stack(axis=2) -> transpose(2,1,0): (W, 24, 40)
  :55-73  align_embeddings: window i joins partition j while (i*.12)+.24 < j*.401 (doubles);
          np.average of float32 rows, stored into a float64 (P,256) array, no re-normalisation.
"""
import numpy as np

WIN = 24
HOP = 12   # int(.12 / hp.data.hop) with hop = 0.01 (dvector_create.py:48, config.yaml:14)


def window_starts(T, win=WIN, hop=HOP):
    out = []
    for j in range(0, T, hop):
        if j + win < T:
            out.append(j)
        else:
            break
    return out


def windows(S, win=WIN, hop=HOP):
    """S (nmels, T) log-mel -> (W, win, nmels) float32; W may be 0 (the reference would raise
    in np.stack at :98 -- undefined there, an empty array here)."""
    starts = window_starts(S.shape[1], win, hop)
    if not starts:
        return np.zeros((0, win, S.shape[0]), dtype=S.dtype)
    frames = np.stack([S[:, j:j + win] for j in starts], axis=2)      # (nmels, win, W)
    return np.transpose(frames, axes=(2, 1, 0))


def partitions(W):
    """dvector_create.py:56-68 -> list of (start, end)."""
    parts = []
    start = 0
    end = 0
    j = 1
    for i in range(W):
        if (i * .12) + .24 < j * .401:
            end = end + 1
        else:
            parts.append((start, end))
            start = end
            end = end + 1
            j += 1
    parts.append((start, end))
    return parts


def align_embeddings(emb):
    """dvector_create.py:55-73.  emb (W,256) float32 -> (P,256) float64."""
    parts = partitions(len(emb))
    out = np.zeros((len(parts), emb.shape[1]))
    for i, (s, e) in enumerate(parts):
        out[i] = np.average(emb[s:e], axis=0)
    return out
