#!/usr/bin/env python3
"""Generate tests/golden/centroids.npz from the UNMODIFIED reference (build container only):
utils.get_centroids (utils.py:27-29) and utils.get_utterance_centroids (utils.py:40-58) on seeded inputs, for the
bit-level test of svb_centroids / svb_utterance_centroids.  Small cases are stored whole; large ones as a checksum of
the float32 bit patterns (sum of the uint32 words mod 2^64, and their XOR) plus 256 sampled values.

    python tests/golden/make_golden_centroids.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

# (N, M, D, seed): training / EER batch shapes of the reference plus long-M cases that leave torch's sequential regime
CASES = [(4, 5, 256, 1), (3, 2, 3, 2), (7, 3, 16, 3), (64, 10, 256, 4), (1024, 3, 256, 5), (5, 40, 64, 6), (3, 300, 32, 7),
         (2, 4500, 16, 8), (4, 5, 16, 9), (6, 10, 40, 10), (3, 21, 33, 11), (2, 70, 100, 12), (2, 5000, 64, 13)]


def case_input(N, M, D, seed):
    r = np.random.RandomState(seed)
    return (r.standard_normal((N, M, D)) * 1.7 + 0.3 * r.standard_normal((N, 1, D))).astype(np.float32)


def bits_checksum(a):
    w = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64).ravel()
    return np.array([w.sum(dtype=np.uint64), np.bitwise_xor.reduce(w)], dtype=np.uint64)


def main():
    import torch
    ref_utils, _ = import_reference()
    out = {}
    for (N, M, D, seed) in CASES:
        E = torch.tensor(case_input(N, M, D, seed))
        C = ref_utils.get_centroids(E).numpy()
        U = ref_utils.get_utterance_centroids(E).numpy()
        tag = f"{N}x{M}x{D}"
        out[f"{tag}.C_sum"] = bits_checksum(C)
        out[f"{tag}.U_sum"] = bits_checksum(U)
        idx = np.random.RandomState(seed + 100).randint(0, U.size, size=256)
        out[f"{tag}.U_idx"] = idx
        out[f"{tag}.U_val"] = U.ravel()[idx]
        if U.size <= 8192:
            out[f"{tag}.C"] = C
            out[f"{tag}.U"] = U
        print(tag, "ok")
    np.savez_compressed(os.path.join(HERE, "centroids.npz"), **out)


if __name__ == "__main__":
    main()
