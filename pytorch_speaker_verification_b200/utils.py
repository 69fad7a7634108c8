"""Drop-in for the loss math of the reference's utils.py (lines 27-29, 40-58, 72-115, 126-132).

Same names, argument meaning and return types; the arithmetic runs in the fused CUDA kernels of
csrc/ge2e.cu (no CPU path).  All three functions are differentiable like the reference's.
"""
import torch

from . import ops


def get_centroids(embeddings):
    """(N, M, D) -> (N, D): mean over each speaker's utterances (utils.py:27-29)."""
    return ops.CentroidsFn.apply(embeddings)


def get_cossim(embeddings, centroids):
    """(N, M, D), (N, D) -> (N, M, N) cosine similarity + 1e-6 with the diagonal computed against the
    leave-one-out centroid of ``embeddings`` (utils.py:72-115), also when ``centroids`` is foreign
    (train_speech_embedder.py:129)."""
    return ops.CossimFn.apply(embeddings, centroids)


def calc_loss(sim_matrix):
    """(N, M, N) -> (loss, per_embedding_loss (N, M)); loss is the SUM over rows (utils.py:126-132)."""
    return ops.CalcLossFn.apply(sim_matrix)


def get_utterance_centroids(embeddings):
    """(N, M, D) -> (N, M, D) leave-one-out centroids (utils.py:40-58).  Helper kept for API completeness;
    the kernels never materialise it, so this view is produced on demand from the sums."""
    c = get_centroids(embeddings)
    M = embeddings.shape[1]
    return (c.unsqueeze(1) * M - embeddings.to(c.dtype)) / (M - 1)
