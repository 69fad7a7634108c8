"""Oracle (test infrastructure): GE2E loss math restated in numpy.

Follows /root/reference/utils.py and speech_embedder_net.py:
  get_centroids            utils.py:27-29
  get_utterance_centroids  utils.py:40-58
  get_cossim               utils.py:72-115   (F.cosine_similarity eps=1e-8, each norm
                                              clamped on its own; +1e-6 at :114;
                                              diagonal overwrite at :113)
  calc_loss                utils.py:126-132  (sum, not mean; +1e-6 inside the log)
  GE2ELoss.forward         speech_embedder_net.py:43-49 (clamp at :44 is a no-op)

All functions are dtype-generic: pass float32 arrays for the fp32 oracle, float64
for the high-precision oracle used for ill-conditioned outputs (db).
``ge2e_fwd_bwd`` additionally restates the closed-form gradient that the CUDA
kernels implement; tests check it against reference autograd (tests/golden).
"""
import numpy as np

COS_EPS = 1e-8     # F.cosine_similarity default eps (utils.py:91,105)
COS_BIAS = 1e-6    # utils.py:114
LOG_BIAS = 1e-6    # utils.py:129


def _cascade(rows):
    """Cascade summation of a list of equally shaped float32 arrays (ATen SumKernel.cpp multi_row_sum): sequential
    into a level-0 accumulator that is flushed into level 1 every 16 rows, level 1 into level 2 every 256, level 2
    into level 3 every 4096; the remaining rows go to level 0 and the levels are added in order."""
    M = len(rows)
    z = np.zeros_like(rows[0]) if M else None
    acc = [z.copy() for _ in range(4)]
    i = 0
    while i + 16 <= M:
        for _ in range(16):
            acc[0] = acc[0] + rows[i]
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = z.copy()
            if (i & (15 << (4 * j))) != 0:
                break
    while i < M:
        acc[0] = acc[0] + rows[i]
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    return acc[0]


def sum_utterances_torch_order(E, vec_chunk=32):
    """``E.sum(dim=1)`` of a contiguous (N, M, D) float32 tensor in the order torch's CPU kernel adds the rows (ATen
    SumKernel.cpp, vectorized_outer_sum with 256-bit vectors -- the kernel torch selects in this image):
      * columns [0, 32 * (D // 32)): cascade summation over the rows (plain sequential for M < 16);
      * the remaining columns: ``row_sum`` -- four interleaved partial sums over rows i = k mod 4 (each a cascade over
        M // 4 rows), the M % 4 tail rows added to partial 0, then p0 + p1 + p2 + p3.
    Pinned against the reference by tests/golden/centroids.npz."""
    N, M, D = E.shape
    out = np.empty((N, D), dtype=E.dtype)
    dv = vec_chunk * (D // vec_chunk)
    if dv:
        out[:, :dv] = _cascade([E[:, i, :dv] for i in range(M)])
    if dv < D:
        q = M // 4
        part = [_cascade([E[:, 4 * i + k, dv:] for i in range(q)]) if q else np.zeros((N, D - dv), dtype=E.dtype)
                for k in range(4)]
        for i in range(4 * q, M):
            part[0] = part[0] + E[:, i, dv:]
        out[:, dv:] = ((part[0] + part[1]) + part[2]) + part[3]
    return out


def get_centroids(E):
    """utils.py:27-29 -- mean over the utterance axis."""
    return E.mean(axis=1)


def get_centroids_bitwise(E):
    """utils.py:27-29 with torch's float32 operation order: cascade sum, then one division by M."""
    return (sum_utterances_torch_order(E) / E.dtype.type(E.shape[1])).astype(E.dtype)


def get_utterance_centroids_bitwise(E):
    """utils.py:40-58 with torch's float32 operation order: (cascade sum - E) / (M - 1)."""
    s = sum_utterances_torch_order(E)[:, None, :]
    return ((s - E).astype(E.dtype) / E.dtype.type(E.shape[1] - 1)).astype(E.dtype)


def get_utterance_centroids(E):
    """utils.py:40-58 -- leave-one-out centroid of every utterance."""
    M = E.shape[1]
    s = E.sum(axis=1, keepdims=True)
    return (s - E) / E.dtype.type(M - 1)


def _cos(x, y):
    """F.cosine_similarity(x, y, dim=-1, eps=1e-8) as torch>=1.12 computes it."""
    eps = x.dtype.type(COS_EPS)
    nx = np.maximum(np.sqrt((x * x).sum(-1, keepdims=True)), eps)
    ny = np.maximum(np.sqrt((y * y).sum(-1, keepdims=True)), eps)
    return ((x / nx) * (y / ny)).sum(-1)


def get_cossim(E, C):
    """utils.py:72-115.  E (N,M,D), C (N',D) with N' == N -> (N,M,N').

    The diagonal [j,:,j] is the cosine to the leave-one-out centroid of E's own
    utterances even when C is foreign (EER use, train_speech_embedder.py:129)."""
    N, M, D = E.shape
    U = get_utterance_centroids(E)
    cos_same = _cos(E.reshape(N * M, D), U.reshape(N * M, D)).reshape(N, M)
    cos = _cos(E[:, :, None, :], C[None, None, :, :])          # (N,M,N')
    idx = np.arange(N)
    cos[idx, :, idx] = cos_same
    return cos + E.dtype.type(COS_BIAS)


def calc_loss(S):
    """utils.py:126-132 -> (loss, per_embedding_loss (N,M))."""
    N = S.shape[0]
    idx = np.arange(N)
    pos = S[idx, :, idx]                                        # (N,M)
    neg = np.log(np.exp(S).sum(axis=2) + S.dtype.type(LOG_BIAS))
    per = -1 * (pos - neg)
    return per.sum(), per


def ge2e_loss(E, w, b):
    """speech_embedder_net.py:43-49."""
    C = get_centroids(E)
    cos = get_cossim(E, C)
    S = E.dtype.type(w) * cos + E.dtype.type(b)
    return calc_loss(S)[0]


def ge2e_fwd_bwd(E, w, b):
    """Closed-form loss and gradient of GE2ELoss.forward.

    Returns dict(loss, per, cos, dE, dw, db).  This is the restatement the CUDA
    kernels follow; it is validated against reference autograd in
    tests/test_oracle_golden.py."""
    dt = E.dtype.type
    N, M, D = E.shape
    eps = dt(COS_EPS)
    ne = np.maximum(np.sqrt((E * E).sum(-1, keepdims=True)), eps)       # (N,M,1)
    Eh = E / ne
    s = E.sum(axis=1, keepdims=True)                                    # (N,1,D)
    c = s[:, 0, :] / dt(M)
    nc = np.maximum(np.sqrt((c * c).sum(-1, keepdims=True)), eps)       # (N,1)
    Ch = c / nc
    U = (s - E) / dt(M - 1)
    nu = np.maximum(np.sqrt((U * U).sum(-1, keepdims=True)), eps)
    Uh = U / nu
    cos0 = np.einsum('jid,kd->jik', Eh, Ch)                             # cos without +1e-6
    idx = np.arange(N)
    cosd0 = (Eh * Uh).sum(-1)                                           # (N,M)
    cos0[idx, :, idx] = cosd0
    cos = cos0 + dt(COS_BIAS)
    S = dt(w) * cos + dt(b)
    ex = np.exp(S)
    den = ex.sum(axis=2) + dt(LOG_BIAS)                                 # (N,M)
    per = -(S[idx, :, idx] - np.log(den))
    loss = per.sum()
    G = ex / den[:, :, None]
    G[idx, :, idx] -= dt(1)
    dw = (G * cos).sum()
    db = G.sum()
    A = dt(w) * G
    a = A[idx, :, idx].copy()                                           # (N,M)
    Aoff = A.copy()
    Aoff[idx, :, idx] = 0
    cos_off = cos0.copy()
    cos_off[idx, :, idx] = 0
    r = (Aoff * cos_off).sum(axis=2)                                    # (N,M)
    R = np.einsum('jik,kd->jid', Aoff, Ch)
    dE = (R - r[:, :, None] * Eh) / ne
    dE += a[:, :, None] * (Uh - cosd0[:, :, None] * Eh) / ne
    P = np.einsum('jik,jid->kd', Aoff, Eh)
    q = (Aoff * cos_off).sum(axis=(0, 1))                               # (N,)
    dC = (P - q[:, None] * Ch) / nc
    dE += dC[:, None, :] / dt(M)
    dU = a[:, :, None] * (Eh - cosd0[:, :, None] * Uh) / nu
    dE += (dU.sum(axis=1, keepdims=True) - dU) / dt(M - 1)
    return dict(loss=loss, per=per, cos=cos, dE=dE, dw=dw, db=db)


# --- loop oracles (utils.py:16-25, 31-38, 60-70, 117-124): element-wise restatement,
# --- used as a second pin on small cases only.
def get_cossim_loops(E, C):
    N, M, D = E.shape
    out = np.zeros((N, M, C.shape[0]), dtype=E.dtype)
    for j in range(N):
        for i in range(M):
            for k in range(C.shape[0]):
                cen = C[k]
                if j == k:
                    cen = (E[j].sum(axis=0) - E[j, i]) / E.dtype.type(M - 1)
                out[j, i, k] = _cos(E[j, i][None], cen[None])[0] + E.dtype.type(COS_BIAS)
    return out


def calc_loss_loops(S):
    N, M, _ = S.shape
    per = np.zeros((N, M), dtype=S.dtype)
    for j in range(N):
        for i in range(M):
            per[j, i] = -(S[j, i, j] - np.log(np.exp(S[j, i]).sum() + S.dtype.type(LOG_BIAS)))
    return per.sum(), per
