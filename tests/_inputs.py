"""Seeded synthetic inputs shared by the golden generator, the tests, smoke() and bench.py.

numpy's legacy RandomState (Mersenne Twister) is used because its stream is frozen by
numpy's compatibility policy, so the GPU box regenerates bit-identical inputs.
Shapes/distributions follow SURVEY.md section 8(d).
"""
import numpy as np

# name -> (N, M, D, kind, w, b)
GE2E_CASES = {
    "c1": (4, 5, 256, "unit", 10.0, -5.0),            # config.yaml train N,M; w,b init
    "nonunit": (7, 3, 16, "raw", 3.0, -1.0),          # E not unit-norm, odd sizes
    "m2": (5, 2, 64, "unit", 10.0, -5.0),             # smallest legal M (M=1 divides by zero)
    "clustered": (16, 6, 256, "clustered", 10.0, -5.0),  # trained-like: tight speaker clusters
    "c2": (64, 10, 256, "unit", 10.0, -5.0),          # BASELINE config 2
}

# name -> (N, M, sigma, alpha, seed);  M is hp.test.M (enrollment M/2 + verification M/2).
# Speaker centres share a common component (c_k = normalize(g + alpha*r_k)) so that
# different-speaker cosines reach the 0.50-0.99 sweep and FAR is not trivially zero.
EER_CASES = {
    "n4": (4, 6, 0.06, 0.5, 777),          # config.yaml test N,M
    "n4_low": (4, 6, 1.0, 0.9, 778),       # similarities all below 0.5 -> EER stays 0 (quirk 8)
    "n4_sep": (4, 6, 0.03, None, 782),     # well separated: FAR=FRR=0 at the first threshold
    "n64": (64, 6, 0.05, 0.5, 779),
    "n64_wide": (64, 6, 0.06, 0.3, 780),
    "n256": (256, 6, 0.06, 0.5, 781),
}


def logmel(B, T, seed, nmels=40):
    """clamp(-3 + 1.5*randn, min=-6): log10(mel+1e-6)-like features, (B,T,nmels) float32."""
    r = np.random.RandomState(seed)
    x = -3.0 + 1.5 * r.standard_normal((B, T, nmels))
    return np.maximum(x, -6.0).astype(np.float32)


def unit_rows(n, d, seed):
    r = np.random.RandomState(seed)
    x = r.standard_normal((n, d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def ge2e_embeddings(N, M, D, kind, seed=2024):
    r = np.random.RandomState(seed + N * 1000 + M)
    if kind == "unit":
        x = r.standard_normal((N, M, D))
        x /= np.linalg.norm(x, axis=2, keepdims=True)
    elif kind == "raw":
        x = r.standard_normal((N, M, D)) * 2.5 + 0.3
    elif kind == "clustered":
        c = r.standard_normal((N, 1, D))
        c /= np.linalg.norm(c, axis=2, keepdims=True)
        x = c + 0.05 * r.standard_normal((N, M, D))
        x /= np.linalg.norm(x, axis=2, keepdims=True)
    else:
        raise ValueError(kind)
    return x.astype(np.float32)


def eer_embeddings(N, M, sigma, alpha, seed, D=256):
    """normalize(c_k + sigma*randn): (enrollment, verification), each (N, M/2, D) float32."""
    r = np.random.RandomState(seed)
    c = r.standard_normal((N, 1, D))
    c /= np.linalg.norm(c, axis=2, keepdims=True)
    if alpha is not None:
        g = r.standard_normal((1, 1, D))
        g /= np.linalg.norm(g)
        c = g + alpha * c
        c /= np.linalg.norm(c, axis=2, keepdims=True)
    x = c + sigma * r.standard_normal((N, M, D))
    x /= np.linalg.norm(x, axis=2, keepdims=True)
    x = x.astype(np.float32)
    return np.ascontiguousarray(x[:, :M // 2]), np.ascontiguousarray(x[:, M // 2:])


def power_spec(T, seed, nmels=40):
    """Positive mel power spectrogram (nmels, T) float32; log10(p+1e-6) is the log-mel."""
    r = np.random.RandomState(seed)
    return np.exp(-7.0 + 3.0 * r.standard_normal((nmels, T))).astype(np.float32)


def saturating_weights(sd_numpy, seed=31337):
    """"Trained-like" parameters: LSTM weights x2.5 and N(0,0.5) biases push the gates into
    saturation (an untrained xavier net keeps them near 0.5).  sd_numpy: name -> ndarray."""
    r = np.random.RandomState(seed)
    out = {}
    for k, v in sd_numpy.items():
        if k.startswith("LSTM_stack.weight"):
            out[k] = (v * 2.5).astype(np.float32)
        elif k.startswith("LSTM_stack.bias"):
            out[k] = (0.5 * r.standard_normal(v.shape)).astype(np.float32)
        else:
            out[k] = v.astype(np.float32)
    return out
