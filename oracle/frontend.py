"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).  PARITY UNPINNED.

CPU restatement (numpy, float64) of the log-mel front end that precedes the hot path: data_preprocess.py:41-45 and
dvector_create.py:43-47 of the reference,

    S = librosa.core.stft(y, n_fft=512, win_length=400, hop_length=160); S = np.abs(S) ** 2
    S = np.log10(np.dot(librosa.filters.mel(sr=16000, n_fft=512, n_mels=40), S) + 1e-6)

The arithmetic lives in librosa (no version pinned by the reference; README.md:14 era = librosa 0.6), which is NOT
installed in this container, so this restates librosa's documented algorithm and cannot be checked against it here:
  stft: center=True with reflect padding of n_fft//2 samples, periodic Hann window of win_length samples zero-padded
        symmetrically to n_fft, frames at multiples of hop_length, 1 + len(y)//hop frames, rfft (257 bins);
  filters.mel: Slaney mel scale (htk=False: linear below 1 kHz, 27 log-spaced steps per factor 6.4 above), fmin 0,
        fmax sr/2, n_mels + 2 band edges equally spaced in mel, triangular weights, Slaney area normalisation
        2 / (f[i+2] - f[i]).
"""
import numpy as np


def hann_window_padded(win_length=400, n_fft=512):
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)        # periodic (fftbins=True)
    out = np.zeros(n_fft)
    lpad = (n_fft - win_length) // 2
    out[lpad:lpad + win_length] = w
    return out


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr=16000, n_fft=512, n_mels=40):
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    edges = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(edges)
    ramps = edges[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[:, None]
    return w


def stft_power(y, n_fft=512, win_length=400, hop=160):
    y = np.asarray(y, dtype=np.float64)
    ypad = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + len(y) // hop
    win = hann_window_padded(win_length, n_fft)
    frames = np.stack([ypad[t * hop:t * hop + n_fft] * win for t in range(n_frames)], axis=1)     # (n_fft, T)
    return np.abs(np.fft.rfft(frames, axis=0)) ** 2                                                 # (257, T)


def log_mel(y, sr=16000, n_fft=512, win_length=400, hop=160, n_mels=40):
    """(n_mels, 1 + len(y)//hop) float64 log10 mel power spectrogram."""
    return np.log10(mel_filterbank(sr, n_fft, n_mels) @ stft_power(y, n_fft, win_length, hop) + 1e-6)
