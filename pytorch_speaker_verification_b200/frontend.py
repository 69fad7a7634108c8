"""Log-mel front end on the GPU: the step before the hot path.

Host mirror of data_preprocess.py:41-45 / dvector_create.py:43-47 (librosa.core.stft -> |.|^2 -> librosa.filters.mel
-> log10(. + 1e-6)) with the reference's config defaults (sr 16 kHz, n_fft 512, 25 ms window, 10 ms hop, 40 mels).
The tables (periodic Hann window padded to n_fft, DFT twiddles, Slaney mel filterbank) are built once per device in
float64 and handed to svb_logmel; the arithmetic runs in libsvb200.so (csrc/frontend.cu), no CPU fallback.
Parity is unpinned: librosa is not available where this was built, see oracle/frontend.py.
"""
import math

import numpy as np
import torch

from . import ops

N_FFT = 512
_tables = {}


def _mel_filterbank(sr, n_fft, n_mels):
    """librosa.filters.mel defaults: Slaney scale (htk=False), fmin 0, fmax sr/2, Slaney area normalisation."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    mels = np.linspace(to_mel(0.0), to_mel(sr / 2.0), n_mels + 2)
    edges = np.where(mels >= min_log_mel, min_log_hz * np.exp(logstep * (mels - min_log_mel)), f_sp * mels)
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    lower = (freqs[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    upper = (edges[2:, None] - freqs[None, :]) / (edges[2:] - edges[1:-1])[:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    return w * (2.0 / (edges[2:] - edges[:-2]))[:, None]


def _get_tables(dev, sr, win_length, n_mels):
    key = (str(dev), sr, win_length, n_mels)
    t = _tables.get(key)
    if t is None:
        win = np.zeros(N_FFT)
        w0 = (N_FFT - win_length) // 2
        win[w0:w0 + win_length] = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win_length) / win_length)
        ang = 2.0 * np.pi * np.arange(N_FFT) / N_FFT
        tw = np.stack([np.cos(ang), np.sin(ang)], axis=1)
        t = (torch.tensor(win, dtype=torch.float32, device=dev), w0, w0 + win_length,
             torch.tensor(tw, dtype=torch.float32, device=dev).contiguous(),
             torch.tensor(_mel_filterbank(sr, N_FFT, n_mels), dtype=torch.float32, device=dev).contiguous())
        _tables[key] = t
    return t


@torch.no_grad()
def log_mel_spectrogram(y, sr=16000, win_length=400, hop=160, n_mels=40):
    """y: 1-D PCM (numpy or tensor, any device) -> (n_mels, 1 + len(y)//hop) float32 CUDA tensor, the `S` of
    data_preprocess.py:45 / dvector_create.py:47 (n_fft = 512)."""
    if win_length > N_FFT:
        raise ValueError("win_length must be <= n_fft = 512")
    y = torch.as_tensor(y)
    dev = ops._target_device(y)
    with torch.cuda.device(dev):
        yg = ops._stage(y, torch.float32, dev).reshape(-1)
        n = int(yg.numel())
        if n <= N_FFT // 2:
            raise ValueError("reflect padding needs more than n_fft/2 samples (librosa raises here too)")
        win, w0, w1, tw, melw = _get_tables(dev, sr, win_length, n_mels)
        out = torch.ops.svb200.logmel(yg, win, tw, melw, int(hop), w0, w1)
    return out
