#!/usr/bin/env python3
"""Runs ONE UNMODIFIED reference script as ``__main__`` from a staged copy of the reference tree (test infrastructure).

    python tests/_run_caller.py --ref baseline/_ref --impl reference|dropin --script train_speech_embedder|dvector_create \
                                --work WORKDIR [--seed S]

* ``--impl reference``: sys.path = [ref] -> the reference's own speech_embedder_net / utils (torch CPU or stock CUDA).
* ``--impl dropin``: sys.path = [pytorch_speaker_verification_b200/dropin, ref] -> the reference's script binds our
  SpeechEmbedder / GE2ELoss / get_centroids / get_cossim; nothing else changes.
The script is executed with ``runpy.run_path(..., run_name="__main__")`` from CWD = WORKDIR, which must hold the
``config/config.yaml`` the reference reads (hparam.py:49) and the data directories that config names.

Environment shims only, no edits to reference files:
  1. PyYAML >= 6: ``yaml.load_all(stream)`` (hparam.py:9) needs a Loader -> default FullLoader;
  2. ``librosa`` is not installed: a stub module providing the two calls dvector_create.py:43-46 makes
     (``core.stft``, ``filters.mel``), computed by oracle/frontend.py's numpy restatement -- the audio front end is
     outside the hot path and is the same stub for both impls;
  3. ``webrtcvad`` is not installed, so ``VAD_segments`` (out of scope) is a stub whose
     ``VAD_chunk(aggressiveness, path)`` returns 0.4 s voiced segments of the float32 samples stored in ``path``
     (.npy bytes under a .wav name), with one gap so that dvector_create.concat_segs merges two runs.
"""
import argparse
import os
import random
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def install_shims():
    import numpy as np
    import yaml
    _orig = yaml.load_all
    yaml.load_all = lambda stream, Loader=None: _orig(stream, Loader=Loader or yaml.FullLoader)

    if ROOT not in sys.path:
        sys.path.append(ROOT)
    from oracle import frontend as ofe

    librosa = types.ModuleType("librosa")
    librosa.core = types.ModuleType("librosa.core")
    librosa.filters = types.ModuleType("librosa.filters")

    def stft(y, n_fft=512, win_length=400, hop_length=160):
        y = np.asarray(y, dtype=np.float64)
        ypad = np.pad(y, n_fft // 2, mode="reflect")
        n_frames = 1 + len(y) // hop_length
        win = ofe.hann_window_padded(win_length, n_fft)
        frames = np.stack([ypad[t * hop_length:t * hop_length + n_fft] * win for t in range(n_frames)], axis=1)
        return np.fft.rfft(frames, axis=0).astype(np.complex64)

    librosa.core.stft = stft
    librosa.stft = stft
    librosa.filters.mel = lambda sr, n_fft=512, n_mels=40: ofe.mel_filterbank(sr, n_fft, n_mels).astype(np.float32)
    sys.modules["librosa"] = librosa
    sys.modules["librosa.core"] = librosa.core
    sys.modules["librosa.filters"] = librosa.filters

    vad = types.ModuleType("VAD_segments")

    def VAD_chunk(aggressiveness, path):
        y = np.load(path, allow_pickle=False).astype(np.float32)
        seg = int(0.4 * 16000)
        n = len(y) // seg
        starts = [0.4 * i + (1.0 if i >= n // 2 else 0.0) for i in range(n)]      # a 1 s gap in the middle: two voiced runs
        segs = [y[i * seg:(i + 1) * seg] for i in range(n)]
        times = [(starts[i], starts[i + 1] if (i + 1 < n and i + 1 != n // 2) else starts[i] + 0.4) for i in range(n)]
        return times, segs

    vad.VAD_chunk = VAD_chunk
    sys.modules["VAD_segments"] = vad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--impl", required=True, choices=["reference", "dropin"])
    ap.add_argument("--script", required=True, choices=["train_speech_embedder", "dvector_create"])
    ap.add_argument("--work", required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    ref = os.path.abspath(a.ref)
    install_shims()
    os.chdir(a.work)
    sys.path.insert(0, ref)
    if a.impl == "dropin":
        sys.path.insert(0, os.path.join(ROOT, "pytorch_speaker_verification_b200", "dropin"))
    import numpy as np
    import torch
    if a.threads:
        torch.set_num_threads(a.threads)
    random.seed(a.seed)
    np.random.seed(a.seed)
    torch.manual_seed(a.seed)
    runpy.run_path(os.path.join(ref, a.script + ".py"), run_name="__main__")
    import speech_embedder_net
    mod = speech_embedder_net.SpeechEmbedder.__module__
    want = "pytorch_speaker_verification_b200" if a.impl == "dropin" else "speech_embedder_net"
    assert mod.startswith(want), (a.impl, mod)
    print(f"CALLER_OK impl={a.impl} SpeechEmbedder={mod}")


if __name__ == "__main__":
    main()
