"""Test infrastructure: synthetic work directories for the reference's scripts and a runner for tests/_run_caller.py.

The reference tree is staged (unmodified, git-ignored) under baseline/_ref/ by ``__graft_entry__.build()`` in the build
container (/root/reference does not exist on the GPU box, baseline/_ref/ travels with the
snapshot).  Everything the scripts read besides their own sources -- config/config.yaml values, the per-speaker
``.npy`` files of data_load.py:75-84, the wav folders dvector_create.py:75,89-92 walks -- is generated here.
"""
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")
REF_FILES = ["hparam.py", "utils.py", "speech_embedder_net.py", "train_speech_embedder.py", "data_load.py",
             "dvector_create.py", os.path.join("config", "config.yaml")]


def reference_staged():
    return all(os.path.isfile(os.path.join(REF, f)) for f in REF_FILES)


def speaker_logmel(n_speakers, utts, frames, seed, nmels=40):
    """(n_speakers, utts, nmels, frames) float32 log-mel-like features with a per-speaker spectral envelope (so that an
    untrained LSTM already separates speakers and the EER sweep is not degenerate): floor -6 like log10(. + 1e-6)."""
    r = np.random.RandomState(seed)
    env = -3.0 + 1.6 * r.standard_normal((n_speakers, 1, nmels, 1))
    tilt = 0.8 * r.standard_normal((n_speakers, utts, nmels, 1))
    x = env + tilt + 1.0 * r.standard_normal((n_speakers, utts, nmels, frames))
    return np.maximum(x, -6.0).astype(np.float32)


def write_config(work, **over):
    """config/config.yaml with the reference's keys (config/config.yaml:1-40); ``over`` uses dotted names."""
    import yaml
    with open(os.path.join(REF, "config", "config.yaml")) as f:
        cfg = {}
        for doc in yaml.load_all(f, Loader=yaml.FullLoader):
            cfg.update(doc)
    for k, v in over.items():
        d = cfg
        parts = k.split("__")
        for p in parts[:-1]:
            d = d[p]
        assert parts[-1] in d, k                    # never add keys the reference does not have
        d[parts[-1]] = v
    os.makedirs(os.path.join(work, "config"), exist_ok=True)
    with open(os.path.join(work, "config", "config.yaml"), "w") as f:
        yaml.safe_dump(cfg, f)
    return cfg


def make_tisv_dirs(work, n_train=8, n_test=8, utts=12, frames=180, seed=5):
    """train_tisv/ and test_tisv/ as data_preprocess.py:46-54 writes them: speaker{i}.npy of shape (U, 40, frames)."""
    tr = speaker_logmel(n_train, utts, frames, seed)
    te = speaker_logmel(n_test, utts, frames, seed + 1)
    for name, arr in (("train_tisv", tr), ("test_tisv", te)):
        d = os.path.join(work, name)
        os.makedirs(d, exist_ok=True)
        for i in range(arr.shape[0]):
            np.save(os.path.join(d, f"speaker{i}.npy"), arr[i])


def make_wav_dirs(work, n_speakers=12, files=2, seconds=3.2, seed=9):
    """TIMIT/TRAIN/DR1/SPK{i}/utt{j}.wav for dvector_create.py:75,89-92 (hp.unprocessed_data './TIMIT/*/*/*/*.wav');
    each "wav" holds float32 samples in .npy format, read by the VAD stub of tests/_run_caller.py."""
    r = np.random.RandomState(seed)
    t = np.arange(int(seconds * 16000)) / 16000.0
    for s in range(n_speakers):
        d = os.path.join(work, "TIMIT", "TRAIN", "DR1", f"SPK{s:02d}")
        os.makedirs(d, exist_ok=True)
        f0 = 90.0 + 12.0 * s
        for j in range(files):
            y = sum((0.3 / k) * np.sin(2 * np.pi * f0 * k * t + r.uniform(0, 6.28)) for k in range(1, 12))
            y = y * (0.3 + np.abs(np.sin(2 * np.pi * (1.0 + 0.2 * j) * t))) + 0.02 * r.standard_normal(len(t))
            with open(os.path.join(d, f"utt{j}.wav"), "wb") as f:
                np.save(f, y.astype(np.float32))


def run_caller(impl, script, work, seed=0, timeout=1500, threads=0):
    """-> stdout of the script; raises with the tail of the output when it fails."""
    cmd = [sys.executable, os.path.join(HERE, "_run_caller.py"), "--ref", REF, "--impl", impl, "--script", script,
           "--work", work, "--seed", str(seed), "--threads", str(threads)]
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0 or "CALLER_OK" not in r.stdout:
        raise RuntimeError(f"{impl}/{script} failed ({r.returncode}):\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}")
    return r.stdout


def parse_losses(text):
    """Loss values of the log lines train_speech_embedder.py:70-72 prints."""
    return [float(m) for m in re.findall(r"\tLoss:([-0-9.naninf]+)\t", text)]


def parse_eer(text):
    """[(EER, thres, FAR, FRR)] per batch (train_speech_embedder.py:151) and the final mean (:154)."""
    rows = [tuple(float(v) for v in m) for m in
            re.findall(r"EER : ([0-9.]+) \(thres:([0-9.]+), FAR:([0-9.]+), FRR:([0-9.]+)\)", text)]
    final = re.findall(r"EER across \d+ epochs: ([0-9.]+)", text)
    return rows, (float(final[0]) if final else None)
