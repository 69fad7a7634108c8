"""CPU check of the caller harness itself (tests/_run_caller.py, tests/_callers.py): the UNMODIFIED reference
train_speech_embedder.py runs as ``__main__`` on synthetic speaker files with the reference's own modules and writes
its log and checkpoints.  The GPU acceptance tests (tests/test_gpu_callers.py) run the same harness with the drop-in.
Skipped where baseline/_ref/ has not been staged."""
import os
import shutil
import tempfile

import numpy as np
import pytest
import torch

import _callers as C

pytestmark = pytest.mark.skipif(not C.reference_staged(), reason="baseline/_ref not staged")


def test_harness_runs_the_reference_train_script():
    work = tempfile.mkdtemp(prefix="svb_callers_cpu_")
    try:
        C.make_tisv_dirs(work, n_train=4, n_test=1, utts=6, frames=165)
        C.write_config(work, training=True, device="cpu", train__epochs=1, train__log_interval=1,
                       train__checkpoint_interval=1, train__num_workers=0, train__checkpoint_dir="./ckpt",
                       train__log_file="./ckpt/Stats")
        out = C.run_caller("reference", "train_speech_embedder", work, threads=4)
        losses = C.parse_losses(out)
        assert len(losses) == 1 and np.isfinite(losses[0]) and 0 < losses[0] < 4 * 5 * np.log(4) * 1.5
        sd = torch.load(os.path.join(work, "ckpt", "final_epoch_1_batch_id_1.model"))
        assert len(sd) == 14 and sd["LSTM_stack.weight_hh_l2"].shape == (3072, 768)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def test_vad_stub_segments_merge_into_two_runs():
    """The VAD stub of the harness yields contiguous 0.4 s segments with one gap (dvector_create.concat_segs:24-36
    must see exactly two voiced runs)."""
    import sys
    import _run_caller as R
    saved = {k: sys.modules.get(k) for k in ("librosa", "librosa.core", "librosa.filters", "VAD_segments")}
    import yaml
    orig = yaml.load_all
    try:
        R.install_shims()
        work = tempfile.mkdtemp(prefix="svb_vad_")
        C.make_wav_dirs(work, n_speakers=1, files=1)
        times, segs = sys.modules["VAD_segments"].VAD_chunk(2, os.path.join(work, "TIMIT", "TRAIN", "DR1", "SPK00", "utt0.wav"))
        assert len(times) == len(segs) == 8
        joins = [times[i][1] == times[i + 1][0] for i in range(len(times) - 1)]
        assert joins.count(False) == 1
        S = sys.modules["librosa"].core.stft(y=segs[0], n_fft=512, win_length=400, hop_length=160)
        assert S.shape == (257, 1 + len(segs[0]) // 160)
        shutil.rmtree(work, ignore_errors=True)
    finally:
        yaml.load_all = orig
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
