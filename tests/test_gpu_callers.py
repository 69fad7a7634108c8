"""Acceptance run of the drop-in boundary: the reference's UNMODIFIED scripts --
train_speech_embedder.py ``train()`` (:19-90) and ``test()`` (:92-154) and dvector_create.py's script body (:75-122) --
are executed as ``__main__`` twice on the same synthetic data and seeds: once binding the reference's own
speech_embedder_net / utils on the CPU, once binding pytorch_speaker_verification_b200/dropin (hp.device = "cuda" for
training; test() and dvector_create keep the model and inputs on the CPU like the reference does, the drop-in stages
them to the GPU).  Logged losses, EER lines, checkpoints and the saved d-vector sequences are compared.

The reference sources come from baseline/_ref/ (staged unmodified by ``__graft_entry__.build()`` in the build
container; git-ignored, travels with the gpurun snapshot).
"""
import os
import shutil
import tempfile

import numpy as np
import pytest
import torch

import _callers as C

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not C.reference_staged(), reason="baseline/_ref not staged (run __graft_entry__.build() "
                                                                  "in the build container)")]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.fixture(scope="module")
def runs():
    work = tempfile.mkdtemp(prefix="svb_callers_")
    C.make_tisv_dirs(work)
    C.make_wav_dirs(work)
    out = {"work": work}
    common = dict(train__epochs=2, train__log_interval=1, train__checkpoint_interval=1, train__num_workers=0,
                  test__epochs=2, test__num_workers=0)
    # ---- train(): 2 epochs x 2 batches of N=4 x M=5 (config.yaml defaults), checkpoint after every epoch
    for impl, dev in (("reference", "cpu"), ("dropin", "cuda")):
        C.write_config(work, training=True, device=dev, train__checkpoint_dir=f"./ckpt_{impl}",
                       train__log_file=f"./ckpt_{impl}/Stats", **common)
        out[f"train_{impl}"] = C.run_caller(impl, "train_speech_embedder", work)
    final = "final_epoch_2_batch_id_2.model"
    # ---- test() and dvector_create: every impl with the checkpoint written by the OTHER impl and by itself
    for impl in ("reference", "dropin"):
        for ck in ("reference", "dropin"):
            C.write_config(work, training=False, device="cpu", model__model_path=f"./ckpt_{ck}/{final}", **common)
            out[f"test_{impl}_{ck}"] = C.run_caller(impl, "train_speech_embedder", work)
            if ck == "reference":
                out[f"dvec_{impl}"] = C.run_caller(impl, "dvector_create", work)
                d = os.path.join(work, f"dvec_{impl}")
                os.makedirs(d, exist_ok=True)
                for n in ("train_sequence", "train_cluster_id", "test_sequence", "test_cluster_id"):
                    shutil.move(os.path.join(work, n + ".npy"), os.path.join(d, n + ".npy"))
    yield out
    shutil.rmtree(work, ignore_errors=True)


def test_unmodified_train_runs_through_the_dropin(runs):
    """train_speech_embedder.py:19-90 unchanged: same batches (same seeds) -> the logged losses agree step by step and
    the saved checkpoints are interchangeable (14 keys, CPU tensors, float32) and close."""
    lr, ld = C.parse_losses(runs["train_reference"]), C.parse_losses(runs["train_dropin"])
    print("losses reference:", lr, "drop-in:", ld)
    assert len(lr) == len(ld) == 4 and all(np.isfinite(ld))
    assert abs(ld[0] - lr[0]) <= 2e-3 * abs(lr[0])            # same weights: embeddings 1e-3 -> loss
    for a, b in zip(ld[1:], lr[1:]):                           # after 1..3 clipped SGD steps on bf16-BPTT gradients
        assert abs(a - b) <= 2e-2 * abs(b), (ld, lr)
    with open(os.path.join(runs["work"], "ckpt_dropin", "Stats")) as f:      # log file of :73-75
        assert len(C.parse_losses(f.read())) == 4
    for name in ("ckpt_epoch_1_batch_id_2.pth", "ckpt_epoch_2_batch_id_2.pth", "final_epoch_2_batch_id_2.model"):
        a = torch.load(os.path.join(runs["work"], "ckpt_reference", name))
        b = torch.load(os.path.join(runs["work"], "ckpt_dropin", name))
        assert list(a.keys()) == list(b.keys()) and len(a) == 14
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype == torch.float32
            assert b[k].device.type == "cpu"                   # saved after .eval().cpu() (:78,:85)
    # the update direction matches: parameter deltas of the two runs agree
    torch.manual_seed(0)
    import pytorch_speaker_verification_b200 as svb
    init = svb.SpeechEmbedder().state_dict()                   # same init as both runs (seed 0, reference RNG order)
    a = torch.load(os.path.join(runs["work"], "ckpt_reference", "final_epoch_2_batch_id_2.model"))
    b = torch.load(os.path.join(runs["work"], "ckpt_dropin", "final_epoch_2_batch_id_2.model"))
    errs = {k: rel_l2((b[k] - init[k]).numpy(), (a[k] - init[k]).numpy()) for k in a}
    print("parameter-delta rel-L2 after 4 steps:", {k: round(float(v), 4) for k, v in errs.items()})
    assert max(errs.values()) < 0.15, errs


def test_unmodified_test_eer_runs_through_the_dropin(runs):
    """train_speech_embedder.py:92-154 unchanged (model and batches on the CPU, as the reference runs it), with the
    checkpoint the reference wrote and with the one the drop-in wrote, loaded by either implementation."""
    for ck in ("reference", "dropin"):
        rr, fr = C.parse_eer(runs[f"test_reference_{ck}"])
        rd, fd = C.parse_eer(runs[f"test_dropin_{ck}"])
        print(f"checkpoint by {ck}: reference EER rows {rr} mean {fr}; drop-in {rd} mean {fd}")
        assert len(rr) == len(rd) == 4 and fr is not None and fd is not None
        assert any(0.0 < row[2] < 1.0 or 0.0 < row[3] < 1.0 for row in rr)          # the sweep is not degenerate
        for a, b in zip(rd, rr):
            assert abs(a[0] - b[0]) <= 0.045 and abs(a[1] - b[1]) <= 0.021, (a, b)  # one count / one threshold step
        assert abs(fd - fr) <= 0.03


def test_unmodified_dvector_create_runs_through_the_dropin(runs):
    """dvector_create.py:75-122 unchanged: same partition counts and cluster ids, aligned d-vectors within 1e-3."""
    w = runs["work"]
    for n in ("train", "test"):
        a = np.load(os.path.join(w, "dvec_reference", f"{n}_sequence.npy"))
        b = np.load(os.path.join(w, "dvec_dropin", f"{n}_sequence.npy"))
        ia = np.load(os.path.join(w, "dvec_reference", f"{n}_cluster_id.npy"))
        ib = np.load(os.path.join(w, "dvec_dropin", f"{n}_cluster_id.npy"))
        assert a.shape == b.shape and a.dtype == b.dtype == np.float64 and a.shape[0] > 0
        assert ia.dtype == ib.dtype and (ia == ib).all()
        err = (np.linalg.norm(a - b, axis=1) / np.linalg.norm(a, axis=1)).max()
        print(n, "sequence", a.shape, "max per-row rel-L2", err)
        assert err < 1e-3, err
