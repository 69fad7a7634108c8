"""CPU: every oracle function against the fixtures produced by the reference itself
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

import _inputs as I
from oracle import dvector, eer, embedder, ge2e

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_toy_block_matches_reference():
    """utils.py:166-173 toy input: vectorised, loop ("prior") and oracle versions agree."""
    g = load("toy.npz")
    E = g["E"]
    assert np.array_equal(g["cossim"], g["cossim_prior"])
    assert float(g["loss"]) == float(g["loss_prior"])
    C = ge2e.get_centroids(E)
    np.testing.assert_array_equal(C, g["centroids"])
    cos = ge2e.get_cossim(E, C)
    np.testing.assert_allclose(cos, g["cossim"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(ge2e.get_cossim_loops(E, C), g["cossim"], rtol=0, atol=1e-7)
    loss, per = ge2e.calc_loss(np.float32(1.0) * cos + np.float32(0.0))
    np.testing.assert_allclose(per, g["per"], rtol=1e-6)
    np.testing.assert_allclose(loss, g["loss"], rtol=1e-6)
    np.testing.assert_allclose(ge2e.calc_loss_loops(cos)[1], g["per"], rtol=1e-6)
    assert abs(float(g["loss"]) - 5.250094413757324) < 1e-6


@pytest.mark.parametrize("name", list(I.GE2E_CASES))
def test_ge2e_closed_form_matches_reference_autograd(name):
    g = load(f"ge2e_{name}.npz")
    N, M, D, kind, w, b = I.GE2E_CASES[name]
    E = I.ge2e_embeddings(N, M, D, kind)
    assert abs(E.astype(np.float64).sum() - float(g["in_sum"])) < 1e-9, "input regeneration differs"
    assert abs((E.astype(np.float64) ** 2).sum() - float(g["in_sq"])) < 1e-9
    # float64 oracle vs float64 reference: tight
    r64 = ge2e.ge2e_fwd_bwd(E.astype(np.float64), w, b)
    assert rel(r64["loss"], g["loss_f64"]) < 1e-12
    assert rel(r64["dw"], g["dw_f64"]) < 1e-10
    assert rel(r64["db"], g["db_f64"]) < 1e-7          # ill-conditioned sum
    assert rel(r64["dE"], g["dE_f64"]) < 2e-6           # fixture stores dE rounded to float32
    assert rel(r64["cos"], g["cos_f64"]) < 2e-6
    # float32 oracle vs float32 reference: 1e-5 (the north_star tolerance)
    r32 = ge2e.ge2e_fwd_bwd(E, w, b)
    assert rel(r32["loss"], g["loss_f32"]) < 1e-5
    assert rel(r32["dw"], g["dw_f32"]) < 1e-5
    assert rel(r32["dE"], g["dE_f64"]) < 1e-5
    assert rel(ge2e.ge2e_loss(E, w, b), g["loss_f32"]) < 1e-5
    if "cos_f32" in g:
        np.testing.assert_allclose(ge2e.get_cossim(E, ge2e.get_centroids(E)), g["cos_f32"], atol=3e-7, rtol=0)
        np.testing.assert_allclose(r32["per"], g["per_f32"], rtol=1e-5, atol=1e-6)


def _sd():
    return embedder.init_state_dict(seed=0)


def test_init_restatement_matches_reference_checksums():
    g = load("embedder_c1.npz")
    sd = _sd()
    assert list(sd) == embedder.PARAM_NAMES
    np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], g["w_sum"], rtol=0, atol=1e-9)
    np.testing.assert_allclose([float((v.double() ** 2).sum()) for v in sd.values()], g["w_sq"], rtol=0, atol=1e-9)


def test_embedder_explicit_matches_reference_forward():
    g = load("embedder_c1.npz")
    x = I.logmel(20, 180, seed=1234)
    assert abs(x.astype(np.float64).sum() - float(g["in_sum"])) < 1e-9
    with torch.no_grad():
        emb = embedder.embedder_explicit(torch.tensor(x), _sd()).numpy()
        emb24 = embedder.embedder_explicit(torch.tensor(I.logmel(7, 24, seed=4321)), _sd()).numpy()
        emb64 = embedder.embedder_explicit(torch.tensor(x[:3]).double(), _sd()).numpy()
    err = np.linalg.norm(emb - g["emb"], axis=1) / np.linalg.norm(g["emb"], axis=1)
    assert err.max() < 2e-5, err.max()
    assert (np.linalg.norm(emb24 - g["emb_T24"], axis=1)).max() < 2e-5
    assert emb64.dtype == np.float32 and np.linalg.norm(emb64 - g["emb_from_f64_input"], axis=1).max() < 2e-5


def test_embedder_saturated_weights():
    g = load("embedder_saturated.npz")
    sat = I.saturating_weights({k: v.numpy() for k, v in _sd().items()})
    sd = {k: torch.tensor(v) for k, v in sat.items()}
    with torch.no_grad():
        emb = embedder.embedder_explicit(torch.tensor(I.logmel(20, 180, seed=1234)), sd).numpy()
    assert (np.linalg.norm(emb - g["emb"], axis=1)).max() < 5e-5


def test_train_step_gradients_match_reference():
    """Explicit-cell autograd + closed-form GE2E vs the reference's loss.backward() (C1)."""
    g = load("embedder_c1.npz")
    sd = {k: v.clone().requires_grad_(True) for k, v in _sd().items()}
    x = torch.tensor(I.logmel(20, 180, seed=1234))
    emb = embedder.embedder_explicit(x, sd)
    w = torch.tensor(10.0, requires_grad=True)
    b = torch.tensor(-5.0, requires_grad=True)
    loss = embedder.library_ge2e_loss(emb.reshape(4, 5, -1), w, b)
    loss.backward()
    assert rel(loss.detach().numpy(), g["loss"]) < 1e-5
    assert rel(w.grad.numpy(), g["dw"]) < 1e-3
    for k in embedder.PARAM_NAMES:
        gr = sd[k].grad.numpy().reshape(-1)
        gn = float(g[f"gnorm.{k}"])
        assert abs(np.sqrt((gr.astype(np.float64) ** 2).sum()) - gn) < 2e-3 * gn, k
        np.testing.assert_allclose(gr[g[f"gidx.{k}"]], g[f"gval.{k}"], atol=2e-3 * gn / np.sqrt(gr.size) + 1e-7, rtol=2e-3)
    # and the closed-form GE2E gradient w.r.t. the embeddings agrees with autograd
    E = emb.detach().numpy().reshape(4, 5, -1)
    r = ge2e.ge2e_fwd_bwd(E, 10.0, -5.0)
    Et = torch.tensor(E, requires_grad=True)
    embedder.library_ge2e_loss(Et, torch.tensor(10.0), torch.tensor(-5.0)).backward()
    assert rel(r["dE"], Et.grad.numpy()) < 1e-5


@pytest.mark.parametrize("name", list(I.EER_CASES))
def test_eer_sweep_bit_exact(name):
    g = load("eer.npz")
    N, M, sigma, alpha, seed = I.EER_CASES[name]
    enr, ver = I.eer_embeddings(N, M, sigma, alpha, seed)
    assert abs(enr.astype(np.float64).sum() + ver.astype(np.float64).sum() - float(g[f"{name}.in_sum"])) < 1e-9
    sim_o = ge2e.get_cossim(ver, ge2e.get_centroids(enr))
    if f"{name}.sim" in g:
        sim = g[f"{name}.sim"]
        np.testing.assert_allclose(sim_o, sim, rtol=0, atol=3e-7)
    else:
        sim = sim_o
        assert abs(sim_o.astype(np.float64).sum() - float(g[f"{name}.sim_sum"])) < 1e-2
    tup = eer.eer_sweep(sim)
    exp = g[f"{name}.tuple"]
    if f"{name}.sim" in g:        # fed the reference's own sim matrix: bit-exact
        assert [float(v) for v in tup] == [float(v) for v in exp], (tup, exp)
    else:
        np.testing.assert_allclose([float(v) for v in tup], exp, atol=2e-3)


def test_eer_float32_threshold_rounding():
    """Quirk 7: sim == float32(t) is NOT above t even when float32(t) > t as doubles."""
    t = eer.THRESHOLDS[7]
    sim = np.full((2, 1, 2), np.float32(t), dtype=np.float32)
    call, cdiag = eer.eer_counts(sim)
    assert call[7].sum() == 0 and call[6].sum() == 4


def test_dvector_windows_and_alignment():
    g = load("dvector.npz")
    for T, cnt in zip(g["win_Ts"], g["win_counts"]):
        assert len(dvector.window_starts(int(T))) == int(cnt)
        assert int(cnt) == (0 if T <= 24 else -(-(int(T) - 24) // 12))
    for T in (37, 160):
        S = np.log10(I.power_spec(T, seed=T) + 1e-6)
        # the reference computes log10(dot(eye, |sqrt(p)|**2) + 1e-6) in float32
        p = np.sqrt(I.power_spec(T, seed=T)) ** 2
        S = np.log10(np.dot(np.eye(40, dtype=np.float32), p) + 1e-6)
        np.testing.assert_array_equal(dvector.windows(S).astype(np.float32), g[f"win_T{T}"])
    assert dvector.windows(np.zeros((40, 24), np.float32)).shape == (0, 24, 40)
    for W in (1, 2, 3, 5, 13, 30, 82):
        out = dvector.align_embeddings(I.unit_rows(W, 256, seed=W))
        assert out.dtype == np.float64
        np.testing.assert_array_equal(out, g[f"align_W{W}"])
    sizes = [e - s for s, e in dvector.partitions(30)]
    assert sizes[:7] == [2, 3, 4, 3, 3, 4, 3]


def test_optimizer_tail_closed_form_matches_library_calls():
    """oracle.optim: the closed form the CUDA kernel follows vs. clip_grad_norm_ + torch.optim.SGD (the entry points
    train_speech_embedder.py:63-65 calls), with a group that clips and one that does not."""
    from oracle import optim as ooptim
    r = np.random.RandomState(5)
    g0 = [(r.randn(3072, 40).astype(np.float32), r.randn(3072, 40).astype(np.float32) * 0.5),
          (r.randn(3072).astype(np.float32), r.randn(3072).astype(np.float32))]
    g1 = [(np.float32(10.0).reshape(()), np.float32(0.3).reshape(())),
          (np.float32(-5.0).reshape(()), np.float32(-0.2).reshape(()))]
    groups = [(g0, 3.0), (g1, 1.0)]
    pl, gl, nl = ooptim.clip_sgd_library(groups, 0.01)
    pn, gn, nn_ = ooptim.clip_sgd_numpy(groups, 0.01)
    assert nl[0] > 3.0 and nl[1] < 1.0
    np.testing.assert_allclose(nl, nn_, rtol=1e-6)
    for a, b in zip(sum(pl, []) + sum(gl, []), sum(pn, []) + sum(gn, [])):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)


def test_front_end_oracle_properties():
    """oracle.frontend (PARITY UNPINNED, librosa absent): properties librosa documents -- hz_to_mel(8000) on the
    Slaney scale, unit-area triangular filters, periodic Hann (sum N/2) centred in the n_fft frame, 1 + n//hop frames,
    Parseval for a frame, and a pure tone landing in the mel band that contains its frequency."""
    from oracle import frontend as ofe
    assert abs(float(ofe.hz_to_mel(8000.0)) - 45.245640471924965) < 1e-9
    assert abs(float(ofe.mel_to_hz(ofe.hz_to_mel(4321.0))) - 4321.0) < 1e-6
    w = ofe.mel_filterbank()
    assert w.shape == (40, 257) and w.min() >= 0.0
    np.testing.assert_allclose(w.sum(axis=1) * (8000.0 / 256), 1.0, atol=0.06)      # area 1 in Hz (31.25 Hz bins)
    win = ofe.hann_window_padded()
    assert abs(win.sum() - 200.0) < 1e-9 and np.all(win[:56] == 0) and np.all(win[456:] == 0) and win[56] == 0.0
    r = np.random.RandomState(2)
    y = r.randn(1603)
    P = ofe.stft_power(y)
    assert P.shape == (257, 1 + 1603 // 160)
    frame = np.pad(y, 256, mode="reflect")[5 * 160:5 * 160 + 512] * win
    full = P[:, 5].sum() * 2 - P[0, 5] - P[256, 5]                                  # two-sided spectrum energy
    assert abs(full / 512 - (frame ** 2).sum()) < 1e-9 * full
    t = np.arange(8000) / 16000.0
    S = ofe.log_mel(0.3 * np.sin(2 * np.pi * 1000.0 * t))
    edges = ofe.mel_to_hz(np.linspace(ofe.hz_to_mel(0.0), ofe.hz_to_mel(8000.0), 42))
    band = int(np.argmax(S[:, 20]))
    assert edges[band] < 1000.0 < edges[band + 2]


def test_centroid_functions_bit_level():
    """utils.get_centroids / get_utterance_centroids (utils.py:27-29, 40-58): the oracle's operation-order restatement
    reproduces the reference's float32 bit patterns, also where M leaves torch's sequential-sum regime."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_centroids import CASES, bits_checksum, case_input
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "centroids.npz"))
    for (N, M, D, seed) in CASES:
        E = case_input(N, M, D, seed)
        tag = f"{N}x{M}x{D}"
        C, U = ge2e.get_centroids_bitwise(E), ge2e.get_utterance_centroids_bitwise(E)
        assert (bits_checksum(C) == g[f"{tag}.C_sum"]).all(), tag
        assert (bits_checksum(U) == g[f"{tag}.U_sum"]).all(), tag
        np.testing.assert_array_equal(U.ravel()[g[f"{tag}.U_idx"]], g[f"{tag}.U_val"])
        if f"{tag}.U" in g:
            np.testing.assert_array_equal(U, g[f"{tag}.U"])
            np.testing.assert_array_equal(C, g[f"{tag}.C"])
        np.testing.assert_allclose(U, ge2e.get_utterance_centroids(E.astype(np.float64)), rtol=2e-4, atol=2e-5)
