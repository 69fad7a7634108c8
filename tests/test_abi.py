"""CPU: the C-ABI library builds, loads and exports every symbol include/svb200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "svb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for s in ("svb_embedder_forward", "svb_embedder_backward", "svb_ge2e", "svb_eer_counts", "svb_eer_finish",
              "svb_dvector_windows", "svb_segment_mean", "svb_centroids", "svb_calc_loss"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from pytorch_speaker_verification_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/svb200.h but not exported"
    assert lib.svb_arch() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pytorch_speaker_verification_b200 as svb
    from pytorch_speaker_verification_b200._lib import SvbError
    with pytest.raises(SvbError):
        svb.SpeechEmbedder()(torch.zeros(2, 30, 40))
    with pytest.raises(SvbError):
        svb.GE2ELoss("cpu")(torch.randn(3, 2, 8))
    with pytest.raises(SvbError):
        svb.get_cossim(torch.randn(3, 2, 8), torch.randn(3, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch_speaker_verification_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
