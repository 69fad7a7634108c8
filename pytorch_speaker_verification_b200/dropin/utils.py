"""Drop-in module named like the reference's utils.py.  The loss math (get_centroids, get_cossim, calc_loss,
get_utterance_centroids) comes from the B200 implementation; every other attribute (mfccs_and_spec, normalize_0_1,
the loop "prior" variants; data_load.py:17 imports mfccs_and_spec from here) is forwarded untouched to the reference's
own utils.py, located as the next ``utils.py`` on sys.path."""
import importlib.util
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.append(_REPO)

from pytorch_speaker_verification_b200.utils import (calc_loss, get_centroids, get_cossim,  # noqa: E402,F401
                                                     get_utterance_centroids)

_HERE = os.path.dirname(os.path.abspath(__file__))
_ref = None


def _reference_utils():
    global _ref
    if _ref is None:
        for d in sys.path:
            cand = os.path.join(d or ".", "utils.py")
            if os.path.isfile(cand) and os.path.abspath(os.path.dirname(cand)) != _HERE:
                spec = importlib.util.spec_from_file_location("_reference_utils", cand)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                _ref = mod
                break
        else:
            raise ImportError("reference utils.py not found on sys.path")
    return _ref


def __getattr__(name):          # PEP 562: anything we do not define is the reference's
    return getattr(_reference_utils(), name)
