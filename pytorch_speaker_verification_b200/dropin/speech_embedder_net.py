"""Drop-in module: put this directory FIRST on sys.path (before the reference checkout) and the reference's
train_speech_embedder.py (:17) and dvector_create.py (:20) import the B200 implementation unchanged."""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.append(_REPO)

from pytorch_speaker_verification_b200.speech_embedder_net import GE2ELoss, SpeechEmbedder  # noqa: E402,F401
from pytorch_speaker_verification_b200.utils import calc_loss, get_centroids, get_cossim     # noqa: E402,F401
