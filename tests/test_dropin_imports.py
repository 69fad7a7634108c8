"""CPU (build container only): with dropin/ first on sys.path the UNMODIFIED reference train_speech_embedder.py
binds our classes; skipped where /root/reference does not exist (GPU box)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, types, os
import yaml
_orig = yaml.load_all
yaml.load_all = lambda stream, Loader=None: _orig(stream, Loader=Loader or yaml.FullLoader)   # PyYAML>=6 shim
sys.modules.setdefault("librosa", types.ModuleType("librosa"))                                  # absent dependency
os.chdir("/root/reference")
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(sys.argv[1], "pytorch_speaker_verification_b200", "dropin"))
import train_speech_embedder as T
import speech_embedder_net, utils
assert T.SpeechEmbedder.__module__.startswith("pytorch_speaker_verification_b200"), T.SpeechEmbedder.__module__
assert T.GE2ELoss.__module__.startswith("pytorch_speaker_verification_b200")
assert T.get_cossim.__module__.startswith("pytorch_speaker_verification_b200")
assert callable(utils.mfccs_and_spec) and utils.mfccs_and_spec.__module__ == "_reference_utils"
net = T.SpeechEmbedder()                      # zero-arg ctor reads the reference's hp singleton
from hparam import hparam as hp
assert net.LSTM_stack.input_size == hp.data.nmels and net.LSTM_stack.hidden_size == hp.model.hidden
assert net.projection.out_features == hp.model.proj
assert len(net.state_dict()) == 14
crit = T.GE2ELoss("cpu")
assert [tuple(p.shape) for p in crit.parameters()] == [(), ()] and float(crit.w) == 10.0 and float(crit.b) == -5.0
print("DROPIN_OK")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_reference_scripts_bind_the_dropin():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=120)
    assert "DROPIN_OK" in r.stdout, r.stdout + r.stderr
