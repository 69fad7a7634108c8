"""CPU oracle for the GE2E hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference hot path
(hwidong-na/PyTorch_Speaker_Verification: speech_embedder_net.py, utils.py,
the EER sweep of train_speech_embedder.py and the windowing/alignment of
dvector_create.py).  It is the checker for the CUDA path, never the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
    ``--impl reference`` legs of ``bench.py`` may import it;
  * nothing under ``pytorch_speaker_verification_b200/`` imports it, and the
    product path raises when the CUDA library is missing.

Where the arithmetic of the reference lives in a third-party dependency
(``torch``: nn.LSTM, nn.Linear, F.cosine_similarity; README.md:14 pins
"PyTorch 0.4.1", this image ships torch 2.11.0) the oracle restates the
published algorithm explicitly (``embedder.py``: LSTM cell equations,
``ge2e.py``: closed-form loss and gradient) and is pinned against the
reference itself executed in the build container:
``tests/golden/make_golden.py`` imports /root/reference (with two environment
shims, no edits) and writes the fixtures in ``tests/golden/``; the
``-m "not gpu"`` tests check every oracle function against them.

Parity status: PINNED against outputs of the reference run in the build
container (the reference ships no tests or golden vectors of its own; its only
executable check is the non-asserting toy block utils.py:166-173, reproduced
in tests/golden/toy.npz).
"""
