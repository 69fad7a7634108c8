"""B200-native (sm_100a) GE2E speaker-verification hot path.

Drop-in for hwidong-na/PyTorch_Speaker_Verification's ``speech_embedder_net`` / ``utils`` loss API,
backed by hand-written CUDA kernels behind the C ABI in include/svb200.h.  No CPU fallback.
"""
__all__ = ["_lib"]
