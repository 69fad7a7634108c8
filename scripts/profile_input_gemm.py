"""The LSTM input projection as a standalone batched GEMM ([102400 x 768] . [768 x 3072], CTA-pair tcgen05 tiles) for
an ncu capture of its tensor-pipe utilisation (north_star target: >= 80 %)."""
import ctypes, sys
import torch
sys.path.insert(0, ".")
from pytorch_speaker_verification_b200 import _lib
from pytorch_speaker_verification_b200._lib import ptr, stream_ptr
L = _lib.lib()
M, N, K = 640 * 160, 3072, 768
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
C = torch.empty(M, N, device="cuda")
PA = (ctypes.c_void_p * 1)(A.data_ptr()); PB = (ctypes.c_void_p * 1)(B.data_ptr())
for _ in range(3):
    assert L.svb_gemm_bf16_2cta(PA, PB, 1, ptr(C), None, M, N, K, ctypes.c_int64(K), ctypes.c_int64(K),
                                ctypes.c_int64(N), 0, stream_ptr()) == 0
torch.cuda.synchronize()
ref = A[:256].float() @ B.float().t()
print("max err", (C[:256] - ref).abs().max().item())
