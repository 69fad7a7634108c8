"""CPU, world_size 2, gloo: host-side logic of the multi-GPU path (speaker sharding, d-vector all-gather, slice of
dL/dE kept per rank, SUM all-reduce of parameter gradients).  The fused CUDA kernel is replaced by the oracle's
closed form through GlobalGE2ELoss's injection point (tests may use the oracle; the product default is the kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _inputs as I
from oracle import ge2e as oge2e


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_compute(E, w, b):
    r = oge2e.ge2e_fwd_bwd(E.numpy(), float(w), float(b))
    f = lambda v: torch.tensor(np.asarray(v, dtype=np.float32))
    return f(r["loss"]), f(r["dE"]), f(r["dw"]), f(r["db"])


def _rows_compute(E_local, C_all, w, b, col0):
    """Design A's arithmetic on the CPU: this rank's rows against all centroids, own column col0 + j replaced by the
    leave-one-out cosine (utils.py:72-115 restricted to a row shard), through torch autograd."""
    import torch.nn.functional as F
    E = E_local.clone().requires_grad_(True)
    C = C_all.clone().requires_grad_(True)
    w_ = w.clone().requires_grad_(True)
    b_ = b.clone().requires_grad_(True)
    N, M, D = E.shape
    with torch.enable_grad():                      # (called from inside an autograd.Function.forward)
        U = (E.sum(dim=1, keepdim=True) - E) / (M - 1)
        cos_same = F.cosine_similarity(E, U, dim=2)
        cos = F.cosine_similarity(E.unsqueeze(2), C.view(1, 1, -1, D), dim=3)
        idx = torch.arange(N)
        cos[idx, :, col0 + idx] = cos_same
        S = w_ * (cos + 1e-6) + b_
        loss = (torch.log(torch.exp(S).sum(dim=2) + 1e-6) - S[idx, :, col0 + idx]).sum()
        loss.backward()
    red = torch.cat([C.grad.reshape(-1), loss.detach().view(1), w_.grad.view(1), b_.grad.view(1)])
    return red, E.grad


def _worker(rank, world, port, out, mode="gather"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_speaker_verification_b200.dist import GlobalGE2ELoss, allreduce_gradients, speaker_shard

    N, M, D = 8, 3, 16
    E = I.ge2e_embeddings(N, M, D, "raw")
    lo, hi = speaker_shard(N, rank, world)
    # a stand-in "embedder": emb = x @ W, shared W, so that parameter gradients need the all-reduce
    torch.manual_seed(0)
    W = torch.nn.Parameter(torch.eye(D) + 0.01 * torch.randn(D, D))

    class Crit(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(3.0))
            self.b = torch.nn.Parameter(torch.tensor(-1.0))

    crit = Crit()
    gl = GlobalGE2ELoss(crit, compute=_oracle_compute if mode == "gather" else _rows_compute, mode=mode)
    x_local = torch.tensor(E[lo:hi])
    loss = gl(x_local @ W)
    (loss * 0.5).backward()
    allreduce_gradients([W])
    out[rank] = (loss.item(), W.grad.clone(), crit.w.grad.item(), crit.b.grad.item())
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("mode", ["rows", "gather"])
def test_global_ge2e_two_ranks_matches_single_process(mode):
    """mode "rows": centroid all-gather + own rows against all centroids + all-reduce of [dC | loss, dw, db];
    mode "gather": d-vector all-gather + the whole global batch on every rank."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, mode), nprocs=world, join=True)
    # single-process reference: whole batch through the oracle-backed autograd path
    N, M, D = 8, 3, 16
    E = torch.tensor(I.ge2e_embeddings(N, M, D, "raw"))
    torch.manual_seed(0)
    W = torch.nn.Parameter(torch.eye(D) + 0.01 * torch.randn(D, D))
    emb = E @ W
    r = oge2e.ge2e_fwd_bwd(emb.detach().numpy(), 3.0, -1.0)
    emb.backward(torch.tensor(0.5 * r["dE"]))
    for rank in range(world):
        loss, gW, gw, gb = out[rank]
        assert abs(loss - float(r["loss"])) < 1e-4 * abs(float(r["loss"]))      # same global loss on every rank
        assert torch.allclose(gW, W.grad, rtol=1e-4, atol=1e-6)                  # SUM over ranks == global gradient
        assert abs(gw - 0.5 * float(r["dw"])) < 1e-4 * abs(float(r["dw"]))       # identical on every rank, not summed
        assert abs(gb - 0.5 * float(r["db"])) < 1e-3 * abs(float(r["db"])) + 1e-6    # (db: sum of +-O(1) terms cancelling to ~5e-6; fp32 autograd noise in "rows" mode)


def test_speaker_shard():
    from pytorch_speaker_verification_b200.dist import speaker_shard
    assert [speaker_shard(512, r, 8) for r in (0, 7)] == [(0, 64), (448, 512)]
    with pytest.raises(ValueError):
        speaker_shard(10, 0, 4)


def test_flat_view_recognises_slices_of_one_storage():
    """allreduce_gradients reduces ONE flat tensor when the gradients are contiguous slices that tile one allocation
    (what svb200::embedder_bwd returns: custom-op outputs, no autograd ``_base``), and refuses anything else."""
    from pytorch_speaker_verification_b200.dist import _flat_view
    flat = torch.arange(24, dtype=torch.float32)
    parts = [flat[0:6].view(2, 3), flat[6:10], flat[10:24].view(7, 2)]
    parts = [torch.as_strided(flat, p.shape, p.stride(), p.storage_offset()) for p in parts]      # no _base, as from an op
    v = _flat_view(parts[::-1])
    assert v is not None and v.numel() == 24 and v.data_ptr() == flat.data_ptr()
    v.mul_(2)
    assert float(parts[1][0]) == 12.0
    assert _flat_view([parts[0], parts[2]]) is None                                     # gap
    assert _flat_view([parts[0], torch.zeros(4)]) is None                               # different storage
    assert _flat_view([flat[0:6].view(2, 3).t(), flat[6:24]]) is None                  # non-contiguous member
