"""GPU, 2 ranks over NCCL: the sharded global GE2E step equals the single-GPU step on the concatenated batch
(BASELINE configs[2] semantics at world_size 2).  Skipped on a 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _inputs as I

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pytorch_speaker_verification_b200 as svb
    from pytorch_speaker_verification_b200.dist import GlobalGE2ELoss, allreduce_gradients, speaker_shard
    torch.manual_seed(0)
    net = svb.SpeechEmbedder().cuda()
    crit = svb.GE2ELoss("cuda")
    N, M, T = 8, 4, 30
    x = torch.tensor(I.logmel(N * M, T, seed=42)).reshape(N, M, T, 40)
    lo, hi = speaker_shard(N, rank, world)
    emb = net(x[lo:hi].reshape(-1, T, 40).cuda())
    loss = GlobalGE2ELoss(crit)(emb.reshape(hi - lo, M, -1))          # default: centroid all-gather + row shards
    loss.backward()
    allreduce_gradients(list(net.parameters()))
    out[rank] = (loss.item(), net.LSTM_stack.weight_hh_l1.grad.cpu(), net.projection.weight.grad.cpu(),
                 crit.w.grad.item())
    # design B (d-vector all-gather, global batch on every rank) gives the same step
    first = [p.grad.clone() for p in net.parameters()]
    w_first = crit.w.grad.clone()
    net.zero_grad()
    crit.zero_grad()
    emb = net(x[lo:hi].reshape(-1, T, 40).cuda())
    lossB = GlobalGE2ELoss(crit, mode="gather")(emb.reshape(hi - lo, M, -1))
    lossB.backward()
    allreduce_gradients(list(net.parameters()))
    assert abs(lossB.item() - loss.item()) < 1e-5 * abs(loss.item())
    assert abs(crit.w.grad.item() - w_first.item()) < 1e-4 * abs(w_first.item())
    # dE of the two designs differs in the last bits, and BPTT stores dG as bf16: a 1e-6 relative perturbation of dE
    # flips roundings worth 2-4e-3 of a parameter gradient at this size (16 rows x 30 frames per rank, nothing to
    # average over) in BOTH BPTT implementations (scripts/check_bptt_sensitivity.py), so the designs are compared
    # inside the gradient tolerance of DESIGN.md section 2 (1.2e-2), not bit for bit
    for p, r in zip(net.parameters(), first):
        assert float((p.grad - r).norm() / r.norm()) < 1.2e-2
    # design A with both exchange steps over NVLink peer memory (symmetric buffers, no NCCL call on the GE2E path):
    # the same loss and, since the sums are taken in rank order from the same partials, the same gradients as the NCCL
    # form up to the summation order of the all-reduce
    net.zero_grad()
    crit.zero_grad()
    emb = net(x[lo:hi].reshape(-1, T, 40).cuda())
    glp = GlobalGE2ELoss(crit, mode="peer")
    lossP = glp(emb.reshape(hi - lo, M, -1))
    lossP.backward()
    allreduce_gradients(list(net.parameters()))
    assert glp._px not in (None, False), "peer memory was not used"
    assert abs(lossP.item() - loss.item()) < 1e-6 * abs(loss.item())
    assert abs(crit.w.grad.item() - w_first.item()) < 1e-5 * abs(w_first.item())
    for p, r in zip(net.parameters(), first):
        assert float((p.grad - r).norm() / r.norm()) < 1.2e-2
    for _ in range(3):                               # buffer reuse across steps (two barriers per step)
        l2 = glp(emb.detach().reshape(hi - lo, M, -1))
        assert abs(l2.item() - lossP.item()) < 1e-6 * abs(lossP.item())
    # two-shot all-reduce of the flat gradient buffer over NVLink peer memory (rank-ordered sums) against NCCL's
    net.zero_grad()
    crit.zero_grad()
    emb = net(x[lo:hi].reshape(-1, T, 40).cuda())
    GlobalGE2ELoss(crit)(emb.reshape(hi - lo, M, -1)).backward()
    local = [p.grad.clone() for p in net.parameters()]
    allreduce_gradients(list(net.parameters()), peer=True)
    from pytorch_speaker_verification_b200 import dist as svb_dist
    assert any(v is not None for v in svb_dist._PEER_ALLREDUCE.values()), "peer all-reduce was not used"
    for p, g in zip(net.parameters(), local):
        ref_sum = g.clone()
        dist.all_reduce(ref_sum, op=dist.ReduceOp.SUM)
        assert torch.allclose(p.grad, ref_sum, rtol=1e-6, atol=1e-9)
    for p, r in zip(net.parameters(), first):
        assert torch.equal(p.grad, r) or float((p.grad - r).norm() / r.norm()) < 1e-6
    # the same step with the bucketed all-reduce started from inside backward: identical sums
    from pytorch_speaker_verification_b200.dist import OverlappedGradReducer
    ref = first                                  # (same GE2E mode: the sums must be bit-identical)
    net.zero_grad()
    crit.zero_grad()
    emb = net(x[lo:hi].reshape(-1, T, 40).cuda())
    loss2 = GlobalGE2ELoss(crit)(emb.reshape(hi - lo, M, -1))
    reducer = OverlappedGradReducer()
    with reducer:
        loss2.backward()
    torch.cuda.synchronize()
    assert reducer.buckets == 4
    for p, r in zip(net.parameters(), ref):
        assert torch.equal(p.grad, r), "bucketed all-reduce differs from the single all-reduce"
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_step_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    import pytorch_speaker_verification_b200 as svb
    torch.manual_seed(0)
    net = svb.SpeechEmbedder().cuda()
    crit = svb.GE2ELoss("cuda")
    N, M, T = 8, 4, 30
    x = torch.tensor(I.logmel(N * M, T, seed=42)).cuda()
    loss = crit(net(x).reshape(N, M, -1))
    loss.backward()
    for rank in range(2):
        l, g1, gp, gw = out[rank]
        assert abs(l - loss.item()) < 1e-5 * abs(loss.item())
        # (through BPTT: bf16 dG roundings flip with the last bits of dE, see the comment in _worker)
        ref1 = net.LSTM_stack.weight_hh_l1.grad.cpu()
        assert float((g1 - ref1).norm() / ref1.norm()) < 1.2e-2
        assert torch.allclose(gp, net.projection.weight.grad.cpu(), rtol=2e-3, atol=1e-6)
        assert abs(gw - crit.w.grad.item()) < 1e-4 * abs(crit.w.grad.item())
