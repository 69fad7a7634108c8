"""Dev: run-to-run bitwise stability of the whole train step (persistent forward + BPTT + weight gradients beside BPTT)
over many repetitions and shapes -- a lost hand-off or a stale read in the pipelines would show as a differing bit."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _inputs as I
import pytorch_speaker_verification_b200 as svb
torch.manual_seed(0)
net = svb.SpeechEmbedder().cuda()
crit = svb.GE2ELoss("cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for (N, M, T) in ((64, 10, 160), (64, 10, 37), (13, 10, 160), (3, 7, 90)):
    x = torch.tensor(I.logmel(N * M, T, seed=5)).cuda()
    ref = None
    bad = 0
    for r in range(reps):
        net.zero_grad(set_to_none=True); crit.zero_grad(set_to_none=True)
        loss = crit(net(x).reshape(N, M, -1)); loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in net.parameters()] + [crit.w.grad.reshape(1), loss.detach().reshape(1)])
        if ref is None: ref = g.clone()
        elif not torch.equal(g, ref): bad += 1
    print(f"N={N} M={M} T={T}: {reps} steps, {bad} differ from the first, finite {bool(torch.isfinite(ref).all())}", flush=True)
