"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's CPU legs, never by the product).

CPU restatement of the optimizer tail of the reference's training step, train_speech_embedder.py:63-65 with the
optimizer built at :33-36: two clip_grad_norm_ calls (embedder parameters at 3.0, GE2E w/b at 1.0) and a plain
SGD step.  Two forms: `clip_sgd_library` calls the same torch entry points the reference calls; `clip_sgd_numpy`
is the closed form the CUDA kernel follows (csrc/optim.cu)."""
import numpy as np
import torch


def clip_sgd_library(groups, lr):
    """groups: list of (list of (param ndarray, grad ndarray), max_norm).  Returns (new params, clipped grads, norms)
    per group, computed by torch.nn.utils.clip_grad_norm_ + torch.optim.SGD on CPU float32."""
    tgroups = []
    for pg, _ in groups:
        ps = []
        for p, g in pg:
            t = torch.nn.Parameter(torch.tensor(np.asarray(p, dtype=np.float32)))
            t.grad = torch.tensor(np.asarray(g, dtype=np.float32))
            ps.append(t)
        tgroups.append(ps)
    opt = torch.optim.SGD([{"params": ps} for ps in tgroups], lr=lr)      # train_speech_embedder.py:33-36
    norms = [float(torch.nn.utils.clip_grad_norm_(ps, mn)) for ps, (_, mn) in zip(tgroups, groups)]   # :63-64
    opt.step()                                                             # :65
    return ([[p.detach().numpy() for p in ps] for ps in tgroups],
            [[p.grad.numpy() for p in ps] for ps in tgroups], norms)


def clip_sgd_numpy(groups, lr):
    out_p, out_g, norms = [], [], []
    for pg, mn in groups:
        tot = np.sqrt(sum(float((np.asarray(g, dtype=np.float64) ** 2).sum()) for _, g in pg))
        coef = np.float32(min(1.0, mn / (tot + 1e-6)))
        gs = [(np.asarray(g, dtype=np.float32) * coef).astype(np.float32) for _, g in pg]
        out_g.append(gs)
        out_p.append([(np.asarray(p, dtype=np.float32) - np.float32(lr) * g).astype(np.float32)
                      for (p, _), g in zip(pg, gs)])
        norms.append(tot)
    return out_p, out_g, norms
