"""CPU emulation of the BPTT path's roundings at C1 (N=4, M=5, T=180): which of them dominates the LSTM parameter
gradient error against exact (float64) autograd?  Knobs: gate stash format (bf16 / fp16 / exact), dG storage (bf16 /
exact), W^T operands (bf16 / exact), split-K partials (fp16 / exact), h operand of the weight gradients (bf16 / exact).
Forward is exact float64 so that only the backward roundings show."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import _inputs as I
from oracle import ge2e as oge2e
from oracle.embedder import init_state_dict

torch.set_num_threads(8)


def rnd(t, mode):
    if mode == "exact":
        return t
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[mode]
    return t.to(torch.float32).to(dt).to(t.dtype)


def forward(x, sd, L=3):
    st = []
    inp = x
    for l in range(L):
        w_ih, w_hh = sd[f"LSTM_stack.weight_ih_l{l}"], sd[f"LSTM_stack.weight_hh_l{l}"]
        bias = sd[f"LSTM_stack.bias_ih_l{l}"] + sd[f"LSTM_stack.bias_hh_l{l}"]
        H = w_hh.shape[1]
        B, T, _ = inp.shape
        gin = inp @ w_ih.t() + bias
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        hs, cs, gs = [h], [c], []
        for t in range(T):
            g = gin[:, t] + h @ w_hh.t()
            i, f, gg, o = g.split(H, dim=1)
            i, f, gg, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(gg), torch.sigmoid(o)
            c = f * c + i * gg
            h = o * torch.tanh(c)
            hs.append(h); cs.append(c); gs.append((i, f, gg, o))
        st.append((inp, hs, cs, gs))
        inp = torch.stack(hs[1:], dim=1)
    y = inp[:, -1] @ sd["projection.weight"].t() + sd["projection.bias"]
    return y, st


def backward(dh_last, st, sd, stash="bf16", dgm="bf16", wm="bf16", part="fp16", hm="bf16", L=3, recompute_c=False):
    grads = {}
    B, T = st[0][0].shape[:2]
    dx_above = None
    for l in reversed(range(L)):
        inp, hs, cs, gs = st[l]
        w_ih, w_hh = sd[f"LSTM_stack.weight_ih_l{l}"], sd[f"LSTM_stack.weight_hh_l{l}"]
        H = w_hh.shape[1]
        whhT, wihT = rnd(w_hh, wm), rnd(w_ih, wm)
        dG = [None] * T
        dc_run = torch.zeros(B, H, dtype=inp.dtype)
        dG_next = None
        for t in reversed(range(T)):
            dh = torch.zeros(B, H, dtype=inp.dtype)
            if dG_next is not None:
                # split-K over 4 slices of the 4H gate columns, partials rounded
                prod = [rnd(dG_next[:, k * H:(k + 1) * H] @ whhT[k * H:(k + 1) * H], part) for k in range(4)]
                dh = prod[0] + prod[1] + prod[2] + prod[3]
            if dx_above is not None:
                dh = dh + dx_above[t]
            elif t == T - 1:
                dh = dh + dh_last
            i, f, gg, o = (rnd(v, stash) for v in gs[t])
            # recompute_c: c_t = f c_{t-1} + i g from the ROUNDED stash instead of reading the stored fp32 c_t
            tc = torch.tanh(f * cs[t] + i * gg if recompute_c else cs[t + 1])
            dc = dh * o * (1 - tc * tc) + dc_run
            dO = dh * tc * o * (1 - o)
            di = dc * gg * i * (1 - i)
            df = dc * cs[t] * f * (1 - f)
            dg = dc * i * (1 - gg * gg)
            dc_run = dc * f
            full = torch.cat([di, df, dg, dO], dim=1)
            grads.setdefault(f"bias_l{l}", torch.zeros(4 * H, dtype=inp.dtype))
            grads[f"bias_l{l}"] += full.sum(0)
            dG[t] = rnd(full, dgm)
            dG_next = dG[t]
        dGs = torch.stack(dG, dim=1).reshape(B * T, 4 * H)
        hprev = torch.stack(hs[:-1], dim=1).reshape(B * T, H)
        grads[f"weight_hh_l{l}"] = dGs.t() @ rnd(hprev, hm)
        grads[f"weight_ih_l{l}"] = dGs.t() @ rnd(inp.reshape(B * T, -1), hm)
        if l > 0:
            dx_above = []
            for t in range(T):
                prod = [rnd(dG[t][:, k * H:(k + 1) * H] @ wihT[k * H:(k + 1) * H], part) for k in range(4)]
                dx_above.append(prod[0] + prod[1] + prod[2] + prod[3])
    return grads


def main():
    N, M, T = 4, 5, 180
    sd = {k: v.double() for k, v in init_state_dict().items()}
    if len(sys.argv) > 1 and sys.argv[1] == "sat":
        sat = I.saturating_weights({k: v.numpy() for k, v in init_state_dict().items()})
        sd = {k: torch.tensor(v).double() for k, v in sat.items()}
    x = torch.tensor(I.logmel(N * M, T, seed=1234)).double()
    y, st = forward(x, sd)
    yn = y / y.norm(dim=1, keepdim=True)
    o = oge2e.ge2e_fwd_bwd(yn.numpy().reshape(N, M, -1), 10.0, -5.0)
    demb = torch.tensor(o["dE"]).reshape(N * M, -1)
    # through the L2 norm and the projection
    nrm = y.norm(dim=1, keepdim=True)
    dy = (demb - yn * (demb * yn).sum(1, keepdim=True)) / nrm
    dh_last = dy @ sd["projection.weight"]
    exact = backward(dh_last, st, sd, "exact", "exact", "exact", "exact", "exact")
    cfgs = {
        "now: stash bf16, dG bf16, W bf16, partials fp16, h bf16": ("bf16", "bf16", "bf16", "fp16", "bf16"),
        "stash fp16": ("fp16", "bf16", "bf16", "fp16", "bf16"),
        "stash exact": ("exact", "bf16", "bf16", "fp16", "bf16"),
        "stash fp16, dG exact": ("fp16", "exact", "bf16", "fp16", "bf16"),
        "stash fp16, W exact": ("fp16", "bf16", "exact", "fp16", "bf16"),
        "stash fp16, partials exact": ("fp16", "bf16", "bf16", "exact", "bf16"),
        "stash fp16, h exact": ("fp16", "bf16", "bf16", "fp16", "exact"),
        "stash fp16, c_t recomputed from the stash": ("fp16", "bf16", "bf16", "fp16", "bf16", 3, True),
        "only stash fp16 + c_t recomputed": ("fp16", "exact", "exact", "exact", "exact", 3, True),
        "only stash fp16": ("fp16", "exact", "exact", "exact", "exact"),
        "only stash bf16": ("bf16", "exact", "exact", "exact", "exact"),
        "only dG bf16": ("exact", "bf16", "exact", "exact", "exact"),
        "only W bf16": ("exact", "exact", "bf16", "exact", "exact"),
    }
    for name, c in cfgs.items():
        g = backward(dh_last, st, sd, *c)
        errs = {k: float((g[k] - exact[k]).norm() / exact[k].norm()) for k in sorted(g)}
        print(f"{name:58s} " + " ".join(f"{k.replace('weight_', 'w').replace('bias', 'b')}={v:.1e}" for k, v in errs.items()))


if __name__ == "__main__":
    main()
