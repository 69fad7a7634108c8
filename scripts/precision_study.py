"""CPU emulation: which operands of the LSTM may be plain bf16 within the 1e-3 embedding tolerance?
Per-layer rounding modes for the input projection (x, W_ih) and the recurrence (h, W_hh)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import _inputs as I
from oracle.embedder import init_state_dict, _rnd, embedder_explicit

def run(x, sd, modes, scale=1.0):
    inp = x
    for l in range(3):
        w_ih = sd[f"LSTM_stack.weight_ih_l{l}"] * scale
        w_hh = sd[f"LSTM_stack.weight_hh_l{l}"] * scale
        bias = sd[f"LSTM_stack.bias_ih_l{l}"] + sd[f"LSTM_stack.bias_hh_l{l}"]
        xi, wi, hr, wr = modes[l]
        H = w_hh.shape[1]
        B, T, _ = inp.shape
        gin = _rnd(inp, xi) @ _rnd(w_ih, wi).t() + bias
        whr = _rnd(w_hh, wr).t()
        h = x.new_zeros(B, H); c = x.new_zeros(B, H); hs = []
        for t in range(T):
            g = gin[:, t] + _rnd(h, hr) @ whr
            i, f, gg, o = g.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            hs.append(h)
        inp = torch.stack(hs, dim=1)
    y = inp[:, -1] @ sd["projection.weight"].t() + sd["projection.bias"]
    return y / y.norm(dim=1, keepdim=True)

torch.set_num_threads(8)
sd = init_state_dict()
B, T = 16, 160
x = torch.tensor(I.logmel(B, T, seed=5))
for scale in (1.0, 2.5):
    ref = run(x.double(), {k: v.double() for k, v in sd.items()}, [("fp32",) * 4] * 3, scale).float()
    cfgs = {
      "now: in split x3 all layers, rec bf16": [("split", "split", "bf16", "bf16")] * 3,
      "l0 split, l1-2 in bf16/bf16": [("split", "split", "bf16", "bf16")] + [("bf16", "bf16", "bf16", "bf16")] * 2,
      "l0 split, l1-2 in split-x/bf16-w": [("split", "split", "bf16", "bf16")] + [("split", "bf16", "bf16", "bf16")] * 2,
      "l0 split, l1-2 in bf16-x/split-w": [("split", "split", "bf16", "bf16")] + [("bf16", "split", "bf16", "bf16")] * 2,
      "all bf16": [("bf16", "bf16", "bf16", "bf16")] * 3,
    }
    for name, m in cfgs.items():
        e = run(x, sd, m, scale)
        err = ((e - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
        print(f"scale {scale}: {name:45s} err {err:.2e}")
