"""Fused optimizer tail of the reference's training step.

``FusedClipSGD`` replaces, in two kernel launches, what train_speech_embedder.py:63-65 runs as stock torch code::

    torch.nn.utils.clip_grad_norm_(embedder_net.parameters(), 3.0)
    torch.nn.utils.clip_grad_norm_(ge2e_loss.parameters(), 1.0)
    optimizer.step()                      # torch.optim.SGD(..., lr=hp.train.lr), :33-36

It is opt-in: the unchanged script keeps working with torch.optim.SGD.  The arithmetic runs in
libsvb200.so (csrc/optim.cu); there is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from ._lib import check
from . import ops  # noqa: F401  (registers torch.ops.svb200)


class FusedClipSGD:
    """groups: list of dicts {'params': iterable of Parameters, 'max_norm': float or None} (the reference's two
    groups: embedder parameters clipped at 3.0, GE2E w/b at 1.0).  Same step()/zero_grad() surface as torch.optim."""

    def __init__(self, groups, lr, write_clipped_grads=True):
        self.param_groups = []
        for g in groups:
            params = [p for p in g["params"]]
            self.param_groups.append({"params": params, "max_norm": g.get("max_norm"), "lr": lr})
        n = sum(len(g["params"]) for g in self.param_groups)
        if n < 1 or n > 32 or len(self.param_groups) > 4:
            raise ValueError("FusedClipSGD handles 1..32 tensors in at most 4 clip groups")
        self.lr = float(lr)
        self.write_clipped_grads = bool(write_clipped_grads)
        self._ws = None
        self._norms = None

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.detach_().zero_()

    @torch.no_grad()
    def step(self):
        """Returns the pre-clip total norms, one per group (device tensor; what clip_grad_norm_ returns)."""
        params, grads, group, max_norm = [], [], [], []
        dev = None
        for gi, g in enumerate(self.param_groups):
            mn = g["max_norm"]
            max_norm.append(float(mn) if mn is not None else 0.0)
            for p in g["params"]:
                if p.grad is None:
                    continue
                if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                    raise _lib.SvbError("FusedClipSGD needs contiguous float32 CUDA parameters (no CPU fallback)")
                gr = p.grad
                if gr.dtype != torch.float32 or not gr.is_contiguous() or gr.device != p.device:
                    raise _lib.SvbError("FusedClipSGD needs contiguous float32 gradients on the parameter's device")
                dev = p.device if dev is None else dev
                if p.device != dev:
                    raise _lib.SvbError("FusedClipSGD: all parameters must live on one device")
                params.append(p)
                grads.append(gr)
                group.append(gi)
        if not params:
            return None
        with torch.cuda.device(dev):
            if self._ws is None or self._ws.device != dev:
                nb = ctypes.c_size_t(0)
                check(_lib.lib().svb_clip_sgd_workspace_bytes(ctypes.byref(nb)), "svb_clip_sgd_workspace_bytes")
                self._ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            self._norms = torch.ops.svb200.clip_sgd(params, grads, group, max_norm, self.lr, self.write_clipped_grads,
                                                    self._ws)
        for p in params:                      # the packed fp16/bf16 weight shadows key on Parameter._version
            torch.autograd.graph.increment_version(p)
        return self._norms
