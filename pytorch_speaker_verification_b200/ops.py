"""Host-side operators over the C ABI (include/svb200.h): thin, allocation + argument marshalling only.

torch is used for device memory, streams and autograd bookkeeping; every arithmetic step runs in
libsvb200.so.  CPU tensors are accepted at the boundary (the reference's test() and dvector_create.py
call the model with CPU tensors): they are staged to the GPU and results return on the caller's device.
"""
import ctypes

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

import os

_POISON = os.environ.get("SVB_POISON_WORKSPACE", "0") == "1"
_i64 = ctypes.c_int64
_sz = ctypes.c_size_t


def _dev():
    if not torch.cuda.is_available():
        raise _lib.SvbError("a CUDA device (B200, sm_100a) is required: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stage(t, dtype=None):
    """-> contiguous CUDA tensor (optionally cast)."""
    d = _dev()
    if t.device.type != "cuda":
        t = t.to(d, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def set_poison_workspace(on):
    """Debug: NaN-fill every forward workspace, so that a kernel reading a slot before it is written shows up."""
    global _POISON
    _POISON = bool(on)


def set_persistent(on):
    """Select the LSTM forward path: True (default) persistent wavefront kernel (csrc/wlstm.cuh), False per-frame
    launches (csrc/lstm.cu)."""
    check(_lib.lib().svb_set_persistent(int(bool(on))), "svb_set_persistent")


def set_persistent_bwd(on):
    """Select the BPTT path: True (default) persistent wavefront kernel (csrc/wbptt.cuh), False per-frame launches."""
    check(_lib.lib().svb_set_persistent_bwd(int(bool(on))), "svb_set_persistent_bwd")


def set_wgrad_overlap(on):
    """True (default): the weight-gradient products over the late frames run on a second stream beside the persistent
    BPTT kernel (which leaves 28 SMs idle); False: all of them after it."""
    check(_lib.lib().svb_set_wgrad_overlap(int(bool(on))), "svb_set_wgrad_overlap")


_GRAD_CB = ctypes.CFUNCTYPE(None, ctypes.c_int, ctypes.c_void_p)
_bucket_hook = None


def set_grad_bucket_hook(fn):
    """fn(bucket: 1-D view of the flat gradient buffer) is called during backward as soon as a bucket's kernels are
    enqueued on the current stream (projection first, then the LSTM layers from the top); None clears it."""
    global _bucket_hook
    _bucket_hook = fn


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


# ------------------------------------------------------------------------------------------ embedder
class PackedWeights:
    """bf16 gate-interleaved shadow of the fp32 master parameters, refreshed when they change."""

    def __init__(self):
        self.key = None
        self.buf = None
        self.dev_params = None

    def get(self, params, I, H, L):
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params)
        if key != self.key:
            dev_params = [_stage(p.detach(), torch.float32) for p in params]
            nbytes = _sz(0)
            check(_lib.lib().svb_embedder_sizes(1, 1, I, H, L, 1, 0, ctypes.byref(nbytes), None), "svb_embedder_sizes")
            if self.buf is None or self.buf.numel() != nbytes.value or self.buf.device != dev_params[0].device:
                self.buf = torch.empty(nbytes.value, dtype=torch.uint8, device=dev_params[0].device)
            check(_lib.lib().svb_embedder_pack_weights(_ptr_array(dev_params[:4 * L]), ptr(self.buf), I, H, L,
                                                       stream_ptr()), "svb_embedder_pack_weights")
            self.key = key
            self.dev_params = dev_params
        return self.buf, self.dev_params


class EmbedderFn(torch.autograd.Function):
    """SpeechEmbedder.forward / BPTT (speech_embedder_net.py:27-33; train_speech_embedder.py:62)."""

    @staticmethod
    def forward(ctx, x, cache, dims, *params):
        I, H, L, P, rec_terms = dims
        if x.dim() != 3 or x.shape[2] != I:
            raise ValueError(f"expected input (batch, frames, {I}), got {tuple(x.shape)}")
        out_device = x.device
        B, T = int(x.shape[0]), int(x.shape[1])
        with torch.cuda.device(_dev()):
            if x.dtype not in (torch.float32, torch.float64):
                x = x.float()                                          # speech_embedder_net.py:28
            xg = _stage(x)
            packed, dev_params = cache.get(params, I, H, L)
            training = bool(any(ctx.needs_input_grad[3:]))
            wbytes = _sz(0)
            check(_lib.lib().svb_embedder_sizes(B, T, I, H, L, P, int(training), None, ctypes.byref(wbytes)),
                  "svb_embedder_sizes")
            ws = torch.empty(wbytes.value, dtype=torch.uint8, device=xg.device)
            if _POISON:          # debug: NaN-fill the workspace so that any read of a not-yet-written slot shows
                ws.fill_(255)
            emb = torch.empty(B, P, dtype=torch.float32, device=xg.device)
            check(_lib.lib().svb_embedder_forward(ptr(xg), 0 if xg.dtype == torch.float32 else 1, ptr(packed),
                                                  ptr(dev_params[4 * L]), ptr(dev_params[4 * L + 1]), ptr(emb), ptr(ws),
                                                  B, T, I, H, L, P, int(training), int(rec_terms), stream_ptr()),
                  "svb_embedder_forward")
        if training:
            ctx.ws, ctx.packed, ctx.dev_params = ws, packed, dev_params
            ctx.shape = (B, T, I, H, L, P)
            ctx.param_devices = [p.device for p in params]
        return emb if out_device.type == "cuda" else emb.to(out_device)

    @staticmethod
    def backward(ctx, demb):
        B, T, I, H, L, P = ctx.shape
        with torch.cuda.device(ctx.ws.device):
            dg = _stage(demb, torch.float32)
            # one flat buffer, parameters as views: a single all-reduce covers every gradient (dist.py)
            flat = torch.empty(sum(p.numel() for p in ctx.dev_params), dtype=torch.float32, device=dg.device)
            grads, off, offs = [], 0, [0]
            for p in ctx.dev_params:
                grads.append(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
                offs.append(off)
            cb, errors = None, []
            if _bucket_hook is not None:
                hook = _bucket_hook

                def ready(bucket, _user):            # bucket L: projection, L-1 ... 0: LSTM layers (svb200.h)
                    try:
                        lo, hi = (offs[4 * L], offs[4 * L + 2]) if bucket == L else (offs[4 * bucket], offs[4 * bucket + 4])
                        hook(flat[lo:hi])
                    except BaseException as exc:         # exceptions cannot cross the C frame: re-raised below
                        errors.append(exc)

                cb = _GRAD_CB(ready)
                check(_lib.lib().svb_set_grad_ready_callback(cb, None), "svb_set_grad_ready_callback")
            try:
                check(_lib.lib().svb_embedder_backward(ptr(dg), ptr(ctx.packed), ptr(ctx.dev_params[4 * L]),
                                                       _ptr_array(grads), ptr(ctx.ws), B, T, I, H, L, P, stream_ptr()),
                      "svb_embedder_backward")
            finally:
                if cb is not None:
                    _lib.lib().svb_set_grad_ready_callback(None, None)
            if errors:
                raise errors[0]
        ctx.ws = None
        grads = [g if d.type == "cuda" else g.to(d) for g, d in zip(grads, ctx.param_devices)]
        return (None, None, None, *grads)


# ------------------------------------------------------------------------------------------ GE2E
def _ge2e_call(E, Cext, w, b, dcos, gscale, want_cos, want_loss, need_grad, fused=True):
    N, M, D = (int(s) for s in E.shape)
    Nc = N if Cext is None else int(Cext.shape[0])
    dev = E.device
    nbytes = _sz(0)
    check(_lib.lib().svb_ge2e_workspace_bytes(N, M, D, Nc, ctypes.byref(nbytes)), "svb_ge2e_workspace_bytes")
    ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    cos = torch.empty(N, M, Nc, **f32) if want_cos else None
    per = torch.empty(N, M, **f32) if want_loss else None
    loss = torch.empty((), **f32) if want_loss else None
    dE = torch.empty(N, M, D, **f32) if need_grad else None
    dC = torch.empty(Nc, D, **f32) if (need_grad and Cext is not None) else None
    dw = torch.empty((), **f32) if (need_grad and w is not None) else None
    db = torch.empty((), **f32) if (need_grad and w is not None) else None
    check(_lib.lib().svb_ge2e(ptr(E), ptr(Cext), N, M, D, Nc, ptr(w), ptr(b), ptr(dcos), ptr(gscale), ptr(cos),
                              ptr(per), ptr(loss), ptr(dE), ptr(dC), ptr(dw), ptr(db), ptr(ws), _sz(nbytes.value),
                              int(fused), stream_ptr()), "svb_ge2e")
    return dict(cos=cos, per=per, loss=loss, dE=dE, dC=dC, dw=dw, db=db)


class GE2ELossFn(torch.autograd.Function):
    """GE2ELoss.forward + backward in one fused kernel (speech_embedder_net.py:43-49, utils.py:27-132)."""

    @staticmethod
    def forward(ctx, E, w, b, fused):
        if E.dim() != 3:
            raise ValueError("embeddings must be (speakers, utterances, dim)")
        if E.shape[1] < 2:
            raise ValueError("GE2E needs at least 2 utterances per speaker (leave-one-out centroid, utils.py:56-57)")
        out_device = E.device
        with torch.cuda.device(_dev()):
            Eg, wg, bg = _stage(E, torch.float32), _stage(w, torch.float32), _stage(b, torch.float32)
            need = bool(any(ctx.needs_input_grad[:3]))
            r = _ge2e_call(Eg, None, wg, bg, None, None, False, True, need, fused)
        if need:
            ctx.saved = (r["dE"], r["dw"], r["db"])
            ctx.devs = (E.device, w.device, b.device)
        return r["loss"] if out_device.type == "cuda" else r["loss"].to(out_device)

    @staticmethod
    def backward(ctx, g):
        dE, dw, db = (t.clone() for t in ctx.saved)
        with torch.cuda.device(dE.device):
            gg = _stage(g, torch.float32)
            check(_lib.lib().svb_scale3(ptr(dE), _sz(dE.numel()), ptr(dw), _sz(1), ptr(db), _sz(1), ptr(gg),
                                        stream_ptr()), "svb_scale3")
        outs = [t if d.type == "cuda" else t.to(d) for t, d in zip((dE, dw, db), ctx.devs)]
        return outs[0], outs[1], outs[2], None


class CentroidsFn(torch.autograd.Function):
    """utils.get_centroids (utils.py:27-29)."""

    @staticmethod
    def forward(ctx, E):
        out_device = E.device
        N, M, D = (int(s) for s in E.shape)
        with torch.cuda.device(_dev()):
            Eg = _stage(E, torch.float32)
            C = torch.empty(N, D, dtype=torch.float32, device=Eg.device)
            check(_lib.lib().svb_centroids(ptr(Eg), ptr(C), N, M, D, stream_ptr()), "svb_centroids")
        ctx.shape, ctx.dev = (N, M, D), out_device
        return C if out_device.type == "cuda" else C.to(out_device)

    @staticmethod
    def backward(ctx, dC):
        N, M, D = ctx.shape
        with torch.cuda.device(_dev()):
            dCg = _stage(dC, torch.float32)
            dE = torch.empty(N, M, D, dtype=torch.float32, device=dCg.device)
            check(_lib.lib().svb_centroids_bwd(ptr(dCg), ptr(dE), N, M, D, stream_ptr()), "svb_centroids_bwd")
        return dE if ctx.dev.type == "cuda" else dE.to(ctx.dev)


class CossimFn(torch.autograd.Function):
    """utils.get_cossim (utils.py:72-115), differentiable w.r.t. embeddings and centroids."""

    @staticmethod
    def forward(ctx, E, C):
        if E.dim() != 3 or C.dim() != 2 or C.shape[1] != E.shape[2]:
            raise ValueError("get_cossim expects embeddings (N,M,D) and centroids (N',D)")
        if C.shape[0] != E.shape[0]:
            # the reference's diagonal index_put (utils.py:112-113) requires N' >= N; it is only ever called with N' == N
            raise ValueError("get_cossim: centroids must have one row per speaker of embeddings")
        out_device = E.device
        with torch.cuda.device(_dev()):
            Eg, Cg = _stage(E, torch.float32), _stage(C, torch.float32)
            r = _ge2e_call(Eg, Cg, None, None, None, None, True, False, False)
        ctx.save_for_backward(Eg, Cg)
        ctx.devs = (E.device, C.device)
        return r["cos"] if out_device.type == "cuda" else r["cos"].to(out_device)

    @staticmethod
    def backward(ctx, dcos):
        Eg, Cg = ctx.saved_tensors
        with torch.cuda.device(Eg.device):
            r = _ge2e_call(Eg, Cg, None, None, _stage(dcos, torch.float32), None, False, False, True)
        outs = [t if d.type == "cuda" else t.to(d) for t, d in zip((r["dE"], r["dC"]), ctx.devs)]
        return outs[0], outs[1]


class CalcLossFn(torch.autograd.Function):
    """utils.calc_loss (utils.py:126-132): (loss, per_embedding_loss)."""

    @staticmethod
    def forward(ctx, S):
        if S.dim() != 3 or S.shape[2] < S.shape[0]:
            raise ValueError("calc_loss expects a similarity matrix (N, M, N)")
        out_device = S.device
        N, M, Nc = (int(s) for s in S.shape)
        with torch.cuda.device(_dev()):
            Sg = _stage(S, torch.float32)
            per = torch.empty(N, M, dtype=torch.float32, device=Sg.device)
            loss = torch.empty((), dtype=torch.float32, device=Sg.device)
            check(_lib.lib().svb_calc_loss(ptr(Sg), N, M, Nc, ptr(per), ptr(loss), None, None, stream_ptr()),
                  "svb_calc_loss")
        ctx.save_for_backward(Sg)
        ctx.dev = out_device
        ctx.mark_non_differentiable(per)
        if out_device.type != "cuda":
            loss, per = loss.to(out_device), per.to(out_device)
        return loss, per

    @staticmethod
    def backward(ctx, gloss, _gper):
        (Sg,) = ctx.saved_tensors
        N, M, Nc = (int(s) for s in Sg.shape)
        with torch.cuda.device(Sg.device):
            per = torch.empty(N, M, dtype=torch.float32, device=Sg.device)
            dS = torch.empty_like(Sg)
            check(_lib.lib().svb_calc_loss(ptr(Sg), N, M, Nc, ptr(per), None, ptr(dS),
                                           ptr(_stage(gloss, torch.float32)), stream_ptr()), "svb_calc_loss")
        return dS if ctx.dev.type == "cuda" else dS.to(ctx.dev)


# ------------------------------------------------------------------------------------------ EER
def eer_counts(sim, thresholds_f32, speaker0=0):
    """Exact counts of sim > t for the row blocks of speakers [speaker0, speaker0 + sim.shape[0])."""
    n, Mv, Nc = (int(s) for s in sim.shape)
    T = int(thresholds_f32.numel())
    ca = torch.empty(n, T, dtype=torch.int32, device=sim.device)
    cd = torch.empty(n, T, dtype=torch.int32, device=sim.device)
    check(_lib.lib().svb_eer_counts(ptr(sim), n, Mv, Nc, int(speaker0), ptr(thresholds_f32), T, ptr(ca), ptr(cd),
                                    stream_ptr()), "svb_eer_counts")
    return ca, cd


def eer_sweep_fused(sim, thresholds_f32):
    """One launch: (out[4 + 2T], cnt_all, cnt_diag); out[1] == -2 asks for the sequential eer_finish."""
    N, Mv, Nc = (int(s) for s in sim.shape)
    T = int(thresholds_f32.numel())
    ca = torch.empty(N, T, dtype=torch.int32, device=sim.device)
    cd = torch.empty(N, T, dtype=torch.int32, device=sim.device)
    scratch = torch.zeros(1 + 16 * T, dtype=torch.int64, device=sim.device)
    out = torch.empty(4 + 2 * T, dtype=torch.float32, device=sim.device)
    check(_lib.lib().svb_eer_sweep(ptr(sim), N, Mv, ptr(thresholds_f32), T, ptr(ca), ptr(cd), ptr(scratch), ptr(out),
                                   stream_ptr()), "svb_eer_sweep")
    return out, ca, cd


def eer_finish(cnt_all, cnt_diag, Mv):
    N, T = (int(s) for s in cnt_all.shape)
    out = torch.empty(4 + 2 * T, dtype=torch.float32, device=cnt_all.device)
    check(_lib.lib().svb_eer_finish(ptr(cnt_all), ptr(cnt_diag), N, int(Mv), T, ptr(out), stream_ptr()),
          "svb_eer_finish")
    return out


# ------------------------------------------------------------------------------------------ d-vectors
def dvector_windows(S, win_start, win):
    """S (nmels, Ttot) float32 CUDA, win_start int32 CUDA (W,) -> (W, win, nmels)."""
    nmels, W = int(S.shape[0]), int(win_start.numel())
    out = torch.empty(W, win, nmels, dtype=torch.float32, device=S.device)
    check(_lib.lib().svb_dvector_windows(ptr(S), _i64(S.stride(0)), nmels, ptr(win_start), W, int(win), ptr(out),
                                         stream_ptr()), "svb_dvector_windows")
    return out


def segment_mean(emb, seg_offsets):
    P, D = int(seg_offsets.numel()) - 1, int(emb.shape[1])
    out = torch.empty(P, D, dtype=torch.float64, device=emb.device)
    check(_lib.lib().svb_segment_mean(ptr(emb), D, ptr(seg_offsets), P, ptr(out), stream_ptr()), "svb_segment_mean")
    return out
