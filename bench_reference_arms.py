"""Reference arms of bench.py: the reference's own code paths timed beside ours (never on our path).

* ``TorchCudaReference``: the UNMODIFIED reference modules (baseline/_ref: speech_embedder_net.SpeechEmbedder /
  GE2ELoss, utils.get_centroids / get_cossim; staged by __graft_entry__.build()) on the same B200 through stock
  torch CUDA -- nn.LSTM -> cuDNN (speech_embedder_net.py:28), eager ATen GE2E (utils.py:72-132), clip_grad_norm_ x2
  and SGD (train_speech_embedder.py:54-65).  This is the kernel-vs-kernel bar of SURVEY.md section 2a / 8d.  When
  baseline/_ref is absent the oracle's port of the same library calls is used and the block says kind = "port".
* ``cpu_secondary``: the reference's CPU path for the secondary rows (GE2E alone, get_cossim + EER loop, per-file
  d-vector extraction), bounded samples, host cores stated.
"""
import contextlib
import os
import statistics
import sys
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "baseline", "_ref")


def _import_reference():
    """-> (speech_embedder_net, utils) of the staged reference, or None.  Same two shims as tests/_run_caller.py."""
    need = ["hparam.py", "utils.py", "speech_embedder_net.py", os.path.join("config", "config.yaml")]
    if not all(os.path.isfile(os.path.join(REF, f)) for f in need):
        return None
    import yaml
    orig = yaml.load_all
    yaml.load_all = lambda stream, Loader=None: orig(stream, Loader=Loader or yaml.FullLoader)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))
    cwd = os.getcwd()
    saved = {k: sys.modules.get(k) for k in ("hparam", "utils", "speech_embedder_net")}
    try:
        os.chdir(REF)                              # hparam.py:49 opens the CWD-relative config/config.yaml
        sys.path.insert(0, REF)
        for k in saved:
            sys.modules.pop(k, None)
        import speech_embedder_net as ref_net
        import utils as ref_utils
        return ref_net, ref_utils
    except Exception as exc:                       # pragma: no cover - reported by the caller
        print(f"reference import failed: {exc!r}", file=sys.stderr)
        return None
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
        for k, v in saved.items():                 # leave no reference module under a generic name behind
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
        yaml.load_all = orig


_eer_code = None


def eer_loop(sim_matrix, N, M):
    """The threshold sweep of train_speech_embedder.py:132-149, executed FROM THE STAGED REFERENCE SOURCE: the
    statements of ``test()`` from ``diff = 1; ...`` through the ``for thres`` loop are cut out of
    baseline/_ref/train_speech_embedder.py with ``ast`` (the script body cannot be imported without running it) and run
    with ``hp.test.N`` / ``hp.test.M`` bound to the benchmark's sizes.  -> (EER, thresh, FAR, FRR) or None."""
    global _eer_code
    path = os.path.join(REF, "train_speech_embedder.py")
    if not os.path.isfile(path):
        return None
    if _eer_code is None:
        import ast
        tree = ast.parse(open(path).read())
        fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "test")
        stmts = None
        for node in ast.walk(fn):
            body = getattr(node, "body", None)
            if isinstance(body, list):
                for i, st in enumerate(body):
                    if (isinstance(st, ast.For) and isinstance(st.target, ast.Name) and st.target.id == "thres"):
                        j = i               # the initialisers share one source line (:132):
                        while (j > 0 and isinstance(body[j - 1], ast.Assign)       # diff = 1; EER=0; EER_thresh = 0; ...
                               and body[j - 1].lineno == body[i - 1].lineno):
                            j -= 1
                        stmts = body[j:i + 1]
        if stmts is None:
            return None
        _eer_code = compile(ast.Module(body=stmts, type_ignores=[]), path, "exec")
    hp = types.SimpleNamespace(test=types.SimpleNamespace(N=N, M=M))
    ns = {"sim_matrix": sim_matrix, "hp": hp}
    exec(_eer_code, ns)
    return ns["EER"], ns["EER_thresh"], ns["EER_FAR"], ns["EER_FRR"]


class TorchCudaReference:
    def __init__(self, torch, dev):
        self.torch, self.dev = torch, dev
        mods = _import_reference()
        if mods is not None:
            self.kind = "reference"
            self.net_mod, self.utils = mods
            self.make_net = lambda: self.net_mod.SpeechEmbedder()
            self.make_loss = lambda d: self.net_mod.GE2ELoss(d)
            self.get_centroids, self.get_cossim = self.utils.get_centroids, self.utils.get_cossim
        else:
            sys.path.insert(0, ROOT)
            from oracle import embedder as oemb        # the port of the same library calls (bench's reference arm only)
            import torch.nn as nn
            self.kind = "port"

            class Loss(nn.Module):
                def __init__(self, d):
                    super().__init__()
                    self.w = nn.Parameter(torch.tensor(10.0).to(d))
                    self.b = nn.Parameter(torch.tensor(-5.0).to(d))

                def forward(self, E):
                    return oemb.library_ge2e_loss(E, self.w, self.b)

            self.make_net = lambda: oemb.LibraryEmbedder()
            self.make_loss = lambda d: Loss(d)
            self.get_centroids = lambda E: E.mean(dim=1)
            self.get_cossim = None

    def _time(self, fn, steps, warmup, flush=None):
        torch = self.torch
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            if flush is not None:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return statistics.median(ms)

    def train_step(self, x_host, N, M, steps, warmup, autocast_bf16=False, flush=None):
        """train_speech_embedder.py:45-65 on the device: returns (ms per full step from pinned host memory incl. H2D,
        clip x2, SGD, D2H loss; ms per fwd+loss+bwd with the batch resident)."""
        torch, dev = self.torch, self.dev
        torch.manual_seed(0)
        net = self.make_net().to(dev)
        crit = self.make_loss(dev)
        opt = torch.optim.SGD([{'params': net.parameters()}, {'params': crit.parameters()}], lr=0.01)
        x_dev = x_host.to(dev)
        loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if autocast_bf16 else contextlib.nullcontext

        def fwd_bwd(x):
            with ctx():
                emb = net(x)
            loss = crit(emb.float().reshape(N, M, -1))
            loss.backward()
            return loss

        def value_step():
            for p in list(net.parameters()) + list(crit.parameters()):
                p.grad = None
            fwd_bwd(x_dev)

        def full_step():
            x = x_host.to(dev, non_blocking=True)
            opt.zero_grad()
            loss = fwd_bwd(x)
            torch.nn.utils.clip_grad_norm_(net.parameters(), 3.0)
            torch.nn.utils.clip_grad_norm_(crit.parameters(), 1.0)
            opt.step()
            loss_host.copy_(loss.detach(), non_blocking=True)

        ms_value = self._time(value_step, steps, warmup, flush)
        ms_full = self._time(full_step, steps, warmup, flush)
        return ms_full, ms_value

    def ge2e_only(self, Enp, steps=10, warmup=3):
        torch, dev = self.torch, self.dev
        crit = self.make_loss(dev)
        E = torch.tensor(Enp, device=dev, requires_grad=True)

        def step():
            E.grad = None
            crit.w.grad = None
            crit.b.grad = None
            crit(E).backward()

        return self._time(step, steps, warmup)

    def eer(self, enr, ver):
        """train_speech_embedder.py:127-149 on the device: (ms get_centroids + get_cossim, ms threshold loop)."""
        torch, dev = self.torch, self.dev
        if self.get_cossim is None:
            return None
        enr, ver = torch.tensor(enr, device=dev), torch.tensor(ver, device=dev)
        N, Mh = int(enr.shape[0]), int(enr.shape[1])
        with torch.no_grad():
            ms_cos = self._time(lambda: self.get_cossim(ver, self.get_centroids(enr)), 3, 1)
            sim = self.get_cossim(ver, self.get_centroids(enr))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tup = eer_loop(sim, N, 2 * Mh)
            torch.cuda.synchronize()
            ms_loop = (time.perf_counter() - t0) * 1e3
        if tup is None:
            return None
        return ms_cos, ms_loop, float(tup[0])

    def forward_windows(self, xw, steps=3, warmup=1):
        """dvector_create.py:100 as one batch on the device (the reference runs it per file on the CPU)."""
        torch = self.torch
        torch.manual_seed(0)
        net = self.make_net().to(self.dev).eval()
        with torch.no_grad():
            return self._time(lambda: net(xw), steps, warmup)


def cpu_secondary(torch, I, budget_s=25.0):
    """Reference CPU path of the secondary rows on all host cores (bounded samples)."""
    sys.path.insert(0, ROOT)
    from oracle import embedder as oemb
    import numpy as np
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {"cores": cores}
    # GE2E fwd+bwd alone at N=64 x M=10 (utils.py:27-132 through autograd)
    E = torch.tensor(I.ge2e_embeddings(64, 10, 256, "unit"), requires_grad=True)
    w = torch.tensor(10.0, requires_grad=True)
    b = torch.tensor(-5.0, requires_grad=True)
    ts = []
    for _ in range(4):
        E.grad = None
        t0 = time.perf_counter()
        oemb.library_ge2e_loss(E, w, b).backward()
        ts.append(time.perf_counter() - t0)
    out["ge2e_fwd_bwd_N64_ms"] = statistics.median(ts[1:]) * 1e3
    # get_cossim + EER loop at N=1024, M=6 (train_speech_embedder.py:127-149): one pass (about 5 s)
    mods = _import_reference()
    if mods is not None:
        _, ref_utils = mods
        enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
        enr, ver = torch.tensor(enr), torch.tensor(ver)
        with torch.no_grad():
            t0 = time.perf_counter()
            sim = ref_utils.get_cossim(ver, ref_utils.get_centroids(enr))
            t1 = time.perf_counter()
            eer_loop(sim, 1024, 6)
            t2 = time.perf_counter()
        out["eer_N1024"] = {"cossim_ms": (t1 - t0) * 1e3, "threshold_loop_ms": (t2 - t1) * 1e3, "kind": "reference"}
    # per-file d-vector extraction (dvector_create.py:98-101): windows -> forward -> align, one utterance at a time
    from oracle import dvector as odv
    torch.manual_seed(0)
    net = oemb.LibraryEmbedder().eval()
    r = np.random.RandomState(4321)
    Ts = r.randint(100, 501, size=4000)
    n_utt, n_win, t_used = 0, 0, 0.0
    with torch.no_grad():
        for T in Ts:
            S = np.log10(I.power_spec(int(T), seed=int(T)) + 1e-6).astype(np.float32)
            t0 = time.perf_counter()
            fr = odv.windows(S)
            if len(fr):
                emb = net(torch.tensor(fr))
                odv.align_embeddings(emb.numpy())
            t_used += time.perf_counter() - t0
            n_utt += 1
            n_win += len(fr)
            if t_used > budget_s * 0.5:
                break
    out["extraction"] = {"utterances": n_utt, "windows": n_win, "seconds": t_used, "windows_per_s": n_win / t_used,
                         "utterances_per_s": n_utt / t_used,
                         "sample": f"first {n_utt} utterances of the synthetic set, per-file loop of dvector_create.py:98-101"}
    return out
