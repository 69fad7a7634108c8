#!/usr/bin/env python3
"""Benchmark of the GE2E training hot path (BASELINE.json metric: train utts/sec, LSTM+GE2E fwd+bwd, N=64 M=10).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-cuda]

One JSON line on stdout (rank 0).  Workload = BASELINE.json configs[1]: 64 speakers x 10 utterances x 160 frames x
40 mel per GPU, synthetic log-mel, reference-initialised weights.  Under torchrun (N > 1) every rank owns 64
speakers (weak scaling), d-vectors are all-gathered for the global GE2E batch (configs[2] at N=8) and parameter
gradients are all-reduced (SUM).

  value : utterances/s of fwd + GE2E + bwd with the batch already resident in HBM
  e2e   : the caller's whole step through the public API -- H2D of the pinned host batch, zero_grad, forward, GE2E,
          backward, the two clip_grad_norm_ and the SGD step of train_speech_embedder.py:54-65, D2H of the loss
  --impl reference : the reference's CPU path (oracle port calling the same torch CPU library entry points the
          reference calls), all host threads, on a bounded sample of the same workload.
  --impl reference-cuda : the UNMODIFIED reference modules (baseline/_ref) on the same GPU through stock torch CUDA
          (cuDNN nn.LSTM, eager GE2E, clip_grad_norm_ x2, SGD): the kernel-vs-kernel bar.  The default run also
          carries these numbers as `gpu_baseline` (N = 1) next to `cpu_baseline`.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_SPK, M_UTT, T_FR, NMELS, HID, NLAYER, PROJ = 64, 10, 160, 40, 768, 3, 256
METRIC = "train utts/sec (LSTM+GE2E fwd+bwd, N=64 M=10)"
PHASES = ["prep", "input_gemm", "recurrent_fwd", "projection", "projection_bwd", "recurrent_bwd", "weight_grads",
          "bias_grads", "dx"]


_RESULT_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region.  NVML from a thread of this process (pynvml, the
    library nvidia-smi itself uses): a looping nvidia-smi subprocess took ~200 ms per query on some boxes and stalled
    kernel launches for tens of ms inside a timed loop.  Falls back to the nvidia-smi loop of the profiling recipe
    (-lms 200) if pynvml is unavailable."""

    def __init__(self, index, uuid=None, period=0.05):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.p, self.thread, self.stop_flag = None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid is not None:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}

            def loop():
                while not self.stop_flag:
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for bit, n in names.items():
                            if r & bit:
                                self.reasons.add(n)
                    except Exception:
                        pass
                    time.sleep(period)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        self.source = "nvidia-smi -lms 200"
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = self.samples
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(self.reasons), "source": self.source}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": self.source}


def launches_per_step(world):
    """Fallback when CUPTI is unavailable: our kernels per `value` step as listed by ncu (profiles/*launches_step*)."""
    fwd = 1 + 1 + 2                                    # prep_x, persistent wavefront LSTM kernel, projection GEMM + finish
    loss = 1                                           # GE2E (per-speaker kernel; general kernel for the global batch)
    bwd = (1 + 5 + 1 + 1                               # scale3, projection bwd (norm, 2 GEMMs, add2, colsum), gradient scale, persistent BPTT
           + NLAYER                                    # frame gates of the late weight-gradient slices
           + 2 * (2 * NLAYER - 1)                      # early + late slices of the 5 wide weight-gradient products
           + 2                                         # layer-0 dW_ih (N = 40) + its slice sum
           + (2 * NLAYER - 1))                         # sum of the three partials per wide product
    return fwd + loss + bwd


def count_launches(L, fn, torch):
    """-> (our kernel launches, other launches, {kernel: count}) of one call of fn, seen through CUPTI's callback API
    (svb_launch_count_*: a launch is ours when its host stub lives in libsvb200.so), or None without CUPTI."""
    if L.svb_launch_count_begin() != 0:
        return None
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        ours, other = ctypes.c_longlong(0), ctypes.c_longlong(0)
        names = ctypes.create_string_buffer(1 << 16)
        L.svb_launch_count_end(ctypes.byref(ours), ctypes.byref(other), names, ctypes.c_size_t(len(names)))
    per = {}
    for item in names.value.decode().split(";"):
        if "=" in item:
            k, v = item.rsplit("=", 1)
            per[k] = int(v)
    return ours.value, other.value, per


def measured_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed summary
    of this round's `ncu --set full` capture (profiles/r2_traffic.json, written by scripts/summarize_ncu.py); None
    when that capture does not exist -- never a number carried over from another build."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        e = t.get(kernel_key)
        return None if e is None else int(e["dram_bytes_read"] + e["dram_bytes_write"])
    except Exception:
        return None


def cpu_reference_rate(steps, warmup, budget_s=150.0):
    """Reference CPU path on all host threads over a bounded sample; returns (utts/s, description, cores)."""
    import torch
    import _inputs as I
    from oracle.embedder import ReferenceLibraryStep
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # probe with 1 speaker-group of 4 x 10 utterances to size the sample
    probe = ReferenceLibraryStep(4, M_UTT)
    xb = torch.tensor(I.logmel(4 * M_UTT, T_FR, seed=1234))
    probe.step(xb)
    t0 = time.perf_counter()
    probe.step(xb)
    per_utt = (time.perf_counter() - t0) / (4 * M_UTT)
    n = N_SPK
    while n > 4 and per_utt * n * M_UTT * (steps + warmup) > budget_s:
        n //= 2
    ref = ReferenceLibraryStep(n, M_UTT)
    xb = torch.tensor(I.logmel(n * M_UTT, T_FR, seed=1234))
    for _ in range(warmup):
        ref.step(xb)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.step(xb)
        ts.append(time.perf_counter() - t0)
    dt = statistics.median(ts)
    sample = (f"{n} speakers x {M_UTT} utts x {T_FR} frames per step (of the 64 x 10 workload), full train step "
              f"(fwd+GE2E+bwd+clip+SGD, train_speech_embedder.py:54-65), {steps} steps after {warmup} warm-up, median")
    return n * M_UTT / dt, sample, cores, dt


def secondary_benchmarks(torch, dist, svb, _lib, I, net, dev, rank, world, n_extract=12500):
    """The other rows of SURVEY section 8: d-vector extraction (configs[3] shape, scaled down to a bounded shard per
    GPU), the EER sweep at configs[4] size and the fused GE2E kernel alone; device time via CUDA events."""
    import ctypes
    import numpy as np
    from pytorch_speaker_verification_b200 import eer as E, ops
    from pytorch_speaker_verification_b200._lib import ptr, stream_ptr
    L = _lib.lib()
    out = {}

    def dev_time(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- d-vector extraction, BASELINE configs[3]: 100k utterances over 8 GPUs = 12,500 per GPU (every rank takes its
    # own 12,500: weak scaling), T_u ~ U[100, 500] frames, from host log-mel arrays
    r = np.random.RandomState(4321 + rank)
    Ts = r.randint(100, 501, size=n_extract)
    specs = [np.log10(I.power_spec(int(T), seed=int(T) + 7 * i) + 1e-6).astype(np.float32) for i, T in enumerate(Ts[:64])]
    specs = [specs[i % 64][:, :int(T)] if specs[i % 64].shape[1] >= T else np.tile(specs[i % 64], (1, 8))[:, :int(T)]
             for i, T in enumerate(Ts)]
    was_training = net.training
    net.eval()
    for _ in range(2):                                          # warm-up (also grows the pinned staging pool)
        outs = svb.extract_dvectors(net, specs)
    dts = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = svb.extract_dvectors(net, specs)
        torch.cuda.synchronize()
        dts.append(time.perf_counter() - t0)
    dt = sorted(dts)[1]                                         # median of 3 whole-job wall times
    nwin = int(sum(max(0, -(-(int(T) - 24) // 12)) for T in Ts))
    ndv = int(sum(len(o) for o in outs))
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["extraction"] = {"utterances_per_gpu": len(specs), "utterances_total": world * len(specs),
                         "windows_per_gpu": nwin, "dvectors_per_gpu": ndv,
                         "seconds": t.item(), "seconds_all_runs": [round(v, 4) for v in dts],
                         "windows_per_s": world * nwin / t.item(),
                         "frac_of_tensor_roofline": (nwin / t.item()) * 572522496.0 / 1e12 /
                         peaks()[0].get("bf16_tflops_sustained", 1400.0),
                         "dvectors_per_s": world * ndv / t.item(), "utterances_per_s": world * len(specs) / t.item(),
                         "path": "host log-mel -> H2D -> window gather -> LSTM fwd (T=24) -> partition mean -> D2H"}
    # device-resident part only: LSTM forward of 32768 windows x 24 frames already in HBM (the tensor-bound core)
    xw = torch.tensor(I.logmel(4096, 24, seed=99)).to(dev).repeat(8, 1, 1)
    with torch.no_grad():
        ms = dev_time(lambda: net(xw), 3, 1)
    out["extraction"]["lstm_forward_windows_per_s_device_resident"] = world * xw.shape[0] / (ms * 1e-3)
    out["extraction"]["lstm_forward_frac_of_tensor_roofline"] = (xw.shape[0] / (ms * 1e-3)) * 572522496.0 / 1e12 / \
        peaks()[0].get("bf16_tflops_sustained", 1400.0)
    net.train(was_training)
    if rank != 0:
        return out

    # ---- the LSTM input projection as a standalone batched GEMM (north_star: tensor-pipe utilisation of the input
    # GEMM; in the default path this product is fused into the persistent forward kernel): all 160 frames x 640
    # utterances of layer 1/2, [102400 x 768] . [768 x 3072], bf16 operands, fp32 out, CTA-pair tcgen05 tiles
    Mg, Ng, Kg = 640 * 160, 3072, 768
    Ag = torch.randn(Mg, Kg, device=dev).to(torch.bfloat16)
    Bg = (torch.randn(Ng, Kg, device=dev) * 0.05).to(torch.bfloat16)
    Cg = torch.empty(Mg, Ng, device=dev)
    PA = (ctypes.c_void_p * 1)(Ag.data_ptr())
    PB = (ctypes.c_void_p * 1)(Bg.data_ptr())

    def input_gemm():
        L.svb_gemm_bf16_2cta(PA, PB, 1, ptr(Cg), None, Mg, Ng, Kg, ctypes.c_int64(Kg), ctypes.c_int64(Kg),
                             ctypes.c_int64(Ng), 0, stream_ptr())

    def input_gemm_persistent():
        L.svb_gemm_persistent(ptr(Ag), ptr(Bg), ptr(Cg), None, Mg, Ng, Kg, ctypes.c_int64(Kg), ctypes.c_int64(Kg),
                              ctypes.c_int64(Ng), 0, stream_ptr())

    ms = dev_time(input_gemm_persistent, 10, 3)
    ms_tile = dev_time(input_gemm, 10, 3)
    tf = 2.0 * Mg * Ng * Kg / (ms * 1e-3) / 1e12
    out["lstm_input_gemm_standalone"] = {"M": Mg, "N": Ng, "K": Kg, "ms": ms, "tflops": tf,
                                         "frac_of_sustained_bf16_peak": tf / peaks()[0].get("bf16_tflops_sustained", 1400.0),
                                         "kernel": "pgemm_kernel (persistent CTA pairs, cta_group::2 256x256 tiles, two TMEM "
                                                   "accumulators: epilogue of tile i under the MMAs of tile i+1)",
                                         "one_tile_per_cta_pair_kernel_ms": ms_tile,
                                         "one_tile_per_cta_pair_tflops": 2.0 * Mg * Ng * Kg / (ms_tile * 1e-3) / 1e12}
    del Ag, Bg, Cg

    # ---- EER sweep at N=1024, M=6 (3 enrollment + 3 verification)
    enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
    enr, ver = torch.tensor(enr).to(dev), torch.tensor(ver).to(dev)
    with torch.no_grad():
        sim = svb.get_cossim(ver, svb.get_centroids(enr)).contiguous()
    thr = E._thresholds_f32(sim.device, E.THRESHOLDS)
    N, Mv, T = 1024, 3, 50
    ca = torch.empty(N, T, dtype=torch.int32, device=dev)
    cd = torch.empty_like(ca)
    scratch = torch.zeros(1 + 16 * T, dtype=torch.int64, device=dev)
    res = torch.empty(4 + 2 * T, device=dev)
    st = stream_ptr()

    def sweep():                      # scratch is re-zeroed by the kernel's last block
        L.svb_eer_sweep(ptr(sim), N, Mv, ptr(thr), T, ptr(ca), ptr(cd), ptr(scratch), ptr(res), st)

    ms = dev_time(sweep, 100, 10)
    with torch.no_grad():
        ms_cos = dev_time(lambda: svb.get_cossim(ver, svb.get_centroids(enr)), 20, 3)
    out["eer"] = {"N": N, "M": 6, "sweep_us": ms * 1e3, "sim_bytes": sim.numel() * 4,
                  "sweep_GBps": sim.numel() * 4 / (ms * 1e-3) / 1e9, "cossim_from_embeddings_us": ms_cos * 1e3,
                  "eer": float(res[0])}

    # ---- fused GE2E forward+backward alone (raw C ABI, preallocated buffers).  Device time per launch: 20 calls
    # captured in a CUDA graph and replayed (a 15 us kernel is shorter than the ctypes + cooperative-launch host
    # path, so a plain loop would time the host); the plain loop is reported next to it.
    def graph_time(fn, reps=20, n=20):
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(reps):
                        fn()
            torch.cuda.current_stream().wait_stream(side)
            return dev_time(g.replay, n, 3) / reps
        except Exception as exc:          # capture of cooperative launches unsupported: report the loop only
            torch.cuda.synchronize()
            return None

    for (Ns, Ms, mode, tag) in ((64, 10, 1, ""), (64, 10, 2, "_general_kernel"), (512, 10, 1, "")):
        Eg = torch.tensor(I.ge2e_embeddings(Ns, Ms, 256, "unit")).to(dev)
        w = torch.tensor(10.0, device=dev)
        b = torch.tensor(-5.0, device=dev)
        nb = ctypes.c_size_t(0)
        L.svb_ge2e_workspace_bytes(Ns, Ms, 256, Ns, ctypes.byref(nb))
        ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        loss = torch.empty((), device=dev)
        dE = torch.empty_like(Eg)
        dw = torch.empty((), device=dev)
        db = torch.empty((), device=dev)

        def call():
            L.svb_ge2e(ptr(Eg), None, Ns, Ms, 256, Ns, ptr(w), ptr(b), None, None, None, None, ptr(loss), ptr(dE),
                       None, ptr(dw), ptr(db), ptr(ws), ctypes.c_size_t(nb.value), mode, stream_ptr())

        ms_loop = dev_time(call, 100, 10)
        ms_graph = graph_time(call)
        ms = ms_graph if ms_graph is not None else ms_loop
        out[f"ge2e_fwd_bwd_N{Ns}{tag}"] = {"us": ms * 1e3, "us_call_loop": ms_loop * 1e3,
                                           "timing": "cuda graph replay" if ms_graph is not None else "call loop",
                                           "algorithmic_bytes": 2 * Eg.numel() * 4,
                                           "GBps": 2 * Eg.numel() * 4 / (ms * 1e-3) / 1e9}
    return out


def workload_config(world):
    """`config` of every arm (ours, --impl reference, --impl reference-cuda): the SAME dict for the same N, so that the
    driver's same-config check compares like with like; what is specific to an arm goes to its `arm_notes`."""
    return {"workload": "GE2E train step, 64 speakers x 10 utts x 160 frames x 40 mel per GPU "
                        "(BASELINE configs[1]; global batch = 64 x n_gpus speakers, configs[2] at 8)",
            "network": "3-layer LSTM 40->768, Linear 768->256, L2 norm; GE2E w=10 b=-5; reference init seed 0",
            "parallelism": f"dp{world} by speaker group" if world > 1 else "single GPU",
            "l2": "GPU arms: 192 MiB buffer written between timed iterations (untimed)"}


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    rate, sample, cores, dt = cpu_reference_rate(steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "utts/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "arm_notes": {"sample": sample, "where": "rank 0's host cores (one process whatever N)"},
            "cpu_baseline": {"value": rate, "unit": "utts/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "utts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def gpu_baseline_block(torch, I, dev, steps, warmup, flush=None, full=True):
    """The reference on the same GPU through stock torch CUDA (bench_reference_arms.TorchCudaReference): C2 train step
    with cuDNN nn.LSTM in fp32 (torch's default: TF32 allowed for cuDNN) and under bf16 autocast, plus -- `full` -- the
    secondary rows (GE2E alone at C2 / C3, get_cossim + threshold loop at C5, a T = 24 extraction batch)."""
    from bench_reference_arms import TorchCudaReference
    ref = TorchCudaReference(torch, dev)
    B = N_SPK * M_UTT
    x_host = torch.tensor(I.logmel(B, T_FR, seed=1234)).pin_memory()
    out = {"kind": ref.kind, "library": f"torch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}",
           "source": "baseline/_ref (unmodified reference modules)" if ref.kind == "reference" else
                     "oracle port of the reference's library calls (baseline/_ref not staged)",
           "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
           "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)}
    for tag, ac in (("fp32", False), ("bf16_autocast", True)):
        ms_full, ms_value = ref.train_step(x_host, N_SPK, M_UTT, steps, warmup, autocast_bf16=ac, flush=flush)
        out[f"train_step_{tag}"] = {"ms_per_step_fwd_loss_bwd": ms_value, "utts_per_s": B / (ms_value * 1e-3),
                                    "ms_per_step_e2e": ms_full, "utts_per_s_e2e": B / (ms_full * 1e-3)}
        torch.cuda.empty_cache()
    if full:
        out["ge2e_fwd_bwd_N64_us"] = ref.ge2e_only(I.ge2e_embeddings(64, 10, 256, "unit")) * 1e3
        out["ge2e_fwd_bwd_N512_us"] = ref.ge2e_only(I.ge2e_embeddings(512, 10, 256, "unit"), steps=5, warmup=2) * 1e3
        torch.cuda.empty_cache()
        enr, ver = I.eer_embeddings(1024, 6, 0.06, 0.5, 4242)
        e = ref.eer(enr, ver)
        if e is not None:
            out["eer_N1024"] = {"cossim_from_embeddings_us": e[0] * 1e3, "threshold_loop_us": e[1] * 1e3, "eer": e[2],
                                "loop": "train_speech_embedder.py:132-149 executed from baseline/_ref on CUDA tensors"}
        torch.cuda.empty_cache()
        xw = torch.tensor(I.logmel(4096, 24, seed=99)).to(dev).repeat(8, 1, 1)
        ms = ref.forward_windows(xw)
        out["lstm_forward_windows_per_s_device_resident"] = xw.shape[0] / (ms * 1e-3)
        del xw
        torch.cuda.empty_cache()
    return out


def run_reference_cuda(args, rank, local_rank):
    """--impl reference-cuda: one JSON line in the main line's format, measured on the reference's stock-torch path."""
    if rank != 0:
        return
    import torch
    import _inputs as I
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    steps, warm = max(1, args.steps), max(3, args.warmup)
    blk = gpu_baseline_block(torch, I, dev, steps, warm, flush=flush, full=True)
    B = N_SPK * M_UTT
    best = blk["train_step_fp32"]
    line = {"impl": "reference-cuda", "metric": METRIC, "value": best["utts_per_s"], "unit": "utts/s", "n_gpus": 1,
            "steps": steps, "warmup": warm, "ms_per_step": best["ms_per_step_fwd_loss_bwd"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (cuDNN, TF32 allowed)", "data": "synthetic",
            "config": workload_config(1),
            "e2e": {"value": best["utts_per_s_e2e"], "unit": "utts/s", "ms_per_step": best["ms_per_step_e2e"],
                    "h2d_bytes_per_step": B * T_FR * NMELS * 4, "d2h_bytes_per_step": 4},
            "gpu_baseline": blk}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--extract-utts", type=int, default=12500, help="utterances per GPU of the extraction row (configs[3]: 100k over 8 GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--recurrent-terms", type=int, default=1)
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: everything libraries print to fd 1 meanwhile (NCCL's version
    # banner, warnings) goes to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "reference-cuda":
        run_reference_cuda(args, rank, local_rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import _inputs as I
    import pytorch_speaker_verification_b200 as svb
    from pytorch_speaker_verification_b200 import _lib
    from pytorch_speaker_verification_b200.dist import GlobalGE2ELoss, allreduce_gradients

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    torch.manual_seed(0)                                   # same init on every rank (reference init order)
    net = svb.SpeechEmbedder().to(dev)
    net.recurrent_terms = args.recurrent_terms
    crit = svb.GE2ELoss(dev)
    # GE2E exchange of the N-rank step: "rows" (default) = NCCL all-gather of the centroids + all-reduce of their
    # gradients; "peer" = the same two steps as plain loads over NVLink peer memory (symmetric buffers, two device-side
    # barriers, no NCCL call).  Measured equal at 2 GPUs (9.51-9.59 vs 9.55-9.70 ms/step: the exchange is ~0.15 ms of
    # launch latencies either way), so the default stays the one with fewer launches; SVB_GE2E_EXCHANGE selects.
    ge2e_exchange = os.environ.get("SVB_GE2E_EXCHANGE", "rows")
    loss_mod = GlobalGE2ELoss(crit, mode=ge2e_exchange) if world > 1 else crit
    params = list(net.parameters())
    opt = torch.optim.SGD([{'params': net.parameters()}, {'params': crit.parameters()}], lr=0.01)
    B = N_SPK * M_UTT
    x_host = torch.tensor(I.logmel(B, T_FR, seed=1234 + rank)).pin_memory()
    x_dev = x_host.to(dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # gradient all-reduce at N > 1: SVB_ALLREDUCE=peer (two-shot all-reduce over NVLink peer memory after backward,
    # csrc/peer.cu), nccl_overlap (four NCCL buckets started from inside backward) or nccl (one NCCL call after backward)
    ar_mode = os.environ.get("SVB_ALLREDUCE", "nccl_overlap")
    if os.environ.get("SVB_ALLREDUCE_OVERLAP", "1") == "0" and ar_mode == "nccl_overlap":
        ar_mode = "nccl"
    reducer = None
    if world > 1 and ar_mode == "nccl_overlap":
        # default at N > 1: the gradient all-reduce runs in four buckets started from inside backward, beside the
        # remaining weight-gradient GEMMs (SVB_ALLREDUCE_OVERLAP=0: one all-reduce after backward)
        from pytorch_speaker_verification_b200.dist import OverlappedGradReducer
        reducer = OverlappedGradReducer()

    loss_events = []                        # (start, end) CUDA events around the loss forward, filled while profiling

    def fwd_bwd(x, profile=False):
        emb = net(x)
        if profile:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
            loss = loss_mod(emb.reshape(N_SPK, M_UTT, PROJ))
            ev[1].record()
            loss_events.append(ev)
        else:
            loss = loss_mod(emb.reshape(N_SPK, M_UTT, PROJ))
        if reducer is not None:               # bucketed all-reduce started from inside backward
            with reducer:
                loss.backward()
            return loss
        loss.backward()
        if world > 1:
            allreduce_gradients(params, peer=(ar_mode == "peer"))
        return loss

    def full_step():
        x = x_host.to(dev, non_blocking=True)
        opt.zero_grad()
        loss = fwd_bwd(x)
        torch.nn.utils.clip_grad_norm_(net.parameters(), 3.0)
        torch.nn.utils.clip_grad_norm_(crit.parameters(), 1.0)
        opt.step()
        loss_host.copy_(loss.detach(), non_blocking=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in evs:
            flush.zero_()                                   # L2 flush between timed iterations (untimed)
            s.record()
            fn()
            e.record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps

    def phase_profile(fn, steps=3):
        """Per-phase device time (CUDA events recorded by the library on the launching stream), averaged."""
        acc = [0.0] * len(PHASES)
        buf = (ctypes.c_float * 16)()
        for _ in range(steps):
            flush.zero_()
            L.svb_profile_enable(1)
            fn()
            torch.cuda.synchronize()
            L.svb_profile_read(buf, 16)
            for i in range(len(PHASES)):
                acc[i] += buf[i]
        L.svb_profile_enable(0)
        return {n: acc[i] / steps for i, n in enumerate(PHASES)}

    def value_step(profile=False):
        for p in list(params) + [crit.w, crit.b]:
            p.grad = None
        return fwd_bwd(x_dev, profile)

    def dist_check():
        """N ranks == 1 rank: the sharded step (speakers split over the ranks, d-vector all-gather, global GE2E, SUM
        all-reduce of the gradients) against the same global batch run by rank 0 alone (40 frames to keep it short)."""
        Tc = 40
        xs = torch.tensor(I.logmel(B, Tc, seed=4321 + rank)).to(dev)
        for p in list(params) + [crit.w, crit.b]:
            p.grad = None
        loss = fwd_bwd(xs)
        torch.cuda.synchronize()
        sharded = [p.grad.clone() for p in params] + [crit.w.grad.clone()]
        x_all = torch.empty(world * B, Tc, NMELS, device=dev)
        dist.all_gather_into_tensor(x_all, xs)
        res = None
        if rank == 0:
            for p in list(params) + [crit.w, crit.b]:
                p.grad = None
            l1 = crit(net(x_all).reshape(world * N_SPK, M_UTT, PROJ))
            l1.backward()
            torch.cuda.synchronize()
            single = [p.grad for p in params] + [crit.w.grad]
            errs = [float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)) for a, b in zip(sharded, single)]
            lerr = abs(float(loss) - float(l1)) / abs(float(l1))
            res = {"ranks": world, "speakers": world * N_SPK, "frames": Tc, "loss_rel_err": lerr,
                   "grad_rel_l2_max": max(errs), "ok": bool(lerr < 1e-5 and max(errs) < 5e-3),
                   "what": "sharded step (all-gather + global GE2E + SUM all-reduce) vs rank 0 alone on the global batch"}
        for p in list(params) + [crit.w, crit.b]:
            p.grad = None
        barrier()
        return res

    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None)) if rank == 0 else None
    ms_value = timed(value_step, args.steps, args.warmup)
    ms_e2e = timed(full_step, args.steps, args.warmup)
    # section 8(f) additions, opt-in for a caller: H2D of the next batch on a copy stream (staging.prefetch) and the
    # fused clip_grad_norm_ x2 + SGD tail (FusedClipSGD) instead of the stock torch calls of `full_step`
    fopt = svb.FusedClipSGD([{"params": net.parameters(), "max_norm": 3.0}, {"params": crit.parameters(), "max_norm": 1.0}], lr=0.01)

    def host_batches():
        while True:
            yield x_host

    staged = svb.prefetch(host_batches(), dev, flatten=False)

    def full_step_fused():
        x = next(staged)
        fopt.zero_grad()
        loss = fwd_bwd(x)
        fopt.step()
        loss_host.copy_(loss.detach(), non_blocking=True)
        return loss

    ms_e2e_fused = timed(full_step_fused, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    phases = phase_profile(lambda: value_step(True))          # every rank: the step contains collectives
    # GE2E forward+gradients as the step runs it: one fused launch on one GPU; at N > 1 the centroid all-gather, this
    # rank's rows against all centroids, the all-reduce of [dC | loss, dw, db] and the waits for the slowest rank
    phases["ge2e_loss_incl_collectives"] = sum(a.elapsed_time(b) for a, b in loss_events) / max(1, len(loss_events))
    barrier()
    loss_val = float(loss_host)
    check = dist_check() if world > 1 else None
    launch_count = count_launches(L, value_step, torch)      # one extra, untimed step seen through CUPTI
    extra = secondary_benchmarks(torch, dist, svb, _lib, I, net, dev, rank, world, args.extract_utts)
    extra["e2e_with_prefetch_and_fused_clip_sgd"] = {
        "value": B * world / (ms_e2e_fused * 1e-3), "unit": "utts/s", "ms_per_step": ms_e2e_fused,
        "step": "svb.prefetch (H2D of the next pinned batch on a copy stream) + zero_grad + fwd + GE2E + bwd + "
                "svb.FusedClipSGD (clip 3.0 / 1.0 + SGD, two launches) + D2H loss",
        "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4}

    if rank == 0:
        pk, pk_src = peaks()
        utts = B * world
        # roofline of the dominant kernel class
        flops_step = 2.0 * B * 4 * HID * HID               # one recurrent frame: (B x H) . (H x 4H), fwd == bwd
        dom = max(("recurrent_fwd", "recurrent_bwd", "input_gemm", "weight_grads"), key=lambda k: phases[k])
        if dom == "recurrent_fwd" and args.recurrent_terms == 1:
            launches = 1                                  # one persistent launch: input projections + recurrence, all layers
            flops_launch = 23838720.0 * T_FR * B
            kern = "wlstm_fwd_kernel (persistent wavefront LSTM forward)"
        elif dom == "recurrent_bwd" and args.recurrent_terms == 1:
            launches = 1                                  # one persistent launch: recurrent + dX products, all layers
            flops_launch = 2.0 * B * T_FR * 4 * HID * HID * (2 * NLAYER - 1)
            kern = "wbptt_kernel (persistent wavefront BPTT)"
        elif dom in ("recurrent_fwd", "recurrent_bwd"):
            launches = NLAYER * T_FR
            flops_launch = flops_step * args.recurrent_terms if dom == "recurrent_fwd" else flops_step
            kern = "tc_gemm_kernel<EpiLstmFwd>" if dom == "recurrent_fwd" else "tc_gemm_kernel<EpiLstmBwd>"
        elif dom == "input_gemm":
            launches = NLAYER
            flops_launch = 2.0 * B * T_FR * 4 * HID * (NMELS + 2 * HID) / 3 * 3   # 3 split-bf16 terms (averaged/layer)
            kern = "tc_gemm_kernel<EpiStoreF32> (input projection)"
        else:
            launches = 2 * NLAYER
            flops_launch = 2.0 * B * T_FR * 4 * HID * (NMELS + 5 * HID) / 6
            kern = "tc_gemm_kernel<EpiStoreF32, MN-major> (weight gradients)"
        avg_ms = phases[dom] / launches
        achieved = flops_launch / (avg_ms * 1e-3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        line = {
            "metric": METRIC, "value": utts / (ms_value * 1e-3), "unit": "utts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16 fwd / bf16 bwd operands, fp32 accumulate",
            "data": "synthetic",
            "config": workload_config(world),
            "arm_notes": {"precision": "forward GEMMs fp16 x fp16 (single term), MUFU.TANH gates; BPTT GEMMs bf16; "
                                       "fp32 accumulate/cell state/loss",
                          "e2e_step": "H2D pinned batch + zero_grad + fwd + GE2E + bwd + clip_grad_norm_ x2 + SGD + D2H loss"},
            "e2e": {"value": utts / (ms_e2e * 1e-3), "unit": "utts/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": (launch_count[0] if launch_count else launches_per_step(world)) * args.steps,
            "gpu_launches_detail": ({"per_step": launch_count[0], "other_libraries_per_step": launch_count[1],
                                     "kernels": launch_count[2],
                                     "source": "CUPTI callback count of one extra untimed step (svb_launch_count_*)"}
                                    if launch_count else {"per_step": launches_per_step(world),
                                                          "source": "formula (CUPTI unavailable)"}),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch from this round's ncu --set full
                         # capture (profiles/r2_traffic.json), else null
                         "traffic": measured_traffic(dom) if args.recurrent_terms == 1 else None,
                         "peak_source": f"{pk_src} (sustained bf16)",
                         "avg_launch_us": avg_ms * 1e3, "launches_per_step": launches,
                         # whole step against the 3x-forward convention of BASELINE.md (11,443,765,248 FLOP/utt)
                         "whole_step_frac": 11443765248.0 * B / 1e12 / (ms_value * 1e-3) / peak},
            "phases_ms": phases,
            "forward_kernel": {"name": "wlstm_fwd_kernel", "ms": phases["recurrent_fwd"],
                               "tflops": 23838720.0 * T_FR * B / (phases["recurrent_fwd"] * 1e-3) / 1e12,
                               "frac_of_sustained_bf16_peak": 23838720.0 * T_FR * B / (phases["recurrent_fwd"] * 1e-3) / 1e12 / peak},
            "secondary": extra,
            "loss": loss_val,
        }
        if check is not None:
            line["dist_check"] = check
        accounted = sum(phases.values())
        line["phases_ms"]["unattributed"] = ms_value - accounted
        line["phases_note"] = ("phases_ms are CUDA-event brackets inside svb_embedder_forward/backward; `unattributed` = "
                               "ms_per_step - their sum: the GE2E launch, autograd / custom-op dispatch gaps between "
                               "the library calls, and (N > 1) the all-gather and the gradient all-reduce")
        if world == 1 and not args.no_gpu_baseline:
            del flush
            torch.cuda.empty_cache()
            try:
                line["gpu_baseline"] = gpu_baseline_block(torch, I, dev, max(3, min(args.steps, 10)), 3)
            except Exception as exc:          # the reference arm must never take the main line down
                line["gpu_baseline"] = {"unavailable": repr(exc)}
        if world == 1 and not args.no_cpu_baseline:
            rate, sample, cores, _ = cpu_reference_rate(2, 1, budget_s=40.0)
            line["cpu_baseline"] = {"value": rate, "unit": "utts/s", "cores": cores, "kind": "port", "sample": sample}
            try:
                from bench_reference_arms import cpu_secondary
                line["cpu_baseline"]["secondary"] = cpu_secondary(torch, I)
            except Exception as exc:
                line["cpu_baseline"]["secondary"] = {"unavailable": repr(exc)}
        gb, cb = line.get("gpu_baseline") or {}, (line.get("cpu_baseline") or {}).get("secondary") or {}
        if world == 1 and ("train_step_fp32" in gb or cb):
            sec = line["secondary"]

            def g(d, *ks):
                for k in ks:
                    d = d.get(k) if isinstance(d, dict) else None
                return d

            rows = {
                "train_step_utts_per_s": {"ours": line["value"], "ours_e2e": line["e2e"]["value"],
                                          "torch_cuda_fp32": g(gb, "train_step_fp32", "utts_per_s"),
                                          "torch_cuda_fp32_e2e": g(gb, "train_step_fp32", "utts_per_s_e2e"),
                                          "torch_cuda_bf16_autocast": g(gb, "train_step_bf16_autocast", "utts_per_s"),
                                          "cpu": g(line, "cpu_baseline", "value")},
                "ge2e_fwd_bwd_N64_us": {"ours": g(sec, "ge2e_fwd_bwd_N64", "us"), "torch_cuda": gb.get("ge2e_fwd_bwd_N64_us"),
                                        "cpu": cb.get("ge2e_fwd_bwd_N64_ms") and cb["ge2e_fwd_bwd_N64_ms"] * 1e3},
                "ge2e_fwd_bwd_N512_us": {"ours": g(sec, "ge2e_fwd_bwd_N512", "us"), "torch_cuda": gb.get("ge2e_fwd_bwd_N512_us")},
                "eer_N1024_cossim_us": {"ours": g(sec, "eer", "cossim_from_embeddings_us"),
                                        "torch_cuda": g(gb, "eer_N1024", "cossim_from_embeddings_us"),
                                        "cpu": g(cb, "eer_N1024", "cossim_ms") and cb["eer_N1024"]["cossim_ms"] * 1e3},
                "eer_N1024_threshold_sweep_us": {"ours": g(sec, "eer", "sweep_us"),
                                                 "torch_cuda": g(gb, "eer_N1024", "threshold_loop_us"),
                                                 "cpu": g(cb, "eer_N1024", "threshold_loop_ms") and cb["eer_N1024"]["threshold_loop_ms"] * 1e3},
                "extraction_windows_per_s": {"ours_e2e_from_host": g(sec, "extraction", "windows_per_s"),
                                             "ours_device_resident": g(sec, "extraction", "lstm_forward_windows_per_s_device_resident"),
                                             "torch_cuda_device_resident": gb.get("lstm_forward_windows_per_s_device_resident"),
                                             "cpu_per_file": g(cb, "extraction", "windows_per_s")},
            }
            lower_is_better = {"ge2e_fwd_bwd_N64_us", "ge2e_fwd_bwd_N512_us", "eer_N1024_cossim_us", "eer_N1024_threshold_sweep_us"}
            lost = []
            for name, r in rows.items():
                ours = [v for k, v in r.items() if k.startswith("ours") and v]
                theirs = [v for k, v in r.items() if k.startswith("torch_cuda") and v]
                if ours and theirs:
                    if (name in lower_is_better and min(theirs) < min(ours)) or (name not in lower_is_better and max(theirs) > max(ours)):
                        lost.append(name)
            line["side_by_side"] = {"rows": rows, "rows_where_stock_torch_cuda_wins": lost,
                                    "cpu_cores": (line.get("cpu_baseline") or {}).get("cores")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
