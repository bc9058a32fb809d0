#!/usr/bin/env python
"""Benchmark of the IDEAS GAN-conv hot path on B200 (contract: see README / DESIGN.md §Measurement).

    python bench.py --gpus 1 --steps 16 --warmup 3                 # our arm, one JSON line
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W     # data parallel, weak scaling
    python bench.py --impl reference --steps K --warmup W          # the reference algorithm on host cores

A "step" is one full IDEAS training iteration (train.py:33-221: D phase, lazy R1 every 16
iterations, G/E/Ex phase with two backwards, EMA) on one synthetic LSUN-shaped batch
(BASELINE.json configs[3]: N=1, sigma=1, 256x256, batch 32 per GPU).  ``value`` is images/s of the
whole job with the batch already resident in HBM; ``e2e`` repeats the measurement through the
public Trainer.step API with the batch copied from pinned host memory and the loss scalars read
back every step.  ``roofline`` times the dominant kernel (the 3x3 modulated-conv implicit GEMM of
BASELINE.json configs[2]) live with CUDA events; ``roofline_hbm`` does the same for the blur kernel
(configs[1]).  ``cpu_baseline`` times the oracle port of the same train step on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ideas_256x256_train_step_images_per_sec"
UNIT = "images/s"
FLOP_PER_IMAGE_STEP = 2.67e12          # SURVEY.md App. B (estimate, used only for a reported TFLOP/s figure)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step (BASELINE configs[3]: 32)")
    ap.add_argument("--image-size", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--case", default=None, help="run ONE roofline kernel alone (for ncu); see CASES in bench.py")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch-g", action="store_true", help="A/B: one batched Generator call per phase instead of three")
    ap.add_argument("--split-dreal", action="store_true", help="A/B: Dreal on the three fake batches separately, on the generator streams")
    ap.add_argument("--no-concurrent-g", action="store_true", help="A/B: the three Generator calls of a phase on ONE stream")
    ap.add_argument("--early-g", action="store_true", help="A/B: start the G phase's generator-side forward under the D backward")
    ap.add_argument("--single-stream", action="store_true", help="A/B: capture the step on one stream (no side-stream branches)")
    ap.add_argument("--no-graphs", action="store_true", help="run the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--prune-dead-backward", action="store_true",
                    help="NOT the default measurement: restrict the loop's second backward (train.py:214-216) to Ex's parameters")
    ap.add_argument("--share-recon", action="store_true",
                    help="NOT the default measurement: evaluate E(X) and G(E(X)) once per iteration instead of once per phase")
    return ap.parse_args()


def workload_name(batch, size):
    return (f"cfg4: full IDEAS N=1 sigma=1 {size}x{size} train step (E+G+Gstru+Ex+Dreal+Dco+Ddist, lazy R1 every 16), "
            f"synthetic LSUN-shaped batch {batch}/GPU")


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
CPU_BATCH = 4          # images per CPU step: a bounded sample of the batch-32 workload


def find_reference_tree():
    """The unmodified reference, when a copy is reachable: /root/reference (build container) or a tree staged for
    one GPU run under baseline/_ref/IDEAS (scripts/stage_reference.sh; git-ignored, never committed)."""
    for cand in (os.environ.get("IDEAS_REFERENCE", ""), "/root/reference", os.path.join(ROOT, "baseline", "_ref", "IDEAS")):
        if cand and os.path.exists(os.path.join(cand, "train.py")) and os.path.exists(os.path.join(cand, "models.py")):
            return cand
    return None


def cpu_reference_step_time(image_size, steps, warmup, budget_s=150.0, batch=CPU_BATCH):
    """The reference algorithm for the path on the host cores, one training iteration (train.py:33-221) on ``batch``
    synthetic images per step.  With a reference tree reachable the UNMODIFIED reference runs (kind "reference",
    scripts/reference_step.py drives its train() on the CPU through the App. D shims); otherwise the oracle port
    (kind "port").  Returns (seconds per step, threads, timed steps, kind)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    ref = find_reference_tree()
    if ref is not None:
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import reference_step
            t, n = reference_step.time_reference(ref, device="cpu", batch=batch, image_size=image_size, steps=steps,
                                                 warmup=warmup, budget_s=budget_s)
            return t, torch.get_num_threads(), n, "reference"
        except Exception as e:       # a broken staged tree must not take the arm down: fall back to the port
            print(f"[bench] reference tree at {ref} unusable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    from oracle.train_step import OracleTrainer
    tr = OracleTrainer(seed=0, image_size=image_size)
    torch.manual_seed(1)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        X = torch.rand(batch, 3, image_size, image_size) * 2 - 1
        t0 = time.perf_counter()
        tr.step(X, i + 1)          # iterations 1.. : lazy R1 falls in only when i+1 is a multiple of 16
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        # keep the whole arm within a few minutes: stop early (after >= 1 timed step) once the budget is spent
        if times and time.perf_counter() - t_start + dt > budget_s:
            break
    return sum(times) / len(times), torch.get_num_threads(), len(times), "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, cores, timed, kind = cpu_reference_step_time(args.image_size, args.steps, min(args.warmup, 1))
    v = CPU_BATCH / t
    what = "the unmodified reference (train.py loop)" if kind == "reference" else "oracle port of train.py:33-221"
    sample = (f"{CPU_BATCH} images per step ({args.image_size}x{args.image_size}), full train iteration on the host cores, "
              f"{what}; {timed} timed steps within the time budget")
    print(json.dumps({
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": timed, "steps_requested": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args.batch, args.image_size), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------
def time_kernel(fn, iters=20, warm=3, warm_ms=0.0):
    """Average launch time after ``warm`` launches (and optionally ``warm_ms`` of GPU time).  The roofline cases are
    timed in BURST mode -- a short cool-down, three warm-up launches, ten timed ones -- to match the burst peaks they
    are divided by: this GPU is power limited, and the same kernel runs ~12 % slower after 60 ms of continuous tensor
    load (measured: cfg-3 forward 735 vs 636 TFLOP/s), which is the regime of the sustained peak, not the burst one."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e3 < warm_ms:
        for _ in range(4):
            fn()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def measure_tf32_gemm_peak(seconds=1.5):
    """cuBLAS TF32 GEMM (fp32 operands, allow_tf32) 8192^3 on this GPU, now: burst = best single launch of 10,
    sustained = back-to-back launches for ``seconds``.  This is the tensor roofline SURVEY.md §8(d) cfg 3 asks for;
    MEASURED_PEAKS.json only carries the bf16 figure (kind::tf32 issues at half that rate)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        c = torch.empty(n, n, device="cuda")
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        flops = 2.0 * n ** 3
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        reps = max(4, int(seconds / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        sustained = e0.elapsed_time(e1) * 1e-3 / reps
        return {"burst": flops / best / 1e12, "sustained": flops / sustained / 1e12,
                "how": f"torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 10 / {reps} back to back"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def load_traffic():
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, one `ncu --set full` capture per
    kernel) of the bench kernels, committed under profiles/ by scripts/ncu_cases.sh + scripts/ncu_summary.py."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


# ---- the kernels bench.py reports a roofline for; `python bench.py --case NAME` runs one of them alone (the
# command ncu captures, scripts/ncu_cases.sh) ---------------------------------------------------------------
def make_case(name):
    """Returns (fn, work, unit_kind, description): fn launches the kernel once through the C ABI on the current
    stream; work = algorithmic FLOPs (tensor bound) or bytes (hbm bound) of one launch."""
    import torch
    from ideas_b200 import _lib
    from ideas_b200._tensor import ptr, stream_ptr
    dev = torch.device("cuda")
    impl = _lib.IMPL_AUTO
    if name in ("conv_fwd_cfg3", "conv_dgrad_cfg3", "conv_wgrad_cfg3"):
        # cfg 3: ModulatedConv2d 512->512 3x3 @64x64, batch 16 (x is 134 MB > 126 MB L2)
        N, C, K, H = 16, 512, 512, 64
        x = torch.randn(N, H, H, C, device=dev)
        wp = torch.randn(9, K, C, device=dev) / (C * 9) ** 0.5
        d = torch.rand(N, K, device=dev) + 0.5
        bias = torch.randn(K, device=dev)
        y = torch.randn(N, H, H, K, device=dev)
        flops = 2.0 * N * H * H * K * C * 9
        if name == "conv_fwd_cfg3":
            def fn():
                _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(bias), N, H, H, C, K, 3, 3,
                          1, 1, _lib.ACT_LRELU, 0.2, 2 ** 0.5, impl, stream_ptr(x))
            return fn, flops, "tensor", "modulated 3x3 conv fwd, 16x512x64x64 -> 512, demod+bias+lrelu fused"
        if name == "conv_dgrad_cfg3":
            def fn():
                _lib.call("ideas_conv2d_dgrad", ptr(x), ptr(y), ptr(wp), ptr(None), ptr(None), ptr(None), N, H, H, C, K, 3, 3,
                          1, 1, H, H, _lib.ACT_NONE, 0.2, 1.0, impl, stream_ptr(x))
            return fn, flops, "tensor", "3x3 conv data gradient, cfg-3 shape"
        dwp = torch.zeros(9, K, C, device=dev)

        def fn():
            _lib.call("ideas_conv2d_wgrad", ptr(dwp), ptr(x), ptr(y), ptr(None), ptr(None), N, H, H, C, K, 3, 3, 1, 1, H, H,
                      impl, stream_ptr(x))
        return fn, flops, "tensor", "3x3 conv weight gradient, cfg-3 shape (16x512x64x64, 512->512)"
    if name in ("conv_halo_g256", "conv_lowch_e256"):
        # the widest low-channel layers: 128->128 @256^2 x32 (G / Dreal) and 32->64 @258^2 valid (E, reflection padded)
        N2, C2, K2, H2, pad = (32, 128, 128, 256, 1) if name == "conv_halo_g256" else (32, 32, 64, 258, 0)
        OH = H2 + 2 * pad - 2
        x2 = torch.randn(N2, H2, H2, C2, device=dev)
        w2 = torch.randn(9, K2, C2, device=dev) / (C2 * 9) ** 0.5
        y2 = torch.empty(N2, OH, OH, K2, device=dev)
        b2 = torch.randn(K2, device=dev)

        def fn():
            _lib.call("ideas_conv2d_forward", ptr(y2), ptr(x2), ptr(w2), ptr(None), ptr(None), ptr(b2), N2, H2, H2, C2, K2, 3, 3,
                      1, pad, _lib.ACT_LRELU, 0.2, 2 ** 0.5, impl, stream_ptr(x2))
        return fn, 2.0 * N2 * OH * OH * C2 * K2 * 9, "tensor", f"3x3 conv fwd, {N2}x{C2}x{H2}x{H2} -> {K2}, bias+lrelu fused"
    B, Cb, Hb = 32, 128, 256
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k = torch.outer(k, k)
    k = (k / k.sum()).contiguous()
    if name == "blur_cfg2":
        # cfg 2 (i): Blur pad (2,2), (32,128,256,256) -> (32,128,257,257)
        xb = torch.randn(B, Hb, Hb, Cb, device=dev)
        yb = torch.empty(B, Hb + 1, Hb + 1, Cb, device=dev)

        def fn():
            _lib.call("ideas_upfirdn2d", ptr(yb), ptr(xb), ptr(k), B, Hb, Hb, Cb, 4, 4, 1, 1, 1, 1, 2, 2, 2, 2, ptr(None), 0.2,
                      1.0, stream_ptr(xb))
        return fn, 4.0 * (xb.numel() + yb.numel()), "hbm", "upfirdn2d up=down=1, 4x4, pad (2,2), 32x128x256x256"
    if name == "blur_act_bwd_cfg2":
        # backward of conv -> FusedLeakyReLU -> Blur(pad 2,2): blur^T of the gradient, activation mask, bias-grad sums
        gz = torch.randn(B, Hb + 1, Hb + 1, Cb, device=dev)
        yact = torch.randn(B, Hb, Hb, Cb, device=dev)
        g1 = torch.empty_like(yact)
        gb = torch.zeros(Cb, device=dev)

        def fn():
            _lib.call("ideas_blur_act_backward", ptr(g1), ptr(gb), ptr(gz), ptr(yact), ptr(k), B, Hb + 1, Hb + 1, Cb, 4, 4,
                      1, 1, 1, 1, 0.2, 2 ** 0.5, stream_ptr(gz))
        return fn, 4.0 * (gz.numel() + 2 * yact.numel()), "hbm", "fused blur^T x lrelu mask + bias-grad, 32x128x257x257 -> 256x256"
    if name in ("lrelu_fwd_cfg2", "lrelu_bwd_cfg2"):
        # cfg 2 (iii): FusedLeakyReLU(128) on (32,128,256,256)
        xa = torch.randn(B, Hb, Hb, Cb, device=dev)
        ba = torch.randn(Cb, device=dev)
        ya = torch.empty_like(xa)
        n = xa.numel()
        if name == "lrelu_fwd_cfg2":
            def fn():
                _lib.call("ideas_fused_bias_act", ptr(ya), ptr(xa), ptr(ba), ptr(None), 3, 0, 0.2, 2 ** 0.5, n, 1, Cb, stream_ptr(xa))
            return fn, 8.0 * n + 4 * Cb, "hbm", "FusedLeakyReLU fwd, 32x128x256x256 (NHWC)"
        gba = torch.zeros(Cb, device=dev)
        ga = torch.randn_like(xa)

        def fn():
            _lib.call("ideas_bias_act_backward", ptr(ya), ptr(gba), ptr(ga), ptr(xa), 0.2, 2 ** 0.5, n, 1, Cb, stream_ptr(xa))
        return fn, 12.0 * n + 4 * Cb, "hbm", "FusedLeakyReLU bwd (input grad + bias grad in one pass), 32x128x256x256"
    raise KeyError(name)


CASES = {"roofline": "conv_fwd_cfg3", "roofline_dgrad": "conv_dgrad_cfg3", "roofline_wgrad": "conv_wgrad_cfg3",
         "roofline_halo": "conv_halo_g256", "roofline_lowch": "conv_lowch_e256", "roofline_hbm": "blur_cfg2",
         "roofline_blur_act_bwd": "blur_act_bwd_cfg2", "roofline_lrelu_fwd": "lrelu_fwd_cfg2",
         "roofline_lrelu_bwd": "lrelu_bwd_cfg2"}


def roofline_sections(peaks, peak_kind):
    """Dominant kernels timed alone, live, with CUDA events on the launching (current) stream.  Runs BEFORE the
    training loop: the kernels are divided by burst peaks (MEASURED_PEAKS.json, and the cuBLAS TF32 burst measured
    here), and this GPU is power limited -- after 30 s of training at the cap the same kernel is 10-15 % slower
    (scripts/bench_sustained.py), which is the regime of the sustained peak.  The sustained GEMM figure is measured
    after the training loop (sustained_tf32_gemm)."""
    import torch
    from ideas_b200 import _lib
    out = {}
    time.sleep(1.0)
    tf32 = measure_tf32_gemm_peak(seconds=0.0)
    tf32["sustained"] = None
    traffic = load_traffic()
    umma = _lib.umma_enabled()
    half_bf16 = peaks["bf16_tflops"] * 0.5
    for key, case in CASES.items():
        fn, work, bound, desc = make_case(case)
        torch.cuda.synchronize()
        time.sleep(1.0)                      # cool-down between cases: every case starts from the same burst state
        t = time_kernel(fn, iters=10, warm=3)
        tr = traffic.get(case, {})
        if bound == "tensor":
            ach = work / t / 1e12
            # peak = the larger of the two tf32 anchors, so the fraction is never flattered: cuBLAS TF32 GEMM measured
            # here a moment ago (burst), and half of the driver-measured bf16 burst (kind::tf32 issues at half rate)
            peak = max(tf32["burst"], half_bf16)
            sec = {"kernel": f"{case}: {desc}", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                   "frac": ach / peak, "traffic": tr.get("dram_bytes"), "traffic_source": tr.get("source"),
                   "peak_source": f"max(cuBLAS TF32 GEMM burst measured in this run {tf32['burst']:.1f}, {peak_kind} "
                                  f"MEASURED_PEAKS.json bf16_tflops {peaks['bf16_tflops']} x 0.5 = {half_bf16:.1f})",
                   "frac_of_cublas_tf32_burst": ach / tf32["burst"],
                   "frac_of_bf16_peak": ach / peaks["bf16_tflops"],
                   "path": "tcgen05 kind::tf32" if umma else "fp32 FFMA (SIMT)", "algorithmic_flops_per_launch": work,
                   "ms_per_launch": t * 1e3}
        else:
            ach = work / t / 1e9
            sec = {"kernel": f"{case}: {desc}", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": ach / peaks["hbm_gbs"], "traffic": tr.get("dram_bytes"), "traffic_source": tr.get("source"),
                   "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs", "algorithmic_bytes_per_launch": work,
                   "ms_per_launch": t * 1e3}
        out[key] = sec
        del fn
        torch.cuda.empty_cache()
    out["tf32_gemm_peak"] = tf32
    return out


def sustained_tf32_gemm(line):
    """cuBLAS TF32 GEMM back to back for 1.5 s, after the training loop (the chip is at its power cap)."""
    m = measure_tf32_gemm_peak(seconds=1.5)
    sus = m["sustained"]
    line["tf32_gemm_peak"]["sustained"] = sus
    line["tf32_gemm_peak"]["how"] = m["how"] + "; burst before the training loop, sustained after it"
    for key, sec in line.items():
        if key.startswith("roofline") and isinstance(sec, dict) and sec.get("bound") == "tensor":
            sec["frac_of_cublas_tf32_sustained"] = sec["achieved"] / sus


def run_case(args):
    """`bench.py --case NAME [--iters N]`: launch one bench kernel alone (what scripts/ncu_cases.sh profiles)."""
    import torch
    torch.cuda.set_device(0)
    fn, work, bound, desc = make_case(args.case)
    t = time_kernel(fn, iters=args.iters, warm=2)
    rate = work / t / (1e12 if bound == "tensor" else 1e9)
    print(json.dumps({"case": args.case, "desc": desc, "ms_per_launch": t * 1e3, "rate": rate,
                      "unit": "TFLOP/s" if bound == "tensor" else "GB/s"}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ideas_b200 import _lib
    from ideas_b200.train_step import Trainer, default_args

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    out = sys.stdout
    if world > 1:
        # stdout carries exactly one JSON line.  NCCL writes its version banner (and NCCL_DEBUG output) to file
        # descriptor 1 from C, so the descriptor itself is pointed at stderr and the JSON line goes to a saved copy.
        sys.stdout.flush()
        out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    roof = None
    if world == 1 and not args.no_roofline:
        peaks, kind = load_peaks()
        roof = roofline_sections(peaks, kind)
        torch.cuda.empty_cache()
    B, S = args.batch, args.image_size
    targs = default_args(batch_size=B, image_size=S)
    tr = Trainer(targs, device=dev, seed=0, cuda_graphs=not args.no_graphs,   # same seed on every rank => identical replicas
                 multi_stream=False if args.single_stream else None,
                 prune_dead_backward=args.prune_dead_backward, share_recon=args.share_recon,
                 batch_generator=args.batch_g, split_dreal=args.split_dreal,
                 concurrent_generator=not args.no_concurrent_g, early_generator=args.early_g)
    tr.broadcast_parameters(0)
    import random
    torch.manual_seed(1000 + rank)                   # per-rank data / Z / T2 / crops
    random.seed(1000 + rank)
    pool = [torch.empty(B, 3, S, S).uniform_(-1, 1).pin_memory() for _ in range(2)]
    resident = [p.to(dev) for p in pool]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: the first step takes the lazy-R1 branch so every code path has run once
    warm_ids = [targs.d_reg_every] + list(range(1, args.warmup))
    for i, it in enumerate(warm_ids[:max(args.warmup, 1)]):
        tr.step(resident[i % 2], it)
    barrier()

    def timed(e2e: bool):
        sampler = ClockSampler(local)
        l0 = _lib.launch_count() + tr.replayed_launches
        d2h = 0
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # e2e: every step copies its batch from pinned host memory and its loss scalars back to pinned host
        # memory; the host consumes step k-1's losses while step k runs (the asynchronous logging a training
        # loop does), and the last step's losses before the clock stops
        host_losses = [torch.zeros(16).pin_memory() for _ in range(2)]
        copied = [torch.cuda.Event() for _ in range(2)]
        seen = 0.0
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        with sampler:
            ev0.record()
            for k in range(args.steps):
                it = args.warmup + 1 + k                 # R1 falls in whenever it % 16 == 0
                marks[k].record()
                if e2e:
                    X = pool[k % 2].to(dev, non_blocking=True)
                    losses = tr.step(X, it)
                    vals = torch.stack([v.detach().reshape(()) for v in losses.values()])
                    host_losses[k % 2][:vals.numel()].copy_(vals, non_blocking=True)
                    copied[k % 2].record()
                    d2h = vals.numel() * 4
                    if k > 0:
                        copied[(k - 1) % 2].synchronize()
                        seen += float(host_losses[(k - 1) % 2][0])
                else:
                    tr.step(resident[k % 2], it)
            if e2e:
                copied[(args.steps - 1) % 2].synchronize()
                seen += float(host_losses[(args.steps - 1) % 2][0])
            marks[args.steps].record()
            ev1.record()
            barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_step = [marks[k].elapsed_time(marks[k + 1]) for k in range(args.steps)]
        return float(ms.item()), _lib.launch_count() + tr.replayed_launches - l0, sampler.summary(), d2h, per_step

    ms, launches, clocks, _, per_step = timed(False)
    value = B * world * args.steps / (ms * 1e-3)
    r1_steps = [k for k in range(args.steps) if (args.warmup + 1 + k) % targs.d_reg_every == 0]
    r1_in_window = len(r1_steps)
    ms_r1 = [per_step[k] for k in r1_steps]
    ms_plain = [per_step[k] for k in range(args.steps) if k not in r1_steps]
    mean = lambda v: (sum(v) / len(v)) if v else None  # noqa: E731
    # amortised over the loop's 16-iteration lazy-R1 period (train.py:105): 15 plain + 1 R1 iteration
    amort = (15 * mean(ms_plain) + mean(ms_r1)) / 16 if (ms_r1 and ms_plain) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if _lib.umma_enabled() else "f32", "data": "synthetic",
        "config": {"workload": workload_name(B, S), "global_batch": B * world, "parallelism": f"dp{world}",
                   "r1_iterations_in_window": r1_in_window, "ms_per_step_plain": mean(ms_plain), "ms_per_step_r1": mean(ms_r1),
                   "ms_per_step_amortised_16": amort, "cuda_graphs": not args.no_graphs,
                   "second_backward": "Ex parameters only (pruned)" if args.prune_dead_backward else "full graph, as train.py:214-216",
                   "recon_forward": ("E(X), G(E(X)) once per iteration (share_recon)" if args.share_recon
                                     else "once per phase, as train.py:58,66,145,154"),
                   "l2": "activations of one step are tens of GB, far larger than the 126 MB L2; no explicit flush",
                   "est_tflops": FLOP_PER_IMAGE_STEP * value / 1e12 if S == 256 else None},
        "clocks": clocks, "gpu_launches": launches,
    }
    if not args.no_e2e:
        ms2, _, _, d2h, _ = timed(True)
        line["e2e"] = {"value": B * world * args.steps / (ms2 * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": B * 3 * S * S * 4, "d2h_bytes_per_step": d2h,
                       "api": "ideas_b200.train_step.Trainer.step"}
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    line["config"]["peak_mem_gib"] = round(peak_mem, 2)
    if world > 1:
        # every rank started from rank 0's weights and applied the same all-reduced gradients: the replicas must be
        # bit-identical after the run
        cs = tr.replica_checksum()
        gathered = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(gathered, cs)
        line["replicas_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
        line["config"]["grad_allreduce"] = "in-place flat fp32 bucket per optimiser (grads are views), NCCL sum / world"
        dist.barrier()
    if rank == 0:
        if world == 1:
            del tr
            torch.cuda.empty_cache()
            if roof is not None:
                line.update(roof)
                sustained_tf32_gemm(line)
            if not args.no_library_baseline:
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import library_baseline
                line["library_baseline"] = library_baseline.run(line)
            if not args.no_cpu_baseline:
                t, cores, n, kind = cpu_reference_step_time(S, 2, 0, budget_s=30.0)
                line["cpu_baseline"] = {"value": CPU_BATCH / t, "unit": UNIT, "cores": cores, "kind": kind,
                                        "sample": f"{n} train iteration(s) on {CPU_BATCH} images {S}x{S} "
                                                  f"({'unmodified reference' if kind == 'reference' else 'oracle port of train.py:33-221'}), {t:.1f} s each"}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        # Orderly teardown: the captured graphs hold NCCL work, so they go first, then the communicator.  A watchdog
        # still ends the process if the teardown blocks (round 1 observed destroy_process_group hanging with live
        # graphs) -- everything has been measured and printed by now.
        sys.stdout.flush()
        sys.stderr.flush()
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        tr.close()
        del tr
        torch.cuda.synchronize()
        dist.destroy_process_group()
        watchdog.cancel()


def main():
    args = parse()
    if args.case:
        run_case(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
