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
    ap.add_argument("--no-graphs", action="store_true", help="run the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--prune-dead-backward", action="store_true",
                    help="NOT the default measurement: restrict the loop's second backward (train.py:214-216) to Ex's parameters")
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
def cpu_reference_step_time(image_size, steps, warmup, budget_s=200.0):
    """The reference algorithm for the path (oracle port of train.py:33-221) on the host cores.
    One step = one training iteration on ONE synthetic image (bounded sample of the batch-32 workload)."""
    import torch
    from oracle.train_step import OracleTrainer
    torch.set_num_threads(os.cpu_count() or 1)
    tr = OracleTrainer(seed=0, image_size=image_size)
    torch.manual_seed(1)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        X = torch.rand(1, 3, image_size, image_size) * 2 - 1
        t0 = time.perf_counter()
        tr.step(X, i + 1)          # iterations 1.. : lazy R1 falls in only when i+1 is a multiple of 16
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        # keep the whole arm within a few minutes: stop early (after >= 1 timed step) once the budget is spent
        if times and time.perf_counter() - t_start + dt > budget_s:
            break
    return sum(times) / len(times), torch.get_num_threads(), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, cores, timed = cpu_reference_step_time(args.image_size, args.steps, args.warmup)
    v = 1.0 / t
    sample = f"1 image per step ({args.image_size}x{args.image_size}), full train iteration on CPU, oracle port of train.py:33-221"
    print(json.dumps({
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": timed, "steps_requested": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(args.batch, args.image_size), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------
def time_kernel(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def roofline_sections(peaks, peak_kind):
    """Dominant kernels timed alone, live, with CUDA events on the launching (current) stream."""
    import torch
    from ideas_b200 import _lib
    from ideas_b200._tensor import ptr, stream_ptr
    dev = torch.device("cuda")
    out = {}
    # --- cfg 3: ModulatedConv2d 512->512 3x3 @64x64, batch 16: the implicit-GEMM forward launch.
    N, C, K, H = 16, 512, 512, 64
    x = torch.randn(N, H, H, C, device=dev)                       # NHWC, 134 MB (> 126 MB L2)
    wp = torch.randn(9, K, C, device=dev) / (C * 9) ** 0.5
    d = torch.rand(N, K, device=dev) + 0.5
    bias = torch.randn(K, device=dev)
    y = torch.empty(N, H, H, K, device=dev)
    impl = _lib.IMPL_AUTO

    def conv():
        _lib.call("ideas_conv2d_forward", ptr(y), ptr(x), ptr(wp), ptr(None), ptr(d), ptr(bias), N, H, H, C, K, 3, 3, 1, 1,
                  _lib.ACT_LRELU, 0.2, 2 ** 0.5, impl, stream_ptr(x))

    t = time_kernel(conv, iters=10, warm=2)
    flops = 2.0 * N * H * H * K * C * 9
    tf32_peak = peaks["bf16_tflops"] * 0.5
    umma = _lib.umma_enabled()
    out["roofline"] = {
        "kernel": "conv_igemm (modulated 3x3 conv fwd, 16x512x64x64 -> 512, demod+bias+lrelu fused)",
        "bound": "tensor", "achieved": flops / t / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
        "frac": flops / t / 1e12 / tf32_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r1b_conv_fwd_cfg3.raw.csv);
        # algorithmic bytes are 277 MB (x 134 + y 134 + packed weights 9.4): weights and part of x are L2 hits
        "traffic": 163231488 + 83914752,
        "peak_source": f"{peak_kind} MEASURED_PEAKS.json bf16_tflops {peaks['bf16_tflops']} x 0.5 (kind::tf32 issues at half the bf16 rate)",
        "frac_of_bf16_peak": flops / t / 1e12 / peaks["bf16_tflops"],
        "path": "tcgen05 kind::tf32" if umma else "fp32 FFMA (SIMT)", "algorithmic_flops_per_launch": flops,
        "ms_per_launch": t * 1e3,
    }
    # --- cfg-3 weight gradient (same shape)
    dy = torch.randn(N, H, H, K, device=dev)
    dwp = torch.zeros(9, K, C, device=dev)

    def wgrad():
        _lib.call("ideas_conv2d_wgrad", ptr(dwp), ptr(x), ptr(dy), ptr(None), ptr(None), N, H, H, C, K, 3, 3, 1, 1, H, H, impl,
                  stream_ptr(x))

    t = time_kernel(wgrad, iters=10, warm=2)
    out["roofline_wgrad"] = {"kernel": "conv_umma_wgrad (cfg-3 shape, 16x512x64x64, 512->512)", "bound": "tensor",
                             "achieved": flops / t / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                             "frac": flops / t / 1e12 / tf32_peak, "traffic": None, "ms_per_launch": t * 1e3}
    del x, y, dy
    # --- halo-reuse kernel on the widest low-channel layer of G / Dreal: 128 -> 128 at 256x256, batch 32
    N2, C2, H2 = 32, 128, 256
    x2 = torch.randn(N2, H2, H2, C2, device=dev)
    w2 = torch.randn(9, C2, C2, device=dev) / (C2 * 9) ** 0.5
    y2 = torch.empty(N2, H2, H2, C2, device=dev)

    def conv2():
        _lib.call("ideas_conv2d_forward", ptr(y2), ptr(x2), ptr(w2), ptr(None), ptr(None), ptr(bias[:C2]), N2, H2, H2, C2, C2, 3, 3,
                  1, 1, _lib.ACT_LRELU, 0.2, 2 ** 0.5, impl, stream_ptr(x2))

    t = time_kernel(conv2, iters=10, warm=2)
    f2 = 2.0 * N2 * H2 * H2 * C2 * C2 * 9
    out["roofline_halo"] = {"kernel": "conv_umma_halo (3x3 conv fwd, 32x128x256x256 -> 128, bias+lrelu fused)", "bound": "tensor",
                            "achieved": f2 / t / 1e12, "peak": tf32_peak, "unit": "TFLOP/s", "frac": f2 / t / 1e12 / tf32_peak,
                            # ncu --set full, profiles/r1b_conv_halo_g256.raw.csv; algorithmic 2.148 GB
                            "traffic": 1076674000 + 1030189000, "ms_per_launch": t * 1e3}
    del x2, y2
    # --- cfg 2: Blur pad (2,2), (32,128,256,256) -> (32,128,257,257)
    B, Cb, Hb = 32, 128, 256
    xb = torch.randn(B, Hb, Hb, Cb, device=dev)
    yb = torch.empty(B, Hb + 1, Hb + 1, Cb, device=dev)
    k = torch.tensor([1., 3., 3., 1.], device=dev)
    k = torch.outer(k, k)
    k = (k / k.sum()).contiguous()

    def blur():
        _lib.call("ideas_upfirdn2d", ptr(yb), ptr(xb), ptr(k), B, Hb, Hb, Cb, 4, 4, 1, 1, 1, 1, 2, 2, 2, 2, ptr(None), 0.2,
                  1.0, stream_ptr(xb))

    t = time_kernel(blur, iters=20, warm=3)
    nbytes = 4.0 * (xb.numel() + yb.numel())
    out["roofline_hbm"] = {
        "kernel": "blur4_nhwc (upfirdn2d up=down=1, 4x4, pad (2,2), 32x128x256x256)", "bound": "hbm",
        "achieved": nbytes / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": nbytes / t / 1e9 / peaks["hbm_gbs"],
        "traffic": 1173469000 + 1033594000,    # ncu --set full, profiles/r1b_blur.raw.csv
        "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs", "algorithmic_bytes_per_launch": nbytes,
        "ms_per_launch": t * 1e3,
    }
    return out


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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    B, S = args.batch, args.image_size
    targs = default_args(batch_size=B, image_size=S)
    tr = Trainer(targs, device=dev, seed=0, cuda_graphs=not args.no_graphs,   # same seed on every rank => identical replicas
                 prune_dead_backward=args.prune_dead_backward)
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
        with sampler:
            ev0.record()
            for k in range(args.steps):
                it = args.warmup + 1 + k                 # R1 falls in whenever it % 16 == 0
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
            ev1.record()
            barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), _lib.launch_count() + tr.replayed_launches - l0, sampler.summary(), d2h

    ms, launches, clocks, _ = timed(False)
    value = B * world * args.steps / (ms * 1e-3)
    r1_in_window = sum(1 for k in range(args.steps) if (args.warmup + 1 + k) % targs.d_reg_every == 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if _lib.umma_enabled() else "f32", "data": "synthetic",
        "config": {"workload": workload_name(B, S), "global_batch": B * world, "parallelism": f"dp{world}",
                   "r1_iterations_in_window": r1_in_window, "cuda_graphs": not args.no_graphs,
                   "second_backward": "Ex parameters only (pruned)" if args.prune_dead_backward else "full graph, as train.py:214-216",
                   "l2": "activations of one step are tens of GB, far larger than the 126 MB L2; no explicit flush",
                   "est_tflops": FLOP_PER_IMAGE_STEP * value / 1e12 if S == 256 else None},
        "clocks": clocks, "gpu_launches": launches,
    }
    if not args.no_e2e:
        ms2, _, _, d2h = timed(True)
        line["e2e"] = {"value": B * world * args.steps / (ms2 * 1e-3), "unit": UNIT,
                       "h2d_bytes_per_step": B * 3 * S * S * 4, "d2h_bytes_per_step": d2h,
                       "api": "ideas_b200.train_step.Trainer.step"}
    peak_mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    line["config"]["peak_mem_gib"] = round(peak_mem, 2)
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1:
            del tr
            torch.cuda.empty_cache()
            peaks, kind = load_peaks()
            if not args.no_roofline:
                line.update(roofline_sections(peaks, kind))
            if not args.no_cpu_baseline:
                t, cores, _ = cpu_reference_step_time(S, 1, 0)
                line["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"1 train iteration on 1 image {S}x{S} (oracle port of train.py:33-221), {t:.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured graphs hold NCCL work: tearing the communicator down while they are alive can block at
        # exit (observed: the JSON line printed, then destroy_process_group never returned).  Everything is
        # measured and printed; leave together and skip the teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
