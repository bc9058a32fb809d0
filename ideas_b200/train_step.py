"""One IDEAS training iteration on the B200 networks, single- or multi-GPU.

Mirrors the loop body of the reference's train.py:33-221 statement for statement: D phase
(:48-102), lazy R1 every ``d_reg_every`` iterations (:105-129), G/E/Ex phase with its two
backwards over one retained graph (:135-216), EMA (:218-221); the three Adam optimisers
and their lazy-regularisation-corrected hyper-parameters are those of train.py:417-432.

Data parallelism (new; the reference is single-process): every rank runs the
same step on its own slice of the batch with its own Z / T2 / crops; gradients are averaged
with ONE flat NCCL all-reduce per optimiser step, issued from an ``optimizer.step`` pre-hook
so the training loop itself never mentions the process group.  Stock DDP is not used: the
second backward without a forward (train.py:214-216) is outside its contract.
"""
from __future__ import annotations

import argparse
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import optim
from torch.nn import functional as F

from . import _lib
from .models import init_model
from .utils import (accumulate, d_logistic_loss, d_r1_loss, draw_crops, g_nonsaturating_loss, patchify_image,
                    requires_grad)

_AB_NO_DCO_HOIST = __import__("os").environ.get("IDEAS_AB_NO_DCO_HOIST", "0") == "1"     # measurement-only A/B switch

NET_CLASSES = [("E", "DisentanglementEncoder"), ("G", "Generator"), ("Gstru", "StructureGenerator"),
               ("Ex", "TensorExtractor"), ("Dreal", "ImageLevelDiscriminator"),
               ("Dco", "CooccurenceDiscriminator"), ("Ddist", "DistributionDiscriminator")]
EMA_KEYS = ("E", "G", "Gstru", "Ex")


def default_args(**overrides) -> argparse.Namespace:
    """Hyper-parameters with the reference's defaults (train.py:331-370)."""
    d = dict(N=1, lambda_Ex=10.0, lr=0.002, batch_size=1, image_size=256, real_r1=10.0, texture_r1=1.0, dist_r1=1.0,
             ref_crop=4, n_crop=8, d_reg_every=16, channel=32, channel_multiplier=1, structure_channel=8,
             texture_channel=2048, num_iters=80000, start_iter=0, blur_kernel=(1, 3, 3, 1))
    d.update(overrides)
    return argparse.Namespace(**d)


_NVTX = __import__("os").environ.get("IDEAS_NVTX", "0") == "1"
_nvtx_open = [False]


def _nvtx_phase(name):
    """IDEAS_NVTX=1: NVTX ranges around the phases of an eager training iteration (discriminators / lazy R1 /
    encoder-generators-extractor), for timeline tools.  ``None`` closes the open range."""
    if not _NVTX:
        return
    if _nvtx_open[0]:
        torch.cuda.nvtx.range_pop()
        _nvtx_open[0] = False
    if name is not None:
        torch.cuda.nvtx.range_push("ideas: " + name)
        _nvtx_open[0] = True


class FlatGradAllReduce:
    """Average the gradients of one optimiser's parameters across ranks with a single all-reduce on a flat fp32
    bucket (d-group 182 MB, g-group 257 MB, ex-group 1.5 MB).

    ``bind()`` (called by the Trainer when world_size > 1) makes every ``p.grad`` a VIEW into the bucket: autograd
    accumulates straight into it, ``zero()`` clears it with one fill, and the all-reduce runs in place -- no
    gather / scatter copies around the collective (1.8 GB of HBM traffic per iteration before).  Unbound, the
    gradients are copied in and out of a scratch bucket (any list of parameters, grads may be None)."""

    def __init__(self, params: List[torch.Tensor], group=None):
        self.params = [p for p in params]
        self.group = group
        self.flat: Optional[torch.Tensor] = None
        self.bound = False
        self.calls = 0

    def bind(self):
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.bound = True

    def zero(self) -> bool:
        """Clear the bound bucket (and with it every p.grad); False when unbound (caller uses optimizer.zero_grad)."""
        if not self.bound:
            return False
        self.flat.zero_()
        return True

    def __call__(self, optimizer=None, args=None, kwargs=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        if self.bound:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)
            self.calls += 1
            return
        ps = [p for p in self.params if p.grad is not None]
        if not ps:
            return
        n = sum(p.numel() for p in ps)
        if self.flat is None or self.flat.numel() != n or self.flat.device != ps[0].device:
            self.flat = torch.empty(n, dtype=torch.float32, device=ps[0].device)
        off = 0
        views = []
        for p in ps:
            v = self.flat[off:off + p.numel()].view_as(p.grad)
            views.append(v)
            off += p.numel()
        torch._foreach_copy_(views, [p.grad for p in ps])
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        torch._foreach_copy_([p.grad for p in ps], views)
        self.calls += 1


class Trainer:
    """The ``trainer`` dict of train.py:390-432 as an object (``trainer[key]`` still works)."""

    def __init__(self, args: argparse.Namespace, device="cuda", seed: Optional[int] = None, fused_adam: bool = True,
                 states: Optional[Dict[str, dict]] = None, cuda_graphs: bool = False, multi_stream: Optional[bool] = None,
                 prune_dead_backward: bool = False, batch_generator: bool = False, split_dreal: bool = False,
                 concurrent_generator: bool = True, early_generator: bool = False, share_recon: bool = False):
        self.args = args
        # train.py evaluates S1, T1 = E(X) and hat_X1 = G(S1, T1) twice per iteration -- at :58,66 for the
        # discriminators and again at :145,154 for the generators -- with the same weights (only d_optim steps in
        # between) and the same input, i.e. with identical results.  With this flag they are evaluated ONCE, with
        # their autograd graph, at the first place; the discriminator phase reads detached copies and the generator
        # phase back-propagates through the graph kept from the first.  Same losses, same parameter trajectory, one
        # E forward and one G forward less per iteration.  Off by default: bench.py measures the loop as written.
        self.share_recon = bool(share_recon)
        self.batch_generator = bool(batch_generator)
        self.concurrent_generator = bool(concurrent_generator)
        self.early_generator = bool(early_generator)
        self.device = torch.device(device)
        self.cuda_graphs = bool(cuda_graphs and self.device.type == "cuda")
        # independent sub-graphs of one iteration (Dreal on the real batch, the co-occurrence branch, the
        # E(container)->Ex branch) run on side streams: their HBM-bound kernels overlap the tensor-bound ones of
        # the main branch, and autograd replays the same streams in backward.  Same arithmetic, same results.
        self.multi_stream = bool((self.cuda_graphs if multi_stream is None else multi_stream) and self.device.type == "cuda")
        self._side_streams: List[torch.cuda.Stream] = []
        self.split_dreal = bool(split_dreal) and not self.batch_generator
        # train.py:214-216 runs Ex_loss.backward() over the whole retained graph, although only ex_optim.step()
        # follows: the gradients it adds to E / G / Gstru are discarded by the next g_optim.zero_grad().  With this
        # flag the second backward is restricted to Ex's parameters -- same parameter trajectory, ~10 % fewer FLOPs
        # (SURVEY.md App. B).  Off by default: the .grad fields of E / G / Gstru then differ from the reference's
        # between iterations, and bench.py measures the reference's loop as written.
        self.prune_dead_backward = bool(prune_dead_backward)
        self._graphs: Dict[tuple, tuple] = {}
        self.replayed_launches = 0        # library kernels launched through graph replays (see bench.py gpu_launches)
        self._pool = None
        self._static_X: Optional[torch.Tensor] = None
        self._static_boxes: Optional[Dict[str, torch.Tensor]] = None
        self._host_boxes: Optional[List[Dict[str, torch.Tensor]]] = None
        if seed is not None:
            torch.manual_seed(seed)
        self.nets: Dict[str, torch.nn.Module] = {}
        for key, cls in NET_CLASSES:                                    # same construction order as train.py:390-399
            self.nets[key] = init_model(cls, args)
        for key in EMA_KEYS:                                            # train.py:400-414
            cls = dict(NET_CLASSES)[key]
            self.nets[key + "_ema"] = init_model(cls, args).eval()
        if states is not None:
            for k, sd in states.items():
                self.nets[k].load_state_dict(sd)
        for k in self.nets:
            self.nets[k].to(self.device)
        for key in EMA_KEYS:
            accumulate(self.nets[key + "_ema"], self.nets[key], 0)
        P = lambda *ks: [p for k in ks for p in self.nets[k].parameters()]  # noqa: E731
        fused = bool(fused_adam and self.device.type == "cuda")
        kw = dict(fused=fused, capturable=self.cuda_graphs)
        self.g_optim = optim.Adam(P("E", "G", "Gstru"), lr=args.lr, betas=(0.0, 0.99), **kw)
        self.ex_optim = optim.Adam(P("Ex"), lr=args.lr, betas=(0.0, 0.99), **kw)
        r = args.d_reg_every / (args.d_reg_every + 1)
        self.d_optim = optim.Adam(P("Dreal", "Dco", "Ddist"), lr=args.lr * r, betas=(0.0 ** r, 0.99 ** r), **kw)
        self.reducers = {}
        from .stylegan2.op.conv import invalidate_step_cache
        for name, opt in (("g", self.g_optim), ("ex", self.ex_optim), ("d", self.d_optim)):
            red = FlatGradAllReduce([p for grp in opt.param_groups for p in grp["params"]])
            opt.register_step_pre_hook(red)
            # packed weights cached for the running iteration are stale once the optimiser has stepped
            opt.register_step_post_hook(lambda *_: invalidate_step_cache())
            self.reducers[name] = red
            if self.device.type == "cuda" and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                red.bind()
        self.accum = 0.5 ** (32 / (10 * 1000))                          # train.py:30

    def __getitem__(self, key):
        if key in self.nets:
            return self.nets[key]
        return {"g_optim": self.g_optim, "ex_optim": self.ex_optim, "d_optim": self.d_optim}[key]

    def keys(self):
        return list(self.nets) + ["g_optim", "ex_optim", "d_optim"]

    def broadcast_parameters(self, src: int = 0):
        """Make every rank start from rank ``src``'s weights (one flat broadcast per network)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        for net in self.nets.values():
            ps = [p.data for p in net.parameters()]
            flat = torch.cat([p.reshape(-1) for p in ps])
            dist.broadcast(flat, src)
            off = 0
            for p in ps:
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()

    # -------------------------------------------------------------------------------------
    def _rand_like_Z(self, X, draws, key, device_rng=False):
        """Z ~ U(-1,1) of shape (B, N, H/16, W/16) (train.py:60-61: the structure code of E is 1/16 resolution)."""
        if draws is not None and key in draws:
            return draws[key].to(self.device)
        shape = (X.shape[0], self.args.N, X.shape[2] // 16, X.shape[3] // 16)
        if device_rng:                     # CUDA-graph path: Philox on the device (no H2D copy inside the graph)
            return torch.rand(shape, dtype=torch.float, device=self.device) * 2 - 1
        # train.py:60-61 draws on the CPU generator and copies; kept so seeded runs match the reference
        return torch.rand(size=shape, dtype=torch.float).to(self.device) * 2 - 1

    def _rand_like_T(self, T, draws, key):
        if draws is not None and key in draws:
            return draws[key].to(self.device)
        return torch.rand_like(T) * 2 - 1

    # ------------------------------------------------------------------------------------- side streams
    def _fork(self, idx: int, *inputs):
        """Context manager: run the body on side stream ``idx`` after everything enqueued so far on the current
        stream.  ``inputs`` are tensors produced on the current stream that the body reads."""
        import contextlib
        if not self.multi_stream:
            return contextlib.nullcontext()
        while len(self._side_streams) <= idx:
            self._side_streams.append(torch.cuda.Stream(device=self.device))
        side = self._side_streams[idx]
        side.wait_stream(torch.cuda.current_stream(self.device))
        for t in inputs:
            t.record_stream(side)
        return torch.cuda.stream(side)

    def _join(self, idx: int, *outputs):
        """Make the current stream wait for side stream ``idx``; ``outputs`` were produced there."""
        if not self.multi_stream:
            return
        main = torch.cuda.current_stream(self.device)
        main.wait_stream(self._side_streams[idx])
        for t in outputs:
            t.record_stream(main)

    # box sets one iteration needs: name -> number of crops (train.py:80-82,168-169)
    def _box_specs(self):
        a = self.args
        return {"fake_crops_d": a.n_crop, "real_crops_d": a.n_crop, "ref_crops_d": a.ref_crop * a.n_crop,
                "fake_crops_g": a.n_crop, "ref_crops_g": a.ref_crop * a.n_crop}

    def step(self, X: torch.Tensor, iter_idx: int, draws: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        """One iteration on batch X (already on the device).  Returns the loss tensors (device
        scalars; ``.item()`` them only when logging, as train.py:224-232 does).  With
        ``cuda_graphs=True`` (and no explicit ``draws``) the whole iteration is replayed as one CUDA
        graph per variant (with / without lazy R1, container switch)."""
        a = self.args
        r1 = iter_idx % a.d_reg_every == 0
        late = iter_idx > a.num_iters * 0.8
        if self.cuda_graphs and draws is None:
            return self._step_graphed(X, r1, late)
        return self._step_eager(X, r1, late, draws)

    # ------------------------------------------------------------------------------------- CUDA graphs
    def _refresh_boxes(self, H, W):
        """Draw this iteration's crop boxes like the reference (CPU RNG) and stage them into the static
        device tensors the captured patchify kernels read."""
        # two pinned staging sets used alternately: set i is rewritten only after the H2D copies issued from it two
        # steps ago have completed (event), so a queued copy can never observe the next iteration's boxes
        slot = self._box_slot
        self._box_slot ^= 1
        if self._box_events[slot] is not None:
            self._box_events[slot].synchronize()
        for key, n in self._box_specs().items():
            host = self._host_boxes[slot][key]
            host.copy_(torch.tensor(draw_crops(n, H, W), dtype=torch.int32))
            self._static_boxes[key].copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._box_events[slot] = ev

    def _snapshot_state(self):
        nets = {k: [t.detach().clone() for t in list(n.parameters()) + list(n.buffers())] for k, n in self.nets.items()}
        opts = []
        for opt in (self.g_optim, self.ex_optim, self.d_optim):
            opts.append({p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                         for p, st in opt.state.items()})
        return nets, opts

    @torch.no_grad()
    def _restore_state(self, snap):
        nets, opts = snap
        for k, n in self.nets.items():
            torch._foreach_copy_(list(n.parameters()) + list(n.buffers()), nets[k])
        for opt, saved in zip((self.g_optim, self.ex_optim, self.d_optim), opts):
            for p, st in opt.state.items():
                for name, v in st.items():
                    if not torch.is_tensor(v):
                        continue
                    if p in saved and name in saved[p]:
                        v.copy_(saved[p][name])
                    else:
                        v.zero_()          # state created by the warm-up: back to Adam's lazily-initialised zeros

    def _step_graphed(self, X, r1, late):
        H, W = X.shape[2], X.shape[3]
        if self._static_X is None:
            self._static_X = torch.empty_like(X)
            self._static_boxes = {k: torch.zeros((n, 4), dtype=torch.int32, device=self.device)
                                  for k, n in self._box_specs().items()}
            self._host_boxes = [{k: torch.zeros((n, 4), dtype=torch.int32).pin_memory()
                                 for k, n in self._box_specs().items()} for _ in range(2)]
            self._box_slot, self._box_events = 0, [None, None]
            self._static_X.copy_(X)
            self._refresh_boxes(H, W)
            # warm-up outside capture (lazy initialisation: Adam state, cuBLAS workspaces, kernel attributes,
            # the flat all-reduce buckets): two eager iterations of the R1 variant, which covers every code path
            # The warm-up iterations are throw-away: parameters, buffers, EMA copies and Adam state are put back
            # afterwards, so the first replayed graph is iteration 1 of the trajectory, exactly as in the eager loop.
            snap = self._snapshot_state()
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step_eager(self._static_X, True, late, None, boxes=self._static_boxes, device_rng=True)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self._restore_state(snap)
            del snap
            for name in ("g", "ex", "d"):
                self._zero_grad(name)
            torch.cuda.empty_cache()
            self._pool = torch.cuda.graph_pool_handle()
        key = (bool(r1), bool(late))
        if key not in self._graphs:
            self._static_X.copy_(X)
            for name in ("g", "ex", "d"):
                self._zero_grad(name)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph, pool=self._pool):
                losses = self._step_eager(self._static_X, r1, late, None, boxes=self._static_boxes, device_rng=True)
                losses = {k: v.detach() for k, v in losses.items()}
            self._graphs[key] = (graph, losses, _lib.launch_count() - n0)
        graph, losses, n_kernels = self._graphs[key]
        self._static_X.copy_(X, non_blocking=True)
        self._refresh_boxes(H, W)
        graph.replay()
        self.replayed_launches += n_kernels
        return losses

    # ------------------------------------------------------------------------------------- eager core
    def _step_eager(self, X, r1, late, draws, boxes=None, device_rng=False) -> Dict[str, torch.Tensor]:
        from .stylegan2.op.conv import step_scope
        with step_scope():                 # packed weights are built once per optimiser step, not once per call
            return self._iteration(X, r1, late, draws, boxes, device_rng)

    def _generate3(self, S1, S2, T1, T2, streams=(2, 3), x1=None, first=None):
        """hat_X1, hat_X2, hat_X3 = G(S1,T1), G(S2,T1), G(S2,T2) (train.py:66-70,154-158) and their concatenation
        (the argument of Dreal, train.py:73,161).  G is per-sample independent (no batch statistics), so the three
        calls may run as ONE call on the concatenated batch (``batch_generator=True``: same values and gradients, a
        third of the launches).  Measured on B200 at batch 32 it is 9 % SLOWER per step (334.7 vs 305.8 ms): at
        three times the working set the activations and weight slabs of consecutive kernels no longer meet in the
        126 MB L2, and the E(container) branch can no longer start under the third call.  Off by default."""
        # ``x1``: hat_X1 already evaluated (share_recon); ``first``: callable evaluating it (with its own autograd
        # mode) in the slot of the first call
        g1 = (lambda: x1) if x1 is not None else (first if first is not None else (lambda: self.nets["G"](S1, T1)))
        if not self.batch_generator or x1 is not None or first is not None:
            if self.concurrent_generator and self.multi_stream:
                # the three calls are independent chains of tensor-bound and HBM-bound kernels: on three streams the
                # blur / activation kernels of one call run under the convolutions of another (autograd replays the
                # same streams in backward).  Measured: 299.6 vs 301.1 ms per step.
                with self._fork(streams[0], S2, T1):
                    x2 = self.nets["G"](S2, T1)
                with self._fork(streams[1], S2, T2):
                    x3 = self.nets["G"](S2, T2)
                x1 = g1()
                self._join(streams[0], x2)
                self._join(streams[1], x3)
                xs = (x1, x2, x3)
            else:
                xs = g1(), self.nets["G"](S2, T1), self.nets["G"](S2, T2)
            return xs, (None if self.split_dreal else torch.cat(xs, 0))
        hat = self.nets["G"](torch.cat((S1, S2, S2), 0), torch.cat((T1, T1, T2), 0))
        return hat.chunk(3, 0), hat

    def _zero_grad(self, name: str, set_to_none: bool = True):
        """optimizer.zero_grad() of train.py:100,122,209,214; one fill of the flat bucket when the grads are views."""
        if not self.reducers[name].zero():
            {"g": self.g_optim, "ex": self.ex_optim, "d": self.d_optim}[name].zero_grad(set_to_none=set_to_none)

    def close(self):
        """Drop the captured CUDA graphs (they hold NCCL work and private memory pools); call before tearing the
        process group down."""
        self._graphs.clear()
        self._pool = None
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def replica_checksum(self) -> torch.Tensor:
        """fp64 sums of every network's parameters: equal on all ranks iff the replicas are bit-identical."""
        return torch.stack([torch.stack([p.detach().double().sum() for p in n.parameters()]).sum() for n in self.nets.values()])

    def _dreal_fake(self, x1, x2, x3, x_all):
        """Dreal(cat(hat_X1, hat_X2, hat_X3)) (train.py:73,161).  Dreal is per-sample independent (no minibatch-stddev
        layer, models.py:369-376), so with side streams the three fake batches each continue on the stream of the
        Generator call that produced them: five chains (three G -> Dreal, E(container) -> Ex, the co-occurrence
        branch) run at the same time.  Measured: no gain over one call on the concatenated batch (285.9 vs 285.6 ms per
        step; a 265 ms reading of one run did not reproduce), so ``split_dreal`` is off by default."""
        if x_all is not None:
            return self.nets["Dreal"](x_all)
        if self.concurrent_generator and self.multi_stream:
            with self._fork(2, x2):
                p2 = self.nets["Dreal"](x2)
            with self._fork(3, x3):
                p3 = self.nets["Dreal"](x3)
            p1 = self.nets["Dreal"](x1)
            self._join(2, p2)
            self._join(3, p3)
            return torch.cat((p1, p2, p3), 0)
        return torch.cat([self.nets["Dreal"](x) for x in (x1, x2, x3)], 0)

    def _iteration(self, X, r1, late, draws, boxes, device_rng) -> Dict[str, torch.Tensor]:
        try:
            return self._iteration_body(X, r1, late, draws, boxes, device_rng)
        finally:
            _nvtx_phase(None)

    def _iteration_body(self, X, r1, late, draws, boxes, device_rng) -> Dict[str, torch.Tensor]:
        a, t = self.args, self.nets
        H, W = X.shape[2], X.shape[3]

        def crops(key, n):
            if boxes is not None:
                return boxes[key]
            return draws[key] if (draws is not None and key in draws) else draw_crops(n, H, W)

        loss = {}
        _nvtx_phase("discriminators")
        # ---------------- discriminators (train.py:48-102)
        for k in EMA_KEYS:
            requires_grad(t[k], False)
        for k in ("Dreal", "Dco", "Ddist"):
            requires_grad(t[k], True)
        # host-side draws in the reference's order (train.py:60-61 then :80-82) before any branch uses them
        Z = self._rand_like_Z(X, draws, "Z_d", device_rng)
        fake_boxes = crops("fake_crops_d", a.n_crop)
        real_boxes = crops("real_crops_d", a.n_crop)
        ref_boxes = crops("ref_crops_d", a.ref_crop * a.n_crop)
        with self._fork(0, X):                                   # Dreal on the real batch needs nothing from E / G
            real_pred = t["Dreal"](X)
        with self._fork(1, X):                                   # so do the real / reference patches of Dco
            real_patch = patchify_image(X, a.n_crop, crops=real_boxes)
            ref_patch = patchify_image(X, a.ref_crop * a.n_crop, crops=ref_boxes)
            # train.py:85-86 evaluates the fake patches first and reuses ref_input for the real ones; ref_input
            # does not depend on the first argument, so the order of the two calls is immaterial
            real_texture_pred, ref_input = t["Dco"](real_patch, ref_patch, ref_batch=a.ref_crop)
        shared = None
        if self.share_recon:
            # E(X) and G(E(X)) once for both phases: built here WITH their graph, read here through detached copies
            requires_grad(t["E"], True)
            S1g, T1g = t["E"](X)
            requires_grad(t["E"], False)
            S1, T1 = S1g.detach(), T1g.detach()
            shared = dict(S1=S1g, T1=T1g)

            def recon():
                requires_grad(t["G"], True)
                shared["hat_X1"] = t["G"](S1g, T1g)
                requires_grad(t["G"], False)
                return shared["hat_X1"].detach()
        else:
            S1, T1 = t["E"](X)
        S2 = t["Gstru"](Z)
        T2 = self._rand_like_T(T1, draws, "T2_d")
        (hat_X1, hat_X2, hat_X3), hat_all = self._generate3(S1, S2, T1, T2, first=recon if self.share_recon else None)
        fake_pred = self._dreal_fake(hat_X1, hat_X2, hat_X3, hat_all)
        fake_patch = patchify_image(hat_X2, a.n_crop, crops=fake_boxes)
        self._join(1, real_texture_pred, ref_input, real_patch, ref_patch)
        fake_texture_pred, _ = t["Dco"](fake_patch, ref_input=ref_input)
        self._join(0, real_pred)
        D_real_loss = d_logistic_loss(real_pred, fake_pred)
        D_texture_loss = d_logistic_loss(real_texture_pred, fake_texture_pred)
        D_dist_loss = d_logistic_loss(t["Ddist"](T2), t["Ddist"](T1))
        loss.update(D_real_loss=D_real_loss, D_texture_loss=D_texture_loss, D_dist_loss=D_dist_loss)

        def generator_draws():
            """Host-side draws of the G / E / Ex phase in the reference's order (train.py:147-148 then :168-169)."""
            return (self._rand_like_Z(X, draws, "Z_g", device_rng), crops("fake_crops_g", a.n_crop),
                    crops("ref_crops_g", a.ref_crop * a.n_crop))

        def generator_forward(g_streams, drawn):
            """First part of the G / E / Ex phase (train.py:144-158): everything that does not read a discriminator."""
            Zg, fb, rb = drawn
            S1g, T1g = (shared["S1"], shared["T1"]) if shared is not None else t["E"](X)
            S2g = t["Gstru"](Zg)
            T2g = self._rand_like_T(T1g, draws, "T2_g")
            hats, hall = self._generate3(S1g, S2g, T1g, T2g, streams=g_streams,
                                         x1=shared["hat_X1"] if shared is not None else None)
            return dict(Z=Zg, fake_boxes=fb, ref_boxes=rb, S1=S1g, T1=T1g, S2=S2g, T2=T2g, hats=hats, hat_all=hall)

        early = None
        if self.early_generator and self.multi_stream and not r1 and not self.share_recon:
            # The generator-side forward of the next phase reads no discriminator, and E / G / Gstru do not change
            # in this one: start it on its own streams now, under the discriminators' backward pass and optimiser
            # step.  Same values, same order of random draws (the backward pass draws nothing).  Measured: no gain
            # (285.0 vs 284.8 ms per step), so it is off by default.
            for k in EMA_KEYS:
                requires_grad(t[k], True)
            drawn = generator_draws()
            with self._fork(4, X, drawn[0]):
                early = generator_forward((5, 6), drawn)
        self._zero_grad("d")
        (D_real_loss + D_texture_loss + D_dist_loss).backward()
        self.d_optim.step()
        # ---------------- lazy R1 (train.py:105-129)
        if r1:
            _nvtx_phase("lazy R1")
            Xr = X.detach().requires_grad_(True)
            r1_real = d_r1_loss(t["Dreal"](Xr), Xr)
            rp = real_patch.detach().requires_grad_(True)
            pred, _ = t["Dco"](rp, ref_patch, ref_batch=a.ref_crop)
            r1_tex = d_r1_loss(pred, rp)
            T2r = T2.detach().requires_grad_(True)
            r1_dist = d_r1_loss(t["Ddist"](T2r), T2r)
            self._zero_grad("d")
            r1_sum = a.real_r1 / 3 * r1_real * a.d_reg_every
            r1_sum = r1_sum + a.texture_r1 / 3 * r1_tex * a.d_reg_every
            r1_sum = r1_sum + a.dist_r1 / 3 * r1_dist * a.d_reg_every
            r1_sum.backward()
            self.d_optim.step()
            loss.update(D_real_r1_loss=r1_real, D_texture_r1_loss=r1_tex, D_dist_r1_loss=r1_dist)
        # ---------------- encoder / generators / extractor (train.py:135-216)
        _nvtx_phase("encoder / generators / extractor")
        for k in EMA_KEYS:
            requires_grad(t[k], True)
        for k in ("Dreal", "Dco", "Ddist"):
            requires_grad(t[k], False)
        def reference_code(rb):
            # the reference-patch code of the co-occurrence branch depends on X and the (now updated, frozen)
            # discriminator alone and has no backward: evaluate it on the branch's stream before the Generator calls
            with self._fork(1, X), torch.no_grad():
                return t["Dco"].reference_code(patchify_image(X, a.ref_crop * a.n_crop, crops=rb), a.ref_crop)

        hoist = self.multi_stream and not _AB_NO_DCO_HOIST
        if early is None:
            drawn = generator_draws()
            ref_input_g = reference_code(drawn[2]) if hoist else None
            gf = generator_forward((2, 3), drawn)
        else:
            gf = early
            ref_input_g = reference_code(gf["ref_boxes"]) if hoist else None
            self._join(4, gf["S1"], gf["T1"], gf["S2"], gf["T2"], gf["Z"], *gf["hats"])
        Z, fake_boxes, ref_boxes = gf["Z"], gf["fake_boxes"], gf["ref_boxes"]
        S1, T1, S2, T2, hat_all = gf["S1"], gf["T1"], gf["S2"], gf["T2"], gf["hat_all"]
        hat_X1, hat_X2, hat_X3 = gf["hats"]
        container = hat_X3 if late else hat_X2
        with self._fork(0, container, S2, Z):                    # E(container) -> Ex branch (train.py:178-189)
            hat_S2, _ = t["E"](container)
            E_stru_loss = F.l1_loss(hat_S2, S2)
            Ex_loss = F.l1_loss(t["Ex"](hat_S2), Z)
        with self._fork(1, hat_X2, X):                           # co-occurrence branch (train.py:168-175)
            fake_patch = patchify_image(hat_X2, a.n_crop, crops=fake_boxes)
            if ref_input_g is None:
                ref_patch = patchify_image(X, a.ref_crop * a.n_crop, crops=ref_boxes)
                fake_patch_pred, _ = t["Dco"](fake_patch, ref_patch, ref_batch=a.ref_crop)
            else:                                                # same stream as the early evaluation above
                fake_patch_pred, _ = t["Dco"](fake_patch, ref_input=ref_input_g)
            G_texture_loss = g_nonsaturating_loss(fake_patch_pred)
        G_rec_loss = F.l1_loss(hat_X1, X)
        G_real_loss = g_nonsaturating_loss(self._dreal_fake(hat_X1, hat_X2, hat_X3, hat_all))
        E_dist_loss = g_nonsaturating_loss(t["Ddist"](T1))
        self._join(0, E_stru_loss, Ex_loss)
        self._join(1, G_texture_loss)
        Loss_total = (G_rec_loss + G_texture_loss + 2 * G_real_loss) + (E_dist_loss + E_stru_loss) + a.lambda_Ex * Ex_loss
        self._zero_grad("g")
        Loss_total.backward(retain_graph=True)
        self.g_optim.step()
        self._zero_grad("ex")
        if self.prune_dead_backward:
            Ex_loss.backward(inputs=[p for p in t["Ex"].parameters()])
        else:
            Ex_loss.backward()
        self.ex_optim.step()
        for k in EMA_KEYS:
            accumulate(t[k + "_ema"], t[k], self.accum)
        loss.update(G_rec_loss=G_rec_loss, G_real_loss=G_real_loss, G_texture_loss=G_texture_loss,
                    E_dist_loss=E_dist_loss, E_stru_loss=E_stru_loss, Ex_loss=Ex_loss, Loss_total=Loss_total)
        return loss
