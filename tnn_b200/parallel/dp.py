"""Data parallelism.

``FusedDataParallel`` is the product path; with ``overlap`` (default) the slab is cut
into gradient buckets (groups of top-level layers) and each bucket's fused reduce+update kernel is
queued on a side stream as soon as backward has produced it, so the NVLink traffic and the optimizer
math overlap the rest of backward.  Summary: parameters, gradients and bf16
compute copies of every replica live in ONE NVLink peer-mapped symmetric slab per GPU
(``torch.distributed._symmetric_memory`` provides the allocation + handle exchange, i.e. the
plumbing); each optimisation step is a single hand-written kernel
(csrc/kernels/optim.cu::k_dp_step) that barriers across ranks with ``.sys``-scope flags,
reduces this rank's 1/W gradient slice from all peers (peer loads, or one
``multimem.ld_reduce`` through NVSwitch when the slab is multicast-bound), applies the
scale + Adam/SGD update on that slice (optimizer state is sharded W ways), broadcasts the new
fp32 + bf16 parameters into every replica (peer stores / ``multimem.st``), barriers again and
zeroes its gradients.  No NCCL call and no separate scale / cast / optimizer / memset kernel
runs on that path.

``AllReduceDataParallel`` is the measured baseline: ``all_reduce`` (NCCL on GPUs, gloo in the CPU
tests) followed by the ordinary flat optimizer kernel.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from ..core import side_stream, stepstate
from ..nn.graph import ALIGN_ELEMS, Graph, graph_from_sequential
from ..nn.optimizers import SGD, Adam, Optimizer
from ..ops import native


def _aligned_numel(seq) -> int:
    total = 0
    for layer in seq.all_layers():
        for d in layer.param_descriptors():
            if d.trainable:
                n = 1
                for s in d.shape:
                    n *= s
                total += (n + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
    return max(total, ALIGN_ELEMS)


class AllReduceDataParallel:
    """Baseline: bucketless all-reduce of the contiguous gradient slab, then the flat optimizer."""

    def __init__(self, graph: Graph, optimizer: Optimizer, group=None):
        self.graph, self.optimizer, self.group = graph, optimizer, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if optimizer.ctx is None:
            optimizer.attach(graph.context())
        self._base_scale = None

    def step(self):
        opt = self.optimizer
        if self._base_scale is None:
            self._base_scale = opt.grad_scale
        if self.world > 1:
            g = self.graph.ctx.grad_slab
            # average (not sum): the optimizer's grad_scale lives in the device step state that TrainStep
            # pushed before the step, so the 1/world factor has to be applied to the gradients themselves
            if g.is_cuda:
                dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(g, group=self.group)
                g.mul_(1.0 / self.world)
        opt.update()

    def broadcast_parameters(self, src: int = 0):
        if self.world > 1:
            dist.broadcast(self.graph.ctx.param_slab, src, group=self.group)
            dist.broadcast(self.graph.ctx.buffer_slab, src, group=self.group)
            self.graph.ctx.refresh_lp()


class FusedDataParallel:
    """Fused all-reduce + optimizer over NVLink peer memory (see module docstring)."""

    FLAG_BYTES = 4096

    def __init__(self, optimizer: Optimizer, group=None, ctas: int | None = None, use_multicast: bool | None = None,
                 overlap: bool | None = None, bucket_elems: int | None = None):
        assert dist.is_initialized(), "init torch.distributed (nccl) first"
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.optimizer = optimizer
        self.ctas = int(ctas or os.environ.get("TNN_DP_CTAS", 64))     # x 256 threads, 4 independent vectors in flight per thread
        self.use_multicast = use_multicast
        self.graph: Graph | None = None
        self.epoch = 0
        self.launches = 0
        # overlap: gradient buckets (groups of top-level layers, in backward-completion order) are
        # reduced + updated by the same kernel on a SIDE stream while backward continues
        self.overlap = overlap if overlap is not None else os.environ.get("TNN_DP_OVERLAP", "1") == "1"
        self.bucket_elems = int(bucket_elems or os.environ.get("TNN_DP_BUCKET_ELEMS", 2 << 20))
        self.buckets: list[dict] = []
        self.side = None

    # ---------------------------------------------------------------- construction
    def build_graph(self, seq, device, seed=None, name=None) -> Graph:
        import torch.distributed._symmetric_memory as symm_mem

        n = _aligned_numel(seq)
        chunk = (n + self.world - 1) // self.world
        chunk = (chunk + 1023) // 1024 * 1024
        self.n, self.chunk = n, chunk
        npad = chunk * self.world                      # every rank's slice is fully backed
        self.off_p, self.off_g = 0, npad * 4
        self.off_lp = self.off_g + npad * 4
        self.off_flag = (self.off_lp + npad * 2 + 255) // 256 * 256
        total = self.off_flag + self.FLAG_BYTES
        self.own_symm = os.environ.get("TNN_DP_OWN_SYMM", "0") == "1"
        if self.own_symm:
            # the framework's own group (csrc/runtime/symmetric_memory.hpp): cuMemCreate slab per rank, POSIX-fd handles
            # over Unix sockets, NVLS multicast window created by rank 0.  Opt-in: it has not run on hardware yet.
            self.buf, ptrs, mc = self._own_symmetric(total, device)
            self.hdl = None
        else:
            self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.hdl = symm_mem.rendezvous(self.buf, self.group)
            ptrs = list(self.hdl.buffer_ptrs)
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        f32 = lambda off, cnt: self.buf[off:off + cnt * 4].view(torch.float32)  # noqa: E731
        self._p, self._g = f32(self.off_p, npad), f32(self.off_g, npad)
        self._lp = self.buf[self.off_lp:self.off_lp + npad * 2].view(torch.bfloat16)
        slabs = {torch.float32: [self._p[:n], self._g[:n]], torch.bfloat16: [self._lp[:n]]}

        def alloc(numel, dtype):
            t = slabs[dtype].pop(0)
            assert t.numel() >= numel
            return t[:numel] if numel != t.numel() else t

        self.graph = graph_from_sequential(seq, device, seed, name, alloc_fn=alloc)
        self.g_peers = [p + self.off_g for p in ptrs]
        self.p_peers = [p + self.off_p for p in ptrs]
        self.pb_peers = [p + self.off_lp for p in ptrs]
        self.flag_peers = [p + self.off_flag for p in ptrs]
        want_mc = self.use_multicast if self.use_multicast is not None else os.environ.get("TNN_DP_MULTICAST", "1") == "1"
        self.mc = mc if (mc and want_mc) else 0
        self.sync = torch.zeros(4, dtype=torch.int32, device=device)
        opt = self.optimizer
        opt.ctx = self.graph.context()
        self._plan_buckets(seq)
        nstate = max(chunk, sum(b["chunk"] for b in self.buckets)) if self.buckets else chunk
        self.m = torch.zeros(nstate, dtype=torch.float32, device=device)
        self.v = torch.zeros(nstate, dtype=torch.float32, device=device) if isinstance(opt, Adam) else None
        if self.overlap and self.buckets and device.type == "cuda":
            self.side = torch.cuda.Stream(device)
        dist.barrier(self.group)
        torch.cuda.synchronize(device)
        if seed is None:
            self.broadcast_parameters(0)      # unseeded replicas would start (and stay) different
            torch.cuda.synchronize(device)
        return self.graph

    def _own_symmetric(self, total: int, device):
        """(uint8 tensor over this rank's slab, [base pointer of every rank], multicast base or 0)."""
        import tempfile

        Cx = native.require()
        box = [tempfile.mkdtemp(prefix="tnn_symm_") if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0), group=self.group)
        mem = Cx.SymmetricMemory.rendezvous(self.rank, self.world, device.index if device.index is not None else torch.cuda.current_device(),
                                            total, box[0], True, 60000)
        self._own_mem = mem                               # owns the mappings: must outlive every view
        if self.rank == 0:
            try:
                os.rmdir(box[0])                          # every mailbox unlinked itself after the last phase
            except OSError:
                pass

        class _Slab:                                      # zero-copy view; keeps `mem` alive through the tensor
            def __init__(s):
                s._keep = mem
                s.__cuda_array_interface__ = {"shape": (total,), "typestr": "|u1", "data": (mem.base(self.rank), False), "version": 3,
                                              "strides": None}

        buf = torch.as_tensor(_Slab(), device=device)
        buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                          # nobody reads a peer's slab before it was cleared
        return buf, [mem.base(r) for r in range(self.world)], int(mem.multicast_base) if mem.has_multicast else 0

    # ---------------------------------------------------------------- gradient buckets
    def _plan_buckets(self, seq):
        """Contiguous slab ranges per group of top-level layers (the model is a Sequential whose
        parameters were packed in layer order), grouped from the END so that a bucket is complete as
        soon as backward has passed its first layer."""
        self.buckets = []
        top = getattr(seq, "layers", None)
        if not top or self.world < 2:
            return
        owner = {}
        for i, l in enumerate(top):
            for sub in l.all_layers():
                owner[id(sub)] = i
        ranges = {}
        for layer, d, off, numel in self.graph.ctx.entries:
            i = owner.get(id(layer))
            if i is None:
                return
            end = off + (numel + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
            lo, hi = ranges.get(i, (off, end))
            ranges[i] = (min(lo, off), max(hi, end))
        groups, hi, first, lo = [], None, None, None
        for i in sorted(ranges, reverse=True):          # last layer first = backward completion order
            lo_i, hi_i = ranges[i]
            if hi is None:
                hi = hi_i
            first, lo = i, lo_i
            if hi - lo >= self.bucket_elems:
                groups.append((first, lo, hi))
                hi = None
        if hi is not None:
            groups.append((first, lo, hi))
        # the groups must tile [0, n) exactly
        covered = sum(h - l for _, l, h in groups)
        if covered != self.n or len(groups) < 2:
            self.buckets = []
            return
        st = 0
        for first, lo, hi in groups:
            c = ((hi - lo + self.world - 1) // self.world + 3) // 4 * 4
            self.buckets.append(dict(first_layer=first, lo=lo, hi=hi, chunk=c, st_off=st))
            st += c

    def _launch(self, lo0=0, hi0=-1, chunk=None, st_off=0):
        opt = self.optimizer
        dev = self.graph.device()
        native.require().dp_fused_step(
            self.world, self.rank, self.n, chunk if chunk is not None else self.chunk, self.g_peers, self.p_peers, self.pb_peers,
            self.flag_peers, (self.mc + self.off_g) if self.mc else 0, (self.mc + self.off_p) if self.mc else 0,
            (self.mc + self.off_lp) if self.mc else 0, self.m, self.v, self.sync, stepstate.get(dev).dev,
            isinstance(opt, Adam), self.ctas, lo0, hi0, st_off)
        self.launches += 1

    # ---------------------------------------------------------------- per-bucket timing / roofline (eager passes only)
    def _timed_launch(self, b):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._launch(b["lo"], b["hi"], b["chunk"], b["st_off"])
        e1.record()
        self._timing.append((b, e0, e1))

    def bucket_report(self, eager_step, x, y, reps: int = 6) -> dict:
        """Runs ``reps`` EAGER steps (``eager_step(x, y)`` = TrainStep._eager) with CUDA events around every bucket
        kernel on the side stream and reports, max over ranks: microseconds per bucket (this includes the wait for
        the slowest rank at the bucket's barrier), the bytes this rank moves over NVLink / HBM for it and the
        achieved fraction of the roofline (the slower of NVLink bytes at 770 GB/s measured peer-copy bandwidth and
        HBM bytes at the measured copy bandwidth)."""
        import json
        import os as _os
        peaks = {"hbm_gbs": 6462.4}
        try:
            peaks.update(json.load(open(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))), "MEASURED_PEAKS.json"))))
        except Exception:  # noqa: BLE001
            pass
        self._timing = []
        for _ in range(2):
            eager_step(x, y)
        torch.cuda.synchronize()
        self._timing = []
        for _ in range(reps):
            eager_step(x, y)
        torch.cuda.synchronize()
        per = {}
        for b, e0, e1 in self._timing:
            per.setdefault((b["lo"], b["hi"]), []).append(e0.elapsed_time(e1) * 1e3)
        self._timing = None
        W = self.world
        rows = []
        for b in (self.buckets or [dict(lo=0, hi=self.n, chunk=self.chunk, st_off=0)]):
            us = sorted(per.get((b["lo"], b["hi"]), [0.0]))
            med = us[len(us) // 2]
            t = torch.tensor([med], dtype=torch.float64, device=self.graph.device())
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            med = float(t[0])
            mine = max(0, min(b["hi"], b["lo"] + (self.rank + 1) * b["chunk"]) - (b["lo"] + self.rank * b["chunk"]))
            nv_in = mine * 4 if self.mc else (W - 1) * mine * 4            # gradients pulled through the switch / from W-1 peers
            nv_out = mine * 6 if self.mc else (W - 1) * mine * 6           # fp32 + bf16 parameters pushed to every replica
            hbm = mine * (4 + 8 + 16 + 2 + 4) + (b["hi"] - b["lo"]) * 4     # own grad read, p rw, m/v rw, bf16 w, (local copy) + grad zeroing
            t_nv = max(nv_in, nv_out) / 770e9
            t_hbm = hbm / (peaks["hbm_gbs"] * 1e9)
            roof = max(t_nv, t_hbm) * 1e6
            rows.append({"elems": b["hi"] - b["lo"], "us": med, "bytes_nvlink_in": nv_in, "bytes_nvlink_out": nv_out, "bytes_hbm": hbm,
                         "roofline_us": roof, "frac_of_roofline": (roof / med) if med > 0 else None,
                         "frac_of_nvlink_roofline": (t_nv * 1e6 / med) if med > 0 else None,
                         "frac_of_hbm_roofline": (t_hbm * 1e6 / med) if med > 0 else None})
        return {"buckets": len(rows), "path": "nvls-multimem" if self.mc else "p2p", "ctas": self.ctas, "per_bucket": rows,
                "us_total": sum(r["us"] for r in rows), "bytes_nvlink": sum(r["bytes_nvlink_in"] + r["bytes_nvlink_out"] for r in rows),
                "note": "eager pass, CUDA events on the side stream, median of %d, max over ranks; a bucket's time includes the wait for "
                        "the slowest rank at its barrier" % reps}

    def backward_hook(self, layer_index: int):
        """Called by the top-level Sequential right after layer `layer_index` finished its backward:
        when that completes a bucket, its fused reduce+update is queued on the side stream."""
        if self.side is None:
            return
        for b in self.buckets:
            if b["first_layer"] == layer_index:
                main = torch.cuda.current_stream(self.graph.device())
                self.side.wait_stream(main)
                if side_stream.has_pending(self.graph.device()):
                    self.side.wait_stream(side_stream.stream(self.graph.device()))   # the bucket's weight gradients
                with torch.cuda.stream(self.side):
                    if getattr(self, "_timing", None) is not None:
                        self._timed_launch(b)
                    else:
                        self._launch(b["lo"], b["hi"], b["chunk"], b["st_off"])
                self._pending = True

    def step_overlapped_finish(self):
        """End of the step: join the side stream (all buckets reduced and updated everywhere)."""
        # (the step state was advanced by TrainStep BEFORE forward: the bucket kernels already used it)
        self.optimizer.step_count += 1
        torch.cuda.current_stream(self.graph.device()).wait_stream(self.side)
        self._pending = False
        for layer in self.graph.ctx.layers:
            layer.mark_lp_fresh()

    # ---------------------------------------------------------------- one fused step
    def step(self):
        if self.side is not None and getattr(self, "_pending", False):
            return self.step_overlapped_finish()
        opt = self.optimizer
        dev = self.graph.device()
        if not opt.external_state:
            stepstate.get(dev).advance(opt)
        opt.step_count += 1
        if self.buckets:
            for b in self.buckets:          # same bucket layout (and optimizer-state layout) without overlap
                self._launch(b["lo"], b["hi"], b["chunk"], b["st_off"])
        else:
            self._launch()
        for layer in self.graph.ctx.layers:
            layer.mark_lp_fresh()

    # ---------------------------------------------------------------- checkpoint / resume (sharded state)
    def _shards(self):
        """(lo, hi, chunk, st_off) of every launch's slab range: rank r owns [lo + r*chunk, min(hi, lo + (r+1)*chunk))
        and keeps its optimizer state at [st_off, st_off + chunk) of self.m / self.v."""
        if self.buckets:
            return [(b["lo"], b["hi"], b["chunk"], b["st_off"]) for b in self.buckets]
        return [(0, self.n, self.chunk, 0)]

    def _gather(self, shard_state: torch.Tensor) -> torch.Tensor:
        full = torch.zeros(self.n, dtype=torch.float32, device=shard_state.device)
        for lo, hi, c, st in self._shards():
            mine = torch.zeros(c, dtype=torch.float32, device=shard_state.device)
            k = min(c, shard_state.numel() - st)
            mine[:k] = shard_state[st:st + k]
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            for r, part in enumerate(parts):
                a, b = lo + r * c, min(hi, lo + (r + 1) * c)
                if b > a:
                    full[a:b] = part[:b - a]
        return full

    def _scatter(self, full: torch.Tensor, shard_state: torch.Tensor):
        full = full.to(shard_state.device, torch.float32).reshape(-1)
        for lo, hi, c, st in self._shards():
            a, b = lo + self.rank * c, min(hi, lo + (self.rank + 1) * c)
            if b > a:
                shard_state[st:st + (b - a)] = full[a:b]

    def state_dict(self) -> dict:
        """Optimizer state with the W-way sharded moments gathered to full slab length (collective: every rank calls)."""
        d = dict(self.optimizer.state_dict())
        if isinstance(self.optimizer, Adam):
            d["m"], d["v"] = self._gather(self.m), self._gather(self.v)
        elif self.m is not None and getattr(self.optimizer, "momentum", 0.0) > 0:
            d["velocity"] = self._gather(self.m)
        return d

    def load_state_dict(self, d: dict):
        Optimizer.load_state_dict(self.optimizer, {k: v for k, v in d.items() if k not in ("m", "v", "velocity")})
        if isinstance(self.optimizer, Adam):
            if d.get("m") is not None:
                self._scatter(d["m"], self.m)
            if d.get("v") is not None:
                self._scatter(d["v"], self.v)
        elif d.get("velocity") is not None:
            self._scatter(d["velocity"], self.m)

    def broadcast_parameters(self, src: int = 0):
        """Make every replica bit-identical to rank `src` (needed when replicas were built without a common seed
        or after rank 0 loaded a checkpoint)."""
        dist.broadcast(self.graph.ctx.param_slab, src, group=self.group)
        dist.broadcast(self.graph.ctx.buffer_slab, src, group=self.group)
        self.graph.ctx.refresh_lp()

    def roofline_bytes(self) -> dict:
        """Bytes this rank moves over NVLink per step (for the achieved-fraction report)."""
        W = self.world
        read = (W - 1) * self.chunk * 4 if not self.mc else self.chunk * 4
        write = (W - 1) * self.chunk * 6 if not self.mc else self.chunk * 6
        return {"nvlink_read": read, "nvlink_write": write, "hbm": self.chunk * (4 * 4 + 4 + 2) + self.n * 4}
