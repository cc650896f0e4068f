#!/usr/bin/env python
"""torch re-creation of the reference's pipeline-parallel configs as a MEASURED library baseline: S stages, one
process per GPU, activations and activation gradients over ``torch.distributed`` p2p (NCCL on GPUs, gloo on CPU),
non-interleaved 1F1B, torch autograd + cuDNN inside each stage, Adam per stage.  Mirrors
/root/reference/torch/torch_resnet9_pipeline*.py (a TensorPipe-RPC 2-stage ResNet-9) and BASELINE.md's
"WRN-16-8 4-stage PP: torch pipeline with NCCL p2p" row; ours to beat is ``tnn_b200.parallel.pipeline``.

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 baseline/torch_pipeline.py \
      --model cifar100_wrn16_8 --batch 256 --microbatches 4 --steps 20 --warmup 5 [--proportions 1,2]

Prints one JSON line on rank 0 (img/s, device-timed, max over ranks).  ``--model tiny`` is a small CNN used by the
CPU test, which checks that S = 2 reproduces the S = 1 loss trajectory.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from torch_wrn import WRN16_8, ResNet9  # noqa: E402


def model_layers(name: str):
    """The model as an ordered list of modules that can be cut between any two of them."""
    if name == "tiny":
        torch.manual_seed(0)
        return [nn.Sequential(nn.Conv2d(3, 8, 3, 1, 1), nn.ReLU()), nn.Sequential(nn.Conv2d(8, 8, 3, 1, 1), nn.ReLU()),
                nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten()), nn.Linear(8, 10)], 10
    torch.manual_seed(0)
    if "resnet9" in name:
        return list(ResNet9(10).net.children()), 10
    m = WRN16_8(100)
    tail = nn.Sequential(m.bn, nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten())
    return [m.conv1, *m.blocks.children(), tail, m.fc], 100


def split(n_layers: int, proportions):
    psum, cur, out = sum(proportions), 0, []
    for i, p in enumerate(proportions):
        end = n_layers if i == len(proportions) - 1 else min(max(cur + n_layers * p // psum, cur + 1), n_layers)
        out.append((cur, end))
        cur = end
    return out


class Stage:
    def __init__(self, a):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))) if torch.cuda.is_available() else torch.device("cpu")
        layers, self.classes = model_layers(a.model)
        props = [int(p) for p in a.proportions.split(",")] if a.proportions else [1] * self.world
        assert len(props) == self.world, "one proportion per stage"
        lo, hi = split(len(layers), props)[self.rank]
        self.net = nn.Sequential(*layers[lo:hi]).to(self.dev)
        if self.dev.type == "cuda":
            self.net = self.net.to(memory_format=torch.channels_last)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=1e-3, eps=1e-3, weight_decay=3e-4,
                                    **({"fused": True} if self.dev.type == "cuda" else {}))
        self.M, self.mb = a.microbatches, a.batch // a.microbatches
        self.first, self.last = self.rank == 0, self.rank == self.world - 1
        self.act_dtype = torch.bfloat16 if self.dev.type == "cuda" else torch.float32
        self.hw = 32
        # shape handshake: each stage tells the next one what it will send
        self.in_shape = (self.mb, 3, self.hw, self.hw)
        if not self.first:
            t = torch.zeros(8, dtype=torch.int64, device=self.dev)
            dist.recv(t, self.rank - 1)
            self.in_shape = tuple(int(v) for v in t.tolist() if v > 0)
        with torch.no_grad(), self._autocast():
            was = self.net.training
            self.net.eval()
            out = self.net(torch.zeros(self.in_shape, device=self.dev))
            self.net.train(was)
        self.out_shape = tuple(out.shape)
        if not self.last:
            t = torch.zeros(8, dtype=torch.int64, device=self.dev)
            t[:len(self.out_shape)] = torch.tensor(self.out_shape)
            dist.send(t, self.rank + 1)

    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16) if self.dev.type == "cuda" else torch.autocast("cpu", enabled=False)

    def _ops(self, send_next=None, recv_next=False, send_prev=None, recv_prev=False):
        """One grouped exchange with the neighbours (grouping send+recv is what keeps NCCL p2p from deadlocking when
        both sides send first).  Returns (activation from prev, gradient from next)."""
        ops, act, grad = [], None, None
        if send_prev is not None:
            ops.append(dist.P2POp(dist.isend, send_prev, self.rank - 1))
        if recv_prev:
            act = torch.empty(self.in_shape, dtype=self.act_dtype, device=self.dev)
            ops.append(dist.P2POp(dist.irecv, act, self.rank - 1))
        if send_next is not None:
            ops.append(dist.P2POp(dist.isend, send_next, self.rank + 1))
        if recv_next:
            grad = torch.empty(self.out_shape, dtype=self.act_dtype, device=self.dev)
            ops.append(dist.P2POp(dist.irecv, grad, self.rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return act, grad

    def _fwd(self, inp, y, saved):
        if not self.first:
            inp.requires_grad_(True)
        with self._autocast():
            out = self.net(inp)
            if self.last:
                out = F.cross_entropy(out.float(), y) / self.M
        saved.append((inp, out))
        return out if self.last else out.detach().to(self.act_dtype).contiguous()

    def _bwd(self, g, saved):
        inp, out = saved.pop(0)
        if self.last:
            out.backward()
        else:
            out.backward(g.to(out.dtype))
        return None if self.first else inp.grad.to(self.act_dtype).contiguous()

    def train_batch(self, x, y):
        """1F1B: min(S - rank - 1, M) warm-up forwards, steady one-forward-one-backward, cool-down backwards."""
        xs = x.chunk(self.M) if self.first else [None] * self.M
        ys = y.chunk(self.M) if self.last else [None] * self.M
        warm = min(self.world - self.rank - 1, self.M)
        remaining = self.M - warm
        saved, f, loss = [], 0, torch.zeros((), device=self.dev)
        for _ in range(warm):
            inp, _ = self._ops(recv_prev=not self.first)
            out = self._fwd(xs[f] if self.first else inp, ys[f], saved)
            f += 1
            self._ops(send_next=out)
        inp = None
        if remaining > 0 and not self.first:
            inp, _ = self._ops(recv_prev=True)
        for k in range(remaining):
            out = self._fwd(xs[f] if self.first else inp, ys[f], saved)
            f += 1
            if self.last:
                loss += out.detach()
                gout = None
            else:
                _, gout = self._ops(send_next=out, recv_next=True)
            gin = self._bwd(gout, saved)
            if self.first:
                inp = None
            elif k == remaining - 1:
                self._ops(send_prev=gin)
            else:
                inp, _ = self._ops(send_prev=gin, recv_prev=True)
        for _ in range(warm):
            _, gout = self._ops(recv_next=True)
            gin = self._bwd(gout, saved)
            if not self.first:
                self._ops(send_prev=gin)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="cifar100_wrn16_8")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--microbatches", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--proportions", default="")
    ap.add_argument("--print-losses", action="store_true")
    a = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        torch.backends.cudnn.benchmark = True
    dist.init_process_group("nccl" if cuda else "gloo")
    st = Stage(a)
    g = torch.Generator().manual_seed(1)
    xs = [torch.rand(a.batch, 3, 32, 32, generator=g) for _ in range(4)]
    ys = [torch.randint(0, st.classes, (a.batch,), generator=g) for _ in range(4)]
    xs = [t.to(st.dev).to(memory_format=torch.channels_last) if cuda else t for t in xs]
    ys = [t.to(st.dev) for t in ys]
    losses = []
    for i in range(a.warmup):
        losses.append(st.train_batch(xs[i % 4], ys[i % 4]))
    dist.barrier()
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    else:
        import time
        t0 = time.perf_counter()
    for i in range(a.steps):
        losses.append(st.train_batch(xs[(a.warmup + i) % 4], ys[(a.warmup + i) % 4]))
    if cuda:
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    else:
        ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=st.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ls = torch.stack(losses).float()
    dist.broadcast(ls, st.world - 1)
    if st.rank == 0:
        res = {"impl": "torch-pipeline", "metric": f"{a.model} training throughput", "value": a.batch * a.steps / (float(t[0]) / 1e3),
               "unit": "img/s", "n_gpus": st.world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(t[0]) / max(1, a.steps),
               "higher_is_better": True, "dtype": "bf16" if cuda else "fp32", "data": "synthetic",
               "config": {"model": a.model, "global_batch": a.batch, "microbatches": a.microbatches,
                          "parallelism": f"pp{st.world}", "schedule": "1F1B", "stack": "torch autograd + cuDNN + p2p"}}
        if a.print_losses:
            res["losses"] = [round(float(v), 6) for v in ls.tolist()]
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
