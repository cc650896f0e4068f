"""torch + cuDNN + NCCL re-creation of the reference WRN-16-8 / CIFAR-100 training config
(mirrors /root/reference/torch/torch_wrn16_8_cifar100.py semantics: pre-activation wide blocks,
dropout 0.3, Adam).  Used only as the MEASURED library baseline (bench.py --impl torch)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class WideBlock(nn.Module):
    def __init__(self, cin, cout, stride, p):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(cin, eps=1e-2)
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=True)
        self.bn2 = nn.BatchNorm2d(cout, eps=1e-2)
        self.drop = nn.Dropout(p)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=True)
        self.short = nn.Conv2d(cin, cout, 1, stride, 0, bias=False) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        o = self.conv1(F.relu(self.bn1(x)))
        o = self.conv2(self.drop(F.relu(self.bn2(o))))
        return o + (self.short(x) if self.short is not None else x)


class WRN16_8(nn.Module):
    def __init__(self, classes=100, p=0.3):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, 1, 1)
        cfg = [(16, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2), (512, 512, 1)]
        self.blocks = nn.Sequential(*[WideBlock(a, b, s, p) for a, b, s in cfg])
        self.bn = nn.BatchNorm2d(512, eps=1e-2)
        self.fc = nn.Linear(512, classes)

    def forward(self, x):
        x = self.blocks(self.conv1(x))
        x = F.relu(self.bn(x))
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


class ResNet9(nn.Module):
    """torch mirror of the zoo's cifar10_resnet9 (reference: torch/torch_resnet9_deepspeed.py model)."""

    def __init__(self, classes=10):
        super().__init__()

        def cb(i, o):
            return nn.Sequential(nn.Conv2d(i, o, 3, 1, 1, bias=False), nn.BatchNorm2d(o, eps=1e-2), nn.ReLU(inplace=True))

        class Res(nn.Module):
            def __init__(self, c):
                super().__init__()
                self.c1 = nn.Conv2d(c, c, 3, 1, 1, bias=False)
                self.b1 = nn.BatchNorm2d(c, eps=1e-2)
                self.c2 = nn.Conv2d(c, c, 3, 1, 1, bias=False)
                self.b2 = nn.BatchNorm2d(c, eps=1e-2)

            def forward(self, x):
                return F.relu(x + self.b2(self.c2(F.relu(self.b1(self.c1(x))))))

        self.net = nn.Sequential(cb(3, 64), cb(64, 128), nn.MaxPool2d(2), Res(128), Res(128), cb(128, 256), nn.MaxPool2d(2),
                                 Res(256), Res(256), cb(256, 512), nn.MaxPool2d(2), Res(512), nn.AvgPool2d(4), nn.Flatten(),
                                 nn.Linear(512, classes))

    def forward(self, x):
        return self.net(x)


def bench_torch_wrn(B, steps, warmup, local, world, model_name="cifar100_wrn16_8", graphed=True):
    """Library arm.  ``graphed`` captures forward + backward + fused Adam in ONE CUDA graph (static inputs,
    capturable optimizer) so this arm is not launch-bound -- the same execution model as tnn_b200's TrainStep."""
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = True
    classes = 10 if "cifar10_" in model_name else 100
    model = (ResNet9(classes) if "resnet9" in model_name else WRN16_8(classes)).to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, eps=1e-3, weight_decay=3e-4, fused=True, capturable=graphed)
    xs = [torch.rand(B, 3, 32, 32, device=dev).to(memory_format=torch.channels_last) for _ in range(4)]
    ys = [torch.randint(0, classes, (B,), device=dev) for _ in range(4)]
    sx, sy = xs[0].clone(memory_format=torch.channels_last), ys[0].clone()

    def eager(x, y):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.cross_entropy(model(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    graph, gloss, mode = None, None, "eager"
    if graphed:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(11 if world > 1 else 3):     # DDP needs 11 warm-up iterations before capture
                    eager(sx, sy)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
                    gloss = F.cross_entropy(model(sx), sy)
                gloss.backward()
                opt.step()
            mode = "cuda-graph"
        except Exception as e:  # noqa: BLE001
            graph, mode = None, f"eager (graph capture failed: {type(e).__name__})"
            torch.cuda.synchronize()

    def step(i):
        if graph is None:
            return eager(xs[i % 4], ys[i % 4])
        sx.copy_(xs[i % 4])
        sy.copy_(ys[i % 4])
        graph.replay()
        return gloss

    for i in range(max(warmup, 3)):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return {"metric": ("WRN-16-8 CIFAR-100" if "wrn" in model_name else model_name) + " training throughput", "value": world * B * steps / (ms / 1e3), "unit": "img/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"model": model_name, "global_batch": world * B, "parallelism": f"ddp{world}",
                       "stack": "torch autocast bf16 + cuDNN channels_last + fused Adam + NCCL", "execution": mode}, "final_loss": float(loss)}
