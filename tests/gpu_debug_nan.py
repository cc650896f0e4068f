import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tnn_b200.models import load_or_create_model
from tnn_b200.nn import LossFactory, OptimizerFactory
from tnn_b200.nn.train import TrainStep
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = load_or_create_model("cifar100_wrn16_8", device=dev, seed=1234)
opt = OptimizerFactory.create_adam(1e-3, 0.9, 0.999, 1e-3, 3e-4).attach(g.ctx)
opt.fused_zero_grad = False
crit = LossFactory.create_logsoftmax_crossentropy()
step = TrainStep(g, crit, opt)
gen = torch.Generator().manual_seed(100)
xs = [(torch.rand((B, 32, 32, 3), generator=gen).to(dev), torch.randint(0, 100, (B,), generator=gen, dtype=torch.int32).to(dev)) for _ in range(4)]
for i in range(16):
    st = step(*xs[i % 4])
    l = float(st[0])
    gs = g.ctx.grad_slab
    print(i, "loss", l, "grad finite", bool(torch.isfinite(gs).all()), "gnorm", float(gs.norm()), "pnorm", float(g.ctx.param_slab.norm()), flush=True)
    if not math.isfinite(l) or not torch.isfinite(gs).all():
        for layer in g.ctx.layers:
            for k, v in layer.grads.items():
                if not torch.isfinite(v).all():
                    print("  bad grad:", layer.name, k, tuple(v.shape))
            for k, v in layer.params.items():
                if not torch.isfinite(v).all():
                    print("  bad param:", layer.name, k)
        break
    gs.zero_()
