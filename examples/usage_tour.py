#!/usr/bin/env python
"""The snippets of docs/USAGE.md as one runnable program: tensors, LayerBuilder / Sequential / TrainStep, data
augmentation, an in-process two-stage pipeline (reference: docs/tensor_usage.md, docs/sequential_usage.en.md,
docs/data_augmentation.md, docs/pipeline_architecture.md).

  python examples/usage_tour.py [--device cpu|cuda]"""
import argparse
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.core import tensor as T  # noqa: E402
from tnn_b200.data import AugmentationBuilder, SyntheticDataLoader  # noqa: E402
from tnn_b200.data.augmentation import AugmentationStrategy  # noqa: E402
from tnn_b200.distributed import Coordinator, CoordinatorConfig, InProcessWorker, ParallelMode, split_microbatches  # noqa: E402
from tnn_b200.models import ExampleModels  # noqa: E402
from tnn_b200.nn import LayerBuilder, LossFactory, OptimizerFactory, Sequential, graph_from_sequential  # noqa: E402
from tnn_b200.nn.train import TrainStep  # noqa: E402
from tnn_b200.parallel import NaivePipelinePartitioner  # noqa: E402


class ColorInversion(AugmentationStrategy):
    """A custom augmentation: invert the colours of an image with probability p."""
    name = "color_inversion"

    def __init__(self, p=0.25):
        self.p = p

    def apply(self, x, gen):                       # x: [B, H, W, C] float, gen: the pipeline's torch.Generator
        hit = torch.rand(x.shape[0], generator=gen) < self.p
        x = x.clone()
        x[hit] = 1.0 - x[hit]
        return x


def tensors(device):
    # tensors ARE torch tensors; core/tensor.py adds the reference's Tensor vocabulary on top
    a = T.make_tensor([2, 3])                                         # fp32, host
    b = torch.tensor([[1., 2., 3.], [4., 5., 6.]])
    T.fill(a, 0.5)
    c = T.mul_scalar(a + b, 2.0) + b
    print("c =", T.head(c, 6), T.stats(c))
    h = T.dtype_cast(c, "BF16")
    d = T.to_device(c, device)
    print(f"bf16 copy: {T.nbytes(h)} bytes; on {d.device}")
    x = T.fill_random_uniform(T.make_tensor([4, 2, 2, 1]), 0.0, 1.0, seed=7)
    T.span(x, 0, 2).zero_()                                           # views share storage
    print("after zeroing the view, sum of the first half:", float(x[:2].sum()))
    buf = io.BytesIO()
    T.save(c, buf)                                                    # {u32 dtype, u64 ndims, u64 dims[], raw bytes}
    buf.seek(0)
    print("round trip max |diff| =", float((T.load(buf) - c).abs().max()))


def sequential(device):
    b = LayerBuilder([32, 32, 3])                                     # batchless NHWC input; the builder tracks the running shape
    b.conv2d(16, 3, 3, 1, 1, 1, 1, False, "conv1").batchnorm(1e-5, 0.1, True, True, "bn1").maxpool2d(2, 2, 2, 2, name="pool1")
    b.basic_residual_block(16, 32, 2, "res1")
    b.avgpool2d(8, 8, 1, 1, name="gap").flatten(name="flat").dense(10, True, "head")
    print("shape after the builder:", b.get_batchless_current_shape())
    g = graph_from_sequential(Sequential(b.build(), "tiny_resnet"), device, seed=1)
    opt = OptimizerFactory.create_sgd(0.05, 0.9)
    step = TrainStep(g, LossFactory.create_logsoftmax_crossentropy(), opt)     # GPU: eager warm-up, then CUDA-graph replays
    data = SyntheticDataLoader((32, 32, 3), 10, 128, seed=3).load_data()
    losses = []
    for _ in range(8):
        x, y = data.get_batch(16)
        losses.append(float(step(x, y)[0]))                           # device tensor {mean loss, #correct}
    print(f"tiny_resnet: loss {losses[0]:.4f} -> {losses[-1]:.4f} over 8 steps; {g.context().num_parameters()} parameters")
    print("graph config is", len(g.get_config().dump()), "bytes of JSON")


def augmentation():
    aug = (AugmentationBuilder(seed=42).horizontal_flip(0.5).random_crop(1.0, 4).brightness(0.3, 0.15)
           .custom_augmentation(ColorInversion(0.25)).normalize([0.5] * 3, [0.25] * 3).build())
    print("pipeline:", " ".join(aug.names()))
    batch = torch.rand(8, 32, 32, 3, generator=torch.Generator().manual_seed(1))
    print("augmented batch mean", float(aug.apply(batch).mean()))
    aug.remove_augmentation("color_inversion")
    aug.remove_augmentation(0)                                        # by index: horizontal_flip
    print("after removals:", aug.size(), "strategies")
    SyntheticDataLoader((32, 32, 3), 10, 64).set_augmentation(aug)   # loaders apply it to every batch they hand out


def pipeline(device):
    workers = [InProcessWorker(device, f"stage{i}") for i in range(2)]
    workers[1].start(background=True)
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, ExampleModels.create("mnist_cnn"), OptimizerFactory.create_sgd(0.05, 0.9), None,
                            NaivePipelinePartitioner([1, 1]), local_worker=workers[0], local_worker_position=0,
                            worker_endpoints=[workers[1].comm.endpoint()])
    co = Coordinator(cfg).start()                                     # partitions the model, ships each stage as JSON
    crit = LossFactory.create_logsoftmax_crossentropy()
    data = SyntheticDataLoader((28, 28, 1), 10, 128, seed=5).load_data()
    try:
        for i in range(3):
            x, y = data.get_batch(32)
            xs, ys = split_microbatches(x, y, 4)
            loss, _ = co.async_train_batch(xs, ys, crit)              # forwards injected, backward per micro-batch as outputs return
            co.update_parameters()
            print(f"pipeline batch {i}: loss {loss:.4f}")
        for l in co.report_load():
            print(f"stage {l['stage']}: {l['avg_forward_time']:.3f} ms fwd, {l['avg_backward_time']:.3f} ms bwd per job")
    finally:
        co.stop()
        for w in workers:
            w.stop()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    dev = torch.device(ap.parse_args().device)
    for title, fn, args in (("tensors", tensors, (dev,)), ("LayerBuilder / Sequential / TrainStep", sequential, (dev,)),
                            ("data augmentation", augmentation, ()), ("pipeline", pipeline, (dev,))):
        print("==", title)
        fn(*args)
    print("usage tour complete")


if __name__ == "__main__":
    main()
