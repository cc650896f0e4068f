#!/usr/bin/env python
"""MSequential: parallel Sequential branches joined by an n-ary layer (reference:
examples/msequential_example.cpp:105-272 -- a two-branch basic case and three heterogeneous branches
of different depth, forward + backward, with the memory-aware branch order printed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.nn import AddLayer, LayerBuilder, MSequential, MulLayer, Sequential  # noqa: E402


def run(mseq, x, title):
    print(f"=== {title} ===")
    mseq.set_training(True)
    print("Forward Pass")
    y = mseq.forward(x, 0)
    print(f"Output shape: {tuple(y.shape)}")
    print("First 5 output values:", " ".join(f"{v:.4f}" for v in y.flatten()[:5].float().tolist()))
    print(f"Branch order chosen by the block (largest transient memory first): {mseq._order}")
    print("Backward Pass")
    dx = mseq.backward(torch.ones_like(y), 0)
    print(f"Input gradient shape: {tuple(dx.shape)}; parameter tensors with gradients: "
          f"{sum(len(l.grads) for s in mseq.sequences for l in s.layers)}")
    print("Backward pass completed successfully.\n")
    return y, dx


def basic(device):
    b1 = Sequential(LayerBuilder([[10]]).dense(20, True, "b1_fc1").activation("relu", "b1_relu").dense(5, True, "b1_fc2").build(), "branch_1")
    b2 = Sequential(LayerBuilder([[10]]).dense(15, True, "b2_fc1").activation("tanh", "b2_tanh").dense(5, True, "b2_fc2").build(), "branch_2")
    ms = MSequential([b1, b2], AddLayer("join_add"), "mseq_basic").init(device, seed=1)
    return run(ms, torch.randn(4, 10, device=device), "MSequential Basic Example")


def heterogeneous(device):
    shallow = Sequential(LayerBuilder([[100]]).dense(50, True, "branch1_dense").build(), "shallow_branch")
    medium = Sequential(LayerBuilder([[100]]).dense(200, True, "branch2_dense1").activation("relu", "branch2_relu")
                        .dense(50, True, "branch2_dense2").build(), "medium_branch")
    deep = Sequential(LayerBuilder([[100]]).dense(400, True, "branch3_dense1").activation("relu", "branch3_relu1")
                      .dense(200, True, "branch3_dense2").activation("relu", "branch3_relu2").dense(50, True, "branch3_dense3").build(),
                      "deep_branch")
    ms = MSequential([shallow, medium, deep], MulLayer("join_mul"), "mseq_hetero").init(device, seed=2)
    return run(ms, torch.randn(8, 100, device=device), "MSequential Heterogeneous Branches Example")


def main():
    device = "cuda" if torch.cuda.is_available() else "cpu"
    basic(device)
    heterogeneous(device)
    print("=== Example Complete ===")
    return 0


if __name__ == "__main__":
    sys.exit(main())
