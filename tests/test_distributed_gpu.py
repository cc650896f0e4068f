"""Device rendezvous of the Coordinator / Worker transport: GPU tensors travel device-to-device into the receiver's
CUDA-IPC slab (``tnn_b200/distributed/rendezvous.py``) instead of being staged through host memory.

Reference: the RoCE backend RDMA-writes into a slab the receiver registered with the NIC
(include/distributed/roce_channel.hpp:17-418, include/device/ibv_allocator.hpp:31-266); round 1 of this repo staged
every job through ``tensor.cpu()`` like the reference's TCP ``Writer`` (include/distributed/io.hpp:46-55).
"""
import multiprocessing as mp
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def _child_device() -> int:
    """Device of the second process.  Same GPU by default (CUDA-IPC between two processes on one device); TNN_TEST_CROSS_GPU=1
    puts it on cuda:1 so the one-sided write crosses NVLink.  (The Python -> C++ test below always uses the second GPU when
    there is one; it passed on a 2-GPU box.  The two all-Python cross-GPU variants failed there in the last run of the round
    because the child never made its device current -- fixed in Worker.run / Layer.forward, not re-run for lack of GPU budget.)"""
    return 1 if (torch.cuda.device_count() > 1 and os.environ.get("TNN_TEST_CROSS_GPU", "0") == "1") else 0


def test_device_rendezvous_same_process_zero_copy_blocks_and_fallbacks():
    _need_gpu()
    from tnn_b200.distributed import CommandType, Job, Message, RoCECommunicator
    dev = torch.device("cuda", 0)
    a = RoCECommunicator(name="a", device=dev, device_slab_bytes=8 << 20).start()
    b = RoCECommunicator(name="b", device=dev, device_slab_bytes=8 << 20).start()
    try:
        a.connect(b.endpoint())
        g = torch.Generator(device=dev).manual_seed(1)
        tensors = [torch.randn(256, 1024, generator=g, device=dev).bfloat16(),                    # 512 KiB bf16
                   torch.randn(3, 64, 1000, generator=g, device=dev),                             # 750 KiB fp32, 3-D
                   torch.arange(1 << 18, device=dev, dtype=torch.int32),                          # 1 MiB int32
                   torch.randn(16, generator=g, device=dev),                                      # tiny: inline host path
                   torch.randn(3 << 20, generator=g, device=dev)]                                 # 12 MiB > slab: host fallback
        for i, t in enumerate(tensors):
            a.send_message(Message(CommandType.FORWARD_JOB, Job(t, i)), b.endpoint())
        got = {}
        for _ in tensors:
            m = b.dequeue_input_message(timeout=60)
            assert m is not None and m.command_type == CommandType.FORWARD_JOB
            got[m.payload.mb_id] = m.payload.data
        for i, t in enumerate(tensors):
            assert got[i].is_cuda and got[i].dtype == t.dtype and got[i].shape == t.shape and torch.equal(got[i], t), i
        sa, sb = a.stats(), b.stats()
        assert sa["device_rendezvous_sent"] == 3 and sb["device_rendezvous_received"] == 3 and sa["device_fallback_host"] == 1
        # the delivered tensor IS the slab block (no receiver-side copy)
        lo, hi = b._dptr, b._dptr + b.device_slab_bytes
        assert all(lo <= got[i].data_ptr() < hi for i in range(3)) and not (lo <= got[3].data_ptr() < hi)
        # blocks return to the slab when the tensors die: 40 x 1 MiB through an 8 MiB slab
        got.clear()
        del m
        big = torch.randn(1 << 18, generator=g, device=dev)
        for i in range(40):
            a.send_message(Message(CommandType.FORWARD_JOB, Job(big, 100 + i)), b.endpoint())
            r = b.dequeue_input_message(timeout=60)
            assert r.payload.mb_id == 100 + i and torch.equal(r.payload.data, big)
            del r
        sb = b.stats()
        assert sb["device_rendezvous_received"] >= 3 + 36, sb       # a few may fall back while frees wait for their event
        assert sb["device_slab_peak"] <= 8 << 20
        # the other direction uses the same pair of connections
        b.send_message(Message(CommandType.BACKWARD_JOB, Job(tensors[0], 9)), a.endpoint())
        back = a.dequeue_input_message(timeout=60)
        assert back.payload.mb_id == 9 and torch.equal(back.payload.data, tensors[0]) and a.stats()["device_rendezvous_received"] == 1
    finally:
        a.stop()
        b.stop()


def _echo_main(port, ready, dev_index):
    """Child process: receive FORWARD_JOBs on the GPU, double them with a kernel of this repo, send them back."""
    import torch as T
    from tnn_b200.distributed import CommandType, Endpoint, Job, Message, RoCECommunicator
    from tnn_b200.ops import tensor_ops
    dev = T.device("cuda", dev_index)
    T.cuda.set_device(dev)               # the kernels launch on the current device's stream
    c = RoCECommunicator("127.0.0.1", port, name="echo", device=dev, device_slab_bytes=32 << 20).start()
    ready.set()
    try:
        while True:
            m = c.dequeue_input_message(timeout=120)
            if m is None or m.command_type == CommandType.SHUTDOWN:
                break
            x = m.payload.data
            assert x.is_cuda
            y = tensor_ops.mul(x.contiguous(), 2.0)
            host, p = m.sender.rsplit(":", 1)
            c.send_message(Message(CommandType.BACKWARD_JOB, Job(y, m.payload.mb_id)), Endpoint.roce(host, int(p)))
        st = c.stats()
        assert st["device_rendezvous_received"] > 0 and st["device_rendezvous_sent"] > 0, st
    finally:
        c.stop()


def test_device_rendezvous_between_two_processes():
    """Two processes, CUDA-IPC: the parent's tensor lands in the child's device slab and the reply in the parent's."""
    _need_gpu()
    from tnn_b200.distributed import CommandType, Endpoint, Job, Message, RoCECommunicator
    ctx = mp.get_context("spawn")
    port = 23000 + (os.getpid() % 20000)
    ready = ctx.Event()
    child_dev = _child_device()
    proc = ctx.Process(target=_echo_main, args=(port, ready, child_dev), daemon=True)
    proc.start()
    assert ready.wait(180)
    dev = torch.device("cuda", 0)
    me = RoCECommunicator("127.0.0.1", 0, name="parent", device=dev, device_slab_bytes=32 << 20).start()
    try:
        peer = Endpoint.roce("127.0.0.1", port)
        me.connect(peer)
        g = torch.Generator(device=dev).manual_seed(2)
        xs = [torch.randn(64, 32, 32, 16, generator=g, device=dev).bfloat16(), torch.randn(1 << 20, generator=g, device=dev)]
        for rep in range(3):
            for i, x in enumerate(xs):
                me.send_message(Message(CommandType.FORWARD_JOB, Job(x, 10 * rep + i)), peer)
            for i, x in enumerate(xs):
                m = me.dequeue_input_message(timeout=120)
                assert m is not None and m.command_type == CommandType.BACKWARD_JOB
                want = xs[m.payload.mb_id % 10]
                assert m.payload.data.is_cuda and torch.equal(m.payload.data.float(), want.float() * 2)
        st = me.stats()
        assert st["device_rendezvous_sent"] == 6 and st["device_rendezvous_received"] == 6 and st["device_fallback_host"] == 0, st
        me.send_message(Message(CommandType.SHUTDOWN), peer)
        proc.join(60)
        assert proc.exitcode == 0
    finally:
        me.stop()
        if proc.is_alive():
            proc.terminate()


def _roce_gpu_worker_main(port, ready, dev_index):
    from tnn_b200.distributed import RoCEWorker
    w = RoCEWorker("127.0.0.1", port, use_gpu=True, gpu_index=dev_index, name=f"g{port}", slab_bytes=16 << 20)
    w.comm.start()
    ready.set()
    w.run()
    w.comm.stop()


def test_roce_pipeline_on_gpus_moves_activations_device_to_device():
    """roce_coordinator + a roce_worker process, both stages on GPUs: activations and their gradients cross through the
    device slabs; the loss trajectory equals the one-device run."""
    _need_gpu()
    from tnn_b200.distributed import CoordinatorConfig, Endpoint, ParallelMode, RoCECoordinator, RoCEWorker, split_microbatches
    from tnn_b200.nn import LayerBuilder, LossFactory, OptimizerFactory, SchedulerFactory, Sequential
    from tnn_b200.nn.graph import graph_from_sequential
    from tnn_b200.parallel import NaivePipelinePartitioner

    def model():
        b = LayerBuilder([16, 16, 16])
        b.conv2d(32, 3, 3, 1, 1, 1, 1, True, "c1").activation("relu", "a1")
        b.conv2d(32, 3, 3, 1, 1, 1, 1, True, "c2").activation("relu", "a2").maxpool2d(2, 2, name="p")
        b.flatten(name="f").dense(64, True, "d1").activation("tanh", "a3").dense(10, True, "d2")
        return Sequential(b.build(), "pp_gpu")

    seed, M, B = 5, 2, 64
    gen = torch.Generator().manual_seed(3)
    batches = [(torch.randn(B, 16, 16, 16, generator=gen), torch.randint(0, 10, (B,), generator=gen, dtype=torch.int32)) for _ in range(3)]
    crit = LossFactory.create_logsoftmax_crossentropy()
    # one-device reference on the GPU (same bf16 kernels)
    gr = graph_from_sequential(model(), "cuda:0", seed)
    opt = OptimizerFactory.create_sgd(0.05, 0.9)
    opt.attach(gr.context())
    net = gr.get_layers()[0]
    ref = []
    for x, y in batches:
        xs, ys = split_microbatches(x.cuda(), y.cuda(), M)
        tot = 0.0
        for i, (xm, ym) in enumerate(zip(xs, ys)):
            out = net.forward(xm, i)
            stats, grad = crit.loss_and_grad(out, ym, scale=1.0 / M)
            net.backward(grad, i)
            tot += float(stats[0]) / M
        opt.update()
        opt.zero_grads()
        ref.append(tot)

    ctx = mp.get_context("spawn")
    port = 25000 + (os.getpid() % 20000)
    ready = ctx.Event()
    child_dev = _child_device()
    proc = ctx.Process(target=_roce_gpu_worker_main, args=(port, ready, child_dev), daemon=True)
    proc.start()
    assert ready.wait(180)
    local = RoCEWorker("127.0.0.1", 0, use_gpu=True, gpu_index=0, name="local", slab_bytes=16 << 20)
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, model(), OptimizerFactory.create_sgd(0.05, 0.9),
                            SchedulerFactory.create_step_lr(OptimizerFactory.create_sgd(0.05, 0.9), 100, 0.5),
                            NaivePipelinePartitioner([1, 1]), local_worker=local, local_worker_position=0,
                            worker_endpoints=[Endpoint.roce("127.0.0.1", port)], seed=seed)
    co = RoCECoordinator(cfg, "127.0.0.1", 0, slab_bytes=16 << 20)
    try:
        co.start()
        losses = []
        for x, y in batches:
            xs, ys = split_microbatches(x, y, M)
            l, _ = co.async_train_batch(xs, ys, crit)
            co.update_parameters()
            losses.append(l)
        for r, l in zip(ref, losses):
            assert abs(r - l) < 5e-2 * max(1.0, abs(r)), (ref, losses)
        st = local.comm.stats()
        assert st["device_rendezvous_sent"] > 0 and st["device_rendezvous_received"] > 0, st      # activations out, gradients back
    finally:
        co.stop()
        local.stop()
        proc.join(20)
        if proc.is_alive():
            proc.terminate()


def test_python_gpu_stage_and_cpp_roce_worker_exchange_device_tensors():
    """Mixed hosts on one node: stage 0 is a Python ``RoCEWorker`` on the GPU, stage 1 the C++ ``roce_worker --gpu``
    (cpp/src/rendezvous.cpp).  Activations go Python -> C++ and gradients C++ -> Python through the device rendezvous of
    both implementations (same frames, same metadata record, CUDA-IPC handles exchanged between the processes)."""
    _need_gpu()
    import socket
    import subprocess
    import time

    from tnn_b200.distributed import CoordinatorConfig, Endpoint, ParallelMode, RoCECoordinator, RoCEWorker, split_microbatches
    from tnn_b200.models import ExampleModels
    from tnn_b200.nn import LossFactory, OptimizerFactory
    from tnn_b200.parallel import NaivePipelinePartitioner
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "cpp", "bin", "roce_worker")
    if not os.access(exe, os.X_OK):
        pytest.skip("C++ host not built")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    child_dev = 1 if torch.cuda.device_count() > 1 else 0
    proc = subprocess.Popen([exe, str(port), "--gpu", str(child_dev)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    local = RoCEWorker("127.0.0.1", 0, use_gpu=True, gpu_index=0, name="py-gpu", slab_bytes=8 << 20)
    try:
        time.sleep(2.0)
        cfgc = CoordinatorConfig(ParallelMode.PIPELINE, ExampleModels.create("mnist_cnn"), OptimizerFactory.create_sgd(0.05, 0.9), None,
                                 NaivePipelinePartitioner([1, 1]), local_worker=local, local_worker_position=0,
                                 worker_endpoints=[Endpoint.roce("127.0.0.1", port)], seed=5)
        co = RoCECoordinator(cfgc, "127.0.0.1", 0, slab_bytes=8 << 20)
        co.start()
        crit = LossFactory.create_logsoftmax_crossentropy()
        g = torch.Generator().manual_seed(0)
        x = torch.rand(64, 28, 28, 1, generator=g)
        y = (torch.arange(64) % 10).to(torch.int32)
        losses = []
        for _ in range(6):
            xs, ys = split_microbatches(x, y, 2)
            l, _ = co.async_train_batch(xs, ys, crit)
            co.update_parameters()
            losses.append(l)
        assert losses[-1] < losses[0], losses
        st = local.comm.stats()
        assert st["device_rendezvous_sent"] >= 12 and st["device_rendezvous_received"] >= 12 and st["device_fallback_host"] == 0, st
        co.stop()
    finally:
        local.stop()
        try:
            proc.wait(15)
        except Exception:  # noqa: BLE001
            proc.kill()
