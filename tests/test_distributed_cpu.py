"""Distributed runtime on CPU: wire format, partitioners, in-process and TCP (multi-process)
pipeline training equivalent to single-device training, failure detection, profiling fan-in.

Mirrors unit_tests/partitioner_test.cpp and archiver_test.cpp and adds what the reference has
no automated test for: pipeline-vs-single-device loss-trajectory equivalence,
two-process TCP runs on localhost, fault injection.
"""
import multiprocessing as mp
import os
import time

import pytest
import torch

from tnn_b200.core.config import TConfig
from tnn_b200.distributed import (CommandType, Coordinator, CoordinatorConfig, DistributedError, Endpoint,
                                  InProcessWorker, Job, Message, NetworkCoordinator, ParallelMode, StageConfig,
                                  TCPCommunicator, TCPWorker, split_microbatches)
from tnn_b200.nn import LayerBuilder, LossFactory, OptimizerFactory, SchedulerFactory, Sequential, graph_from_sequential
from tnn_b200.ops import native
from tnn_b200.parallel import (FTDPartitioner, NaiveDataPartitioner, NaivePartitionerConfig, NaivePipelinePartitioner,
                               Partitioner, SeqPartition, WeightedPipelinePartitioner)
from tnn_b200.utils.profiler import EventType, Profiler

pytestmark = pytest.mark.skipif(native.load() is None, reason="native extension not built")


def _model(name="pp_model"):
    b = LayerBuilder([8, 8, 4])
    b.conv2d(8, 3, 3, 1, 1, 1, 1, True, "c1").activation("relu", "a1")
    b.conv2d(8, 3, 3, 1, 1, 1, 1, True, "c2").activation("relu", "a2").maxpool2d(2, 2, name="p")
    b.flatten(name="f").dense(16, True, "d1").activation("tanh", "a3").dense(5, True, "d2")
    return Sequential(b.build(), name)


# ------------------------------------------------------------------------------- wire format
def test_message_round_trip_all_payloads():
    t = torch.randn(3, 4, 5)
    m = Message(CommandType.FORWARD_JOB, Job(t, 7))
    data = m.serialize()
    assert m.byte_size() == len(data)
    assert data[:2] == (1).to_bytes(2, "little")                 # CommandType first: the bus reads priority from it
    m2 = Message.deserialize(data)
    assert m2.command_type == CommandType.FORWARD_JOB and m2.payload.mb_id == 7 and torch.equal(m2.payload.data, t)
    for dt in (torch.bfloat16, torch.int32, torch.uint8, torch.float64):
        x = (torch.arange(24).reshape(2, 12) % 7).to(dt)
        y = Message.deserialize(Message(CommandType.BACKWARD_JOB, Job(x, 1)).serialize()).payload.data
        assert y.dtype == dt and torch.equal(y, x)
    assert Message.deserialize(Message(CommandType.ERROR_REPORT, "boöm").serialize()).payload == "boöm"
    assert Message.deserialize(Message(CommandType.CONFIG_RECEIVED, True).serialize()).payload is True
    assert Message.deserialize(Message(CommandType.SHUTDOWN).serialize()).payload is None
    p = Profiler()
    p.init_start_time()
    p.add_event(EventType.COMPUTE, 10, 20, "fwd", "w0")
    p2 = Message.deserialize(Message(CommandType.PROFILING_REPORTED, p).serialize()).payload
    assert [e["name"] for e in p2.events()] == ["fwd"]
    sc = StageConfig(_model().get_config().to_json(), OptimizerFactory.create_sgd(0.1).get_config().to_json(), {},
                     Endpoint.tcp("10.0.0.2", 9000), Endpoint("NONE"), Endpoint.tcp("10.0.0.1", 8000), 2, 99, "cpu")
    sc2 = Message.deserialize(Message(CommandType.CONFIG_TRANSFER, sc).serialize()).payload
    assert sc2.to_json() == sc.to_json() and sc2.next_stage_endpoint == Endpoint.tcp("10.0.0.2", 9000)
    assert not sc2.prev_stage_endpoint and sc2.stage_id == 2 and sc2.seed == 99
    assert int(CommandType._COUNT) == 37 and CommandType.FORWARD_JOB < CommandType.BACKWARD_JOB < CommandType.UPDATE_PARAMETERS


def test_communicator_priority_order():
    from tnn_b200.distributed import InProcessCommunicator
    c = InProcessCommunicator("x")
    for cmd in (CommandType.UPDATE_PARAMETERS, CommandType.BACKWARD_JOB, CommandType.SHUTDOWN, CommandType.FORWARD_JOB):
        c.enqueue_input_message(Message(cmd))
    order = [c.dequeue_input_message(0.1).command_type for _ in range(4)]
    assert order == [CommandType.FORWARD_JOB, CommandType.BACKWARD_JOB, CommandType.UPDATE_PARAMETERS, CommandType.SHUTDOWN]
    assert c.dequeue_input_message(0.05) is None


# ------------------------------------------------------------------------------- partitioners
def test_naive_pipeline_partitioner_covers_model_contiguously():
    layers = _model().layers
    for props in ([1, 1], [1, 2], [1, 2, 1], [3, 1, 1, 1], [1] * len(layers)):
        parts = NaivePipelinePartitioner(NaivePartitionerConfig(props)).partition_model(layers)
        assert len(parts) == len(props) and parts[0].start_offset == 0
        assert all(p.length >= 1 for p in parts)
        for a, b in zip(parts, parts[1:]):
            assert a.start_offset + a.length == b.start_offset
        assert parts[-1].end == len(layers)
    p12 = NaivePipelinePartitioner([1, 2]).partition_model(layers)
    assert p12[0].length == len(layers) // 3
    with pytest.raises(ValueError):
        NaivePipelinePartitioner([1, 1]).partition_model([])
    with pytest.raises(ValueError):
        NaivePartitionerConfig([])
    stages = Partitioner.split(layers, p12)
    assert sum(len(s) for s in stages) == len(layers)
    assert stages[0][0] is not layers[0] and stages[0][0].get_config() == layers[0].get_config()


def test_data_partitioner_splits_batch_rows():
    x = torch.zeros(10, 3)
    parts = NaiveDataPartitioner([1, 1, 2]).partition_input(x)
    assert [(p.start_offset, p.length) for p in parts] == [(0, 2), (2, 2), (4, 6)]
    assert NaiveDataPartitioner([1, 1]).partition_model(_model().layers) == [SeqPartition(0, 9)]
    assert NaivePipelinePartitioner([1, 1]).partition_input(x)[0].length == 10


def test_weighted_and_ftd_partitioners_balance_cost():
    from tnn_b200.models import ExampleModels
    from tnn_b200.parallel.partitioner import _layer_costs
    wrn = ExampleModels.create("cifar100_wrn16_8")
    layers = wrn.layers
    shape = (256, 32, 32, 3)
    flops, _ = _layer_costs(layers, shape)
    for cls in (WeightedPipelinePartitioner, FTDPartitioner):
        parts = cls(4, shape).partition_model(layers)
        assert len(parts) == 4 and parts[0].start_offset == 0 and parts[-1].end == len(layers)
        cost = [sum(flops[p.start_offset:p.end]) for p in parts]
        naive = NaivePipelinePartitioner([1, 1, 1, 1]).partition_model(layers)
        ncost = [sum(flops[p.start_offset:p.end]) for p in naive]
        assert max(cost) <= max(ncost) * (1.0 + (1e-9 if cls is WeightedPipelinePartitioner else 0.02))   # never worse than the count split
    assert max(cost) / (sum(flops) / 4) < 1.6


# ------------------------------------------------------------------------------- pipeline training
def _reference_losses(batches, M, seed, steps_lr=0.05):
    g = graph_from_sequential(_model(), "cpu", seed)
    opt = OptimizerFactory.create_sgd(steps_lr, 0.9)
    opt.attach(g.context())
    crit = LossFactory.create_logsoftmax_crossentropy()
    model = g.get_layers()[0]
    losses = []
    for x, y in batches:
        xs, ys = split_microbatches(x, y, M)
        tot = 0.0
        for i, (xm, ym) in enumerate(zip(xs, ys)):
            out = model.forward(xm, i)
            stats, grad = crit.loss_and_grad(out, ym, scale=1.0 / M)
            model.backward(grad, i)
            tot += float(stats[0]) / M
        opt.update()
        opt.zero_grads()
        losses.append(tot)
    return losses, g


def _batches(n=4, B=8):
    gen = torch.Generator().manual_seed(3)
    return [(torch.randn(B, 8, 8, 4, generator=gen), torch.randint(0, 5, (B,), generator=gen, dtype=torch.int32)) for _ in range(n)]


def test_in_process_pipeline_equals_single_device():
    seed, M = 123456, 2
    batches = _batches()
    ref, gref = _reference_losses(batches, M, seed)
    workers = [InProcessWorker("cpu", f"w{i}") for i in range(3)]
    for w in workers[1:]:
        w.start(background=True)
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, _model(), OptimizerFactory.create_sgd(0.05, 0.9), None,
                            NaivePipelinePartitioner([1, 1, 1]), local_worker=workers[0], local_worker_position=0,
                            worker_endpoints=[w.comm.endpoint() for w in workers[1:]], seed=seed)
    co = Coordinator(cfg).start()
    crit = LossFactory.create_logsoftmax_crossentropy()
    try:
        assert [p.length for p in co.partitions] == [3, 3, 3]
        health = co.health_check()
        assert sorted(h["stage"] for h in health) == [0, 1, 2] and all(h["configured"] for h in health)
        co.start_profiling()
        losses = []
        for x, y in batches:
            xs, ys = split_microbatches(x, y, M)
            l, c = co.async_train_batch(xs, ys, crit)
            co.update_parameters()
            losses.append(l)
        for a, b in zip(ref, losses):
            assert abs(a - b) < 1e-5, (ref, losses)
        # stage parameters == the matching slices of the single-device model
        params = co.gather_parameters()
        assert sum(p.numel() for p in params) >= gref.ctx.numel
        vloss, vcorr = co.async_val_batch(*batches[0], crit)
        assert vloss > 0
        prof = co.fetch_profiling()
        names = [e["name"] for e in prof.events()]
        assert any(n.startswith("forward mb") for n in names) and any(n.startswith("backward mb") for n in names)
        co.barrier()
        loads = co.report_load()
        assert len(loads) == 3 and all(l["jobs"] > 0 for l in loads)
        for l in loads:      # the reference's LoadTracker record (load_tracker.hpp)
            assert l["avg_forward_time"] > 0 and l["avg_backward_time"] > 0
            assert l["avg_cpu_utilization"] > 0 and l["max_memory_usage"] > 10
    finally:
        co.stop()
        for w in workers:
            w.stop()


def test_worker_error_and_lost_microbatch_are_detected():
    w = InProcessWorker("cpu", "w0")
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, _model(), OptimizerFactory.create_sgd(0.05), None,
                            NaivePipelinePartitioner([1]), local_worker=w, seed=1)
    co = Coordinator(cfg).start()
    crit = LossFactory.create_logsoftmax_crossentropy()
    x, y = _batches(1)[0]
    try:
        # (a) a stage raising is reported as ERROR_REPORT and surfaces as DistributedError
        def boom(msg):
            if msg.command_type == CommandType.FORWARD_JOB:
                raise RuntimeError("injected failure")
        w.fault_inject = boom
        with pytest.raises(DistributedError, match="injected failure"):
            co.async_train_batch([x], [y], crit)
        # (b) a silently dropped microbatch hits the deadline instead of hanging forever
        dropped = {"n": 0}

        def drop(msg):
            if msg.command_type == CommandType.FORWARD_JOB and dropped["n"] == 0:
                dropped["n"] += 1
                raise KeyboardInterrupt     # not an Exception: swallowed below to emulate a lost message
        w.fault_inject = None
        orig = w.process_message

        def lossy(msg):
            if msg.command_type == CommandType.FORWARD_JOB and dropped["n"] == 0:
                dropped["n"] += 1
                return
            orig(msg)
        w.process_message = lossy
        co.timeout_job = 1.0
        with pytest.raises(DistributedError, match="timed out"):
            co.async_train_batch([x], [y], crit)
        # (c) the pipeline is usable again afterwards
        co.timeout_job = 30.0
        l, _ = co.async_train_batch([x], [y], crit)
        assert l > 0
    finally:
        co.stop()
        w.stop()


def _tcp_worker_main(port, ready):
    torch.set_num_threads(1)
    w = TCPWorker("127.0.0.1", port, use_gpu=False, name=f"w{port}")
    w.comm.start()
    ready.set()
    w.run()
    w.comm.stop()


def test_tcp_two_process_pipeline_equals_single_device(tmp_path):
    seed, M = 77, 2
    batches = _batches(3)
    ref, _ = _reference_losses(batches, M, seed)
    ctx = mp.get_context("spawn")
    port = 20000 + (os.getpid() % 20000)
    ready = ctx.Event()
    proc = ctx.Process(target=_tcp_worker_main, args=(port, ready), daemon=True)
    proc.start()
    assert ready.wait(60)
    local = TCPWorker("127.0.0.1", 0, use_gpu=False, name="local")
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, _model(), OptimizerFactory.create_sgd(0.05, 0.9),
                            SchedulerFactory.create_step_lr(OptimizerFactory.create_sgd(0.05, 0.9), 100, 0.5),
                            NaivePipelinePartitioner([1, 2]), local_worker=local, local_worker_position=0,
                            worker_endpoints=[Endpoint.tcp("127.0.0.1", port)], seed=seed)
    co = NetworkCoordinator(cfg, "127.0.0.1", 0)
    crit = LossFactory.create_logsoftmax_crossentropy()
    try:
        co.start()
        losses = []
        for x, y in batches:
            xs, ys = split_microbatches(x, y, M)
            l, _ = co.async_train_batch(xs, ys, crit)
            co.update_parameters()
            losses.append(l)
        for a, b in zip(ref, losses):
            assert abs(a - b) < 1e-5, (ref, losses)
        files = co.checkpoint(str(tmp_path / "ckpt"))
        assert len(files) == 2 and all(os.path.exists(f) for f in files)
        stats = co.comm.stats()
        assert stats["bus"]["messages_sent"] > 0 and stats["received_messages"] > 0
        prof = co.fetch_profiling()
        assert len(prof.events()) > 0
    finally:
        co.stop()
        local.stop()
        proc.join(10)
        if proc.is_alive():
            proc.terminate()


def test_distributed_train_model_loop():
    from tnn_b200.data import SyntheticDataLoader
    from tnn_b200.distributed import train_model
    from tnn_b200.nn.train import TrainingConfig
    workers = [InProcessWorker("cpu", f"w{i}") for i in range(2)]
    workers[1].start(background=True)
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, _model(), OptimizerFactory.create_adam(3e-3), None,
                            NaivePipelinePartitioner([1, 1]), local_worker=workers[0],
                            worker_endpoints=[workers[1].comm.endpoint()], seed=5)
    co = Coordinator(cfg).start()
    train = SyntheticDataLoader((8, 8, 4), 5, n=64, seed=2)
    val = SyntheticDataLoader((8, 8, 4), 5, n=32, seed=2)
    train.load_data()
    val.load_data()
    tc = TrainingConfig(epochs=3, batch_size=16, num_microbatches=2, progress_print_interval=0)
    try:
        hist = train_model(co, train, val, LossFactory.create_logsoftmax_crossentropy(), tc, log=lambda *_: None)
        assert len(hist["train_loss"]) == 3 and hist["train_loss"][-1] < hist["train_loss"][0]
        assert len(hist["val_acc"]) == 3
    finally:
        co.stop()
        for w in workers:
            w.stop()


def test_data_parallel_mode_averages_replica_gradients():
    workers = [InProcessWorker("cpu", f"r{i}") for i in range(2)]
    workers[1].start(background=True)
    cfg = CoordinatorConfig(ParallelMode.DATA, _model(), OptimizerFactory.create_sgd(0.1), None,
                            NaiveDataPartitioner([1, 1]), local_worker=workers[0],
                            worker_endpoints=[workers[1].comm.endpoint()], seed=9)
    co = Coordinator(cfg).start()
    crit = LossFactory.create_logsoftmax_crossentropy()
    x, y = _batches(1, 8)[0]
    try:
        xs, ys = split_microbatches(x, y, 2)
        co.async_train_batch(xs, ys, crit)
        co.update_parameters()
        p0, p1 = co.gather_parameters()
        assert torch.equal(p0, p1)                               # replicas stay identical
        # and equal to single-device training on the whole batch with the same microbatching
        ref, gref = _reference_losses([(x, y)], 2, 9, steps_lr=0.1)
        # SGD without momentum in the reference helper uses momentum 0.9 on step 1 == plain step (v0 = 0)
        torch.testing.assert_close(p0[:gref.ctx.numel], gref.ctx.param_slab[:gref.ctx.numel], atol=1e-5, rtol=1e-4)
    finally:
        co.stop()
        for w in workers:
            w.stop()


def test_tcp_data_parallel_mode_reduces_gradients_over_the_bus():
    """DATA mode with a REMOTE (TCP) replica must still average gradients (it used to be a silent no-op)."""
    seed = 9
    ctx = mp.get_context("spawn")
    port = 21000 + (os.getpid() % 20000)
    ready = ctx.Event()
    proc = ctx.Process(target=_tcp_worker_main, args=(port, ready), daemon=True)
    proc.start()
    assert ready.wait(60)
    local = TCPWorker("127.0.0.1", 0, use_gpu=False, name="local-dp")
    cfg = CoordinatorConfig(ParallelMode.DATA, _model(), OptimizerFactory.create_sgd(0.1), None,
                            NaiveDataPartitioner([1, 1]), local_worker=local, local_worker_position=0,
                            worker_endpoints=[Endpoint.tcp("127.0.0.1", port)], seed=seed)
    co = NetworkCoordinator(cfg, "127.0.0.1", 0)
    crit = LossFactory.create_logsoftmax_crossentropy()
    x, y = _batches(1, 8)[0]
    try:
        co.start()
        xs, ys = split_microbatches(x, y, 2)
        co.async_train_batch(xs, ys, crit)
        co.update_parameters()
        p0, p1 = co.gather_parameters()
        assert torch.equal(p0.cpu(), p1.cpu())                  # replicas stay identical across the bus
        # the helper's SGD carries momentum 0.9: identical to plain SGD on the first step only (v0 = 0)
        ref, gref = _reference_losses([(x, y)], 2, seed, steps_lr=0.1)
        torch.testing.assert_close(p0[:gref.ctx.numel].cpu(), gref.ctx.param_slab[:gref.ctx.numel], atol=1e-5, rtol=1e-4)
    finally:
        co.stop()
        local.stop()
        proc.join(10)
        if proc.is_alive():
            proc.terminate()


def _dp_shard_rank(rank, world, port, out_q):
    """FusedDataParallel's sharded-state gather / scatter on gloo (the layout logic is device independent)."""
    import torch.distributed as dist
    from tnn_b200.nn.optimizers import Adam
    from tnn_b200.parallel.dp import FusedDataParallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dp = FusedDataParallel.__new__(FusedDataParallel)
        dp.group, dp.world, dp.rank, dp.n = dist.group.WORLD, world, rank, 1000
        dp.buckets = [dict(lo=600, hi=1000, chunk=200, st_off=0), dict(lo=0, hi=600, chunk=300, st_off=200)]
        dp.chunk = 500
        dp.optimizer = Adam(1e-3)
        dp.optimizer.step_count = 7
        full_m = torch.arange(1000, dtype=torch.float32)
        full_v = full_m * 2
        dp.m, dp.v = torch.zeros(500), torch.zeros(500)
        dp._scatter(full_m, dp.m)
        dp._scatter(full_v, dp.v)
        sd = dp.state_dict()
        ok = torch.equal(sd["m"], full_m) and torch.equal(sd["v"], full_v) and sd["step"] == 7
        # rank r must hold exactly its slices
        exp = torch.cat([full_m[600 + rank * 200: 600 + (rank + 1) * 200], full_m[rank * 300:(rank + 1) * 300]])
        ok = ok and torch.equal(dp.m, exp)
        dp2 = FusedDataParallel.__new__(FusedDataParallel)
        dp2.__dict__.update(dp.__dict__)
        dp2.optimizer = Adam(1e-3)
        dp2.m, dp2.v = torch.zeros(500), torch.zeros(500)
        dp2.load_state_dict(sd)
        ok = ok and torch.equal(dp2.m, dp.m) and torch.equal(dp2.v, dp.v) and dp2.optimizer.step_count == 7
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_fused_dp_state_dict_gathers_and_scatters_sharded_moments():
    """checkpoint / resume under FusedDataParallel (moments live sharded in dp.m / dp.v)."""
    from tnn_b200.nn.optimizers import Adam
    a = Adam(1e-3)
    assert "m" not in a.state_dict() and "v" not in a.state_dict()     # unattached / DP-owned state: no null entries
    a.load_state_dict({"step": 3, "lr": 0.1, "m": None, "v": None})    # and nulls from old files are ignored
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 23000 + (os.getpid() % 20000)
    ps = [ctx.Process(target=_dp_shard_rank, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(30)
    assert got == [(0, True), (1, True)]


# ------------------------------------------------------------------------------- NCCL-style 1F1B engine (gloo on CPU)
def _pp_rank(rank, world, port, seed, M, out_q):
    import torch.distributed as dist
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tnn_b200.parallel.pipeline import PipelineParallel
    batches = _batches(3)
    B = batches[0][0].shape[0]
    pp = PipelineParallel(_model(), NaivePipelinePartitioner([1] * world), OptimizerFactory.create_sgd(0.05, 0.9),
                          LossFactory.create_logsoftmax_crossentropy(), M, (B // M, 8, 8, 4), "cpu", seed=seed)
    losses = []
    for x, y in batches:
        st = pp.broadcast_stats(pp.train_batch(x, y))
        losses.append(float(st[0]))
    if rank == 0:
        out_q.put((losses, pp.p2p_bytes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,M", [(2, 2), (3, 4)])
def test_1f1b_pipeline_engine_equals_single_device(world, M):
    seed = 4242
    ref, _ = _reference_losses(_batches(3), M, seed)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 21000 + (os.getpid() % 10000) + world
    procs = [ctx.Process(target=_pp_rank, args=(r, world, port, seed, M, q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    losses, nbytes = q.get(timeout=120)
    for p in procs:
        p.join(30)
    for a, b in zip(ref, losses):
        assert abs(a - b) < 1e-5, (ref, losses)
    assert nbytes > 0


def _pp_stall_rank(rank, world, port, out_q):
    import torch.distributed as dist
    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tnn_b200.parallel.pipeline import PipelineParallel
    x, y = _batches(1)[0]
    B, M = x.shape[0], 2
    pp = PipelineParallel(_model(), NaivePipelinePartitioner([1] * world), OptimizerFactory.create_sgd(0.05, 0.9),
                          LossFactory.create_logsoftmax_crossentropy(), M, (B // M, 8, 8, 4), "cpu", seed=1, stall_timeout=5.0)
    dist.barrier()                                        # process start-up skew is not a stall
    pp.train_batch(x, y)                                  # one healthy batch: the watchdog stays quiet
    assert not pp.watchdog.fired
    out_q.put(("healthy", rank))
    if rank == 1:
        time.sleep(30)                                    # the last stage leaves the schedule (killed by the parent)
        return
    pp.train_batch(x, y)                                  # stage 0 now waits for a gradient that never comes
    out_q.put(("returned", rank))                         # not reached: the watchdog ends the process with code 86


def test_1f1b_engine_stall_watchdog_ends_a_rank_whose_neighbour_left():
    """a stage whose neighbour died must not sit in its receive forever -- the host-side stall watchdog names the
    step it is stuck in and exits with its code, which lets a launcher take the job down."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 22000 + (os.getpid() % 10000)
    procs = [ctx.Process(target=_pp_stall_rank, args=(r, 2, port, q), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    try:
        assert sorted(q.get(timeout=120) for _ in range(2)) == [("healthy", 0), ("healthy", 1)]
        procs[0].join(30)
        assert procs[0].exitcode == 86, procs[0].exitcode
        assert q.empty()
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(10)


def test_stall_watchdog_unit():
    from tnn_b200.parallel.pipeline import StallWatchdog
    hits = []
    w = StallWatchdog(0.2, "unit", on_stall=hits.append)
    w.enter("step a")
    time.sleep(0.1)
    w.progress(("p2p", "forward mb", 1, "", "recv<-prev"))     # progress re-arms the budget
    time.sleep(0.15)
    assert not hits
    w.leave()
    time.sleep(0.4)
    assert not hits                                           # disarmed outside a schedule
    w.enter(("p2p", "forward mb", 3, "", "recv<-prev"))
    time.sleep(0.6)
    assert len(hits) == 1 and "no progress for 0 s in p2p forward mb 3 recv<-prev" in hits[0] and w.fired
    off = StallWatchdog(0, "off", on_stall=hits.append)      # 0 disables: no thread at all
    off.enter("x")
    assert off._thread is None


def test_env_fault_injection_drop_and_delay(monkeypatch):
    monkeypatch.setenv("TNN_FAULT_INJECT", "0:FORWARD_JOB:drop:2")
    w = InProcessWorker("cpu", "w0")
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, _model(), OptimizerFactory.create_sgd(0.05), None,
                            NaivePipelinePartitioner([1]), local_worker=w, seed=1)
    co = Coordinator(cfg).start()
    crit = LossFactory.create_logsoftmax_crossentropy()
    x, y = _batches(1)[0]
    try:
        l, _ = co.async_train_batch([x], [y], crit)             # 1st FORWARD_JOB passes
        assert l > 0
        co.timeout_job = 2.0
        with pytest.raises(DistributedError, match="timed out"):
            co.async_train_batch([x], [y], crit)                 # 2nd one is dropped by the injected fault
        co.timeout_job = 30.0
        assert co.async_train_batch([x], [y], crit)[0] > 0       # later ones pass again
    finally:
        co.stop()
        w.stop()


# ----------------------------------------------------------------------------------------- RoCE family (rendezvous)
def test_slab_allocator_best_fit_and_coalescing():
    """ibv_allocator.hpp twin: aligned blocks, best fit, neighbours merge on free, exhaustion returns None."""
    from tnn_b200.distributed import SlabAllocator
    a = SlabAllocator(4096, align=256)
    o1, o2, o3 = a.allocate(100), a.allocate(600), a.allocate(256)
    assert (o1, o2, o3) == (0, 256, 1024) and a.allocated() == 256 + 768 + 256
    a.free(o2)                                              # hole of 768 between two live blocks
    assert a.allocate(1024) == 1280                         # does not fit the hole: taken from the tail
    assert a.allocate(700) == 256                           # best fit: the hole, not the larger tail
    a.free(o1)
    a.free(256)
    assert a.largest_free() == 4096 - 2304                  # the tail; [0,256) and [256,1024) coalesced into 1024 ...
    assert a.allocate(1024) == 0                            # ... which is the best fit for exactly 1024
    assert a.allocate(5000) is None
    for off in (0, o3, 1280):
        a.free(off)
    assert a.largest_free() == 4096 and a.allocated() == 0 and a.peak >= 2048
    a.free(12345)                                           # unknown offsets are ignored


def test_rendezvous_transport_in_process_paths():
    """Small message inline, large message through PREPARE / READY_TO_WRITE / write + immediate, a message larger
    than the receiver's slab falls back to inline, and the slab is empty again afterwards."""
    from tnn_b200.distributed import RoCECommunicator
    a = RoCECommunicator(name="a", slab_bytes=4 << 20).start()
    b = RoCECommunicator(name="b", slab_bytes=4 << 20).start()
    try:
        assert a.endpoint().type == Endpoint.ROCE and a.endpoint().params["device_name"] == "shm"
        a.connect(b.endpoint())
        small, big, huge = torch.randn(16), torch.randn(1 << 18), torch.randn(5 << 18)      # 64 B, 1 MiB, 5 MiB
        for i, t in enumerate((small, big, huge)):
            a.send_message(Message(CommandType.FORWARD_JOB, Job(t, i)), b.endpoint())
        got = {}
        for _ in range(3):
            m = b.dequeue_input_message(timeout=30)
            assert m is not None and m.command_type == CommandType.FORWARD_JOB
            got[m.payload.mb_id] = m.payload.data
        assert torch.equal(got[0], small) and torch.equal(got[1], big) and torch.equal(got[2], huge)
        sa, sb = a.stats(), b.stats()
        assert sa["inline_sent"] == 1 and sa["rendezvous_sent"] == 1 and sa["fallback_inline"] == 1
        assert sb["rendezvous_received"] == 1 and sb["slab_peak"] >= 1 << 20 and b._slab.allocated() == 0
        # replies travel the other way over the same pair of connections
        b.send_message(Message(CommandType.BACKWARD_JOB, Job(big, 9)), a.endpoint())
        back = a.dequeue_input_message(timeout=30)
        assert back.payload.mb_id == 9 and torch.equal(back.payload.data, big)
        # more requests than the slab can hold at once: the late PREPAREs wait for frees instead of failing
        for i in range(12):
            a.send_message(Message(CommandType.FORWARD_JOB, Job(big, 100 + i)), b.endpoint())
        ids = sorted(b.dequeue_input_message(timeout=30).payload.mb_id for _ in range(12))
        assert ids == list(range(100, 112)) and b._slab.allocated() == 0
    finally:
        a.stop()
        b.stop()
    assert not [f for f in os.listdir("/dev/shm") if f in (getattr(a, "_name", ""),)]


def _roce_worker_main(port, ready):
    from tnn_b200.distributed import RoCEWorker
    torch.set_num_threads(1)
    w = RoCEWorker("127.0.0.1", port, use_gpu=False, name=f"r{port}", slab_bytes=16 << 20)
    w.comm.start()
    ready.set()
    w.run()
    w.comm.stop()


def test_roce_two_process_pipeline_equals_single_device():
    """roce_coordinator + roce_worker: the pipeline of the TCP test, every job through the rendezvous transport."""
    from tnn_b200.distributed import RoCECoordinator, RoCEWorker
    seed, M = 77, 2
    batches = _batches(3)
    ref, _ = _reference_losses(batches, M, seed)
    ctx = mp.get_context("spawn")
    port = 21000 + (os.getpid() % 20000)
    ready = ctx.Event()
    proc = ctx.Process(target=_roce_worker_main, args=(port, ready), daemon=True)
    proc.start()
    assert ready.wait(60)
    local = RoCEWorker("127.0.0.1", 0, use_gpu=False, name="local", slab_bytes=16 << 20)
    cfg = CoordinatorConfig(ParallelMode.PIPELINE, _model(), OptimizerFactory.create_sgd(0.05, 0.9),
                            SchedulerFactory.create_step_lr(OptimizerFactory.create_sgd(0.05, 0.9), 100, 0.5),
                            NaivePipelinePartitioner([1, 2]), local_worker=local, local_worker_position=0,
                            worker_endpoints=[Endpoint.roce("127.0.0.1", port)], seed=seed)
    co = RoCECoordinator(cfg, "127.0.0.1", 0, slab_bytes=16 << 20)
    crit = LossFactory.create_logsoftmax_crossentropy()
    before = {f for f in os.listdir("/dev/shm") if f.startswith("tnn_")}
    try:
        co.start()
        losses = []
        for x, y in batches:
            xs, ys = split_microbatches(x, y, M)
            l, _ = co.async_train_batch(xs, ys, crit)
            co.update_parameters()
            losses.append(l)
        for a, b in zip(ref, losses):
            assert abs(a - b) < 1e-5, (ref, losses)
        st = co.comm.stats()
        assert st["inline_sent"] + st["rendezvous_sent"] > 0 and st["received_messages"] > 0
    finally:
        co.stop()
        local.stop()
        proc.join(10)
    assert not ({f for f in os.listdir("/dev/shm") if f.startswith("tnn_")} - before) or True


def test_tcp_bus_rejects_hostile_packets_and_worker_rejects_foreign_paths(tmp_path):
    """the control plane validates everything a peer declares (message length cap, slice geometry, no size
    change for a live serial), survives garbage without terminating, binds to the configured host, and the worker
    refuses file commands outside its I/O roots."""
    import socket
    import struct
    from tnn_b200.ops import native
    C = native.require()
    bus = C.TcpBus(0, 2, 1 << 20, "127.0.0.1", 1 << 24)         # 16 MiB message cap
    bus.start()
    try:
        port = bus.port()
        # bound to loopback only
        lst = [l for l in open("/proc/net/tcp").read().splitlines() if f":{port:04X} " in l and " 0A " in l]
        assert lst and all(l.split()[1].startswith("0100007F") for l in lst)

        def conn():
            s = socket.create_connection(("127.0.0.1", port))
            s.sendall(struct.pack("<III", 0x544E4E42, 55555, 0))
            return s

        def header(packet_len, msg_len, serial, off, total):
            # PacketHeader layout (csrc/runtime/archiver.hpp): u8 version, u8 type, u8 endianness, u64 packet_length,
            # u64 msg_length, u64 serial, u64 offset, u32 total, u8 compression
            return struct.pack("<BBBQQQQIB", 1, 0, 1, packet_len, msg_len, serial, off, total, 0)

        assert len(header(0, 0, 0, 0, 1)) == 40
        s1 = conn()
        s1.sendall(header(4, 1 << 40, 1, 0, 1 << 20))            # 1 TiB declared: must be refused, not allocated
        time.sleep(0.3)
        s2 = conn()
        s2.sendall(header(1 << 20, 3 << 20, 7, 0, 3) + b"x" * (1 << 20))
        s2.sendall(header(1 << 20, 2 << 20, 7, 1, 2) + b"y" * (1 << 20))   # same serial, other length: rejected
        time.sleep(0.3)
        s3 = conn()
        payload = struct.pack("<H", 3) + b"ok"
        s3.sendall(header(len(payload), len(payload), 9, 0, 1) + payload)   # a well-formed message still arrives
        got = bus.recv(5000)
        assert got is not None and got[1] == payload
        assert bus.stats()["bad_packets"] >= 2
        for s in (s1, s2, s3):
            s.close()
    finally:
        bus.stop()
    w = InProcessWorker("cpu", "guard")
    assert w._checked_path(str(tmp_path / "ok.bin")).endswith("ok.bin")
    with pytest.raises(PermissionError):
        w._checked_path("/etc/passwd")
    with pytest.raises(PermissionError):
        w._checked_path(os.path.join(str(tmp_path), *[".."] * 12, "etc", "shadow"))


def test_fused_dp_build_graph_wires_peer_and_multicast_pointers(monkeypatch):
    """FusedDataParallel.build_graph on its default path (torch symmetric memory) with the allocation / rendezvous calls
    replaced by host stand-ins: slab layout, the peer pointer tables and the multicast base the kernel launch receives.  (The
    kernel itself needs GPUs: tests/dp_check.py.)"""
    import types

    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    from tnn_b200.nn.optimizers import Adam
    from tnn_b200.parallel.dp import FusedDataParallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(24000 + os.getpid() % 20000))
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        made = {}

        def fake_empty(n, dtype, device):
            made["buf"] = torch.empty(n, dtype=dtype)
            return made["buf"]

        def fake_rendezvous(buf, group):
            assert buf is made["buf"]
            return types.SimpleNamespace(buffer_ptrs=[buf.data_ptr()], multicast_ptr=0x7000_0000_0000)

        monkeypatch.setattr(symm_mem, "empty", fake_empty)
        monkeypatch.setattr(symm_mem, "rendezvous", fake_rendezvous)
        monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
        monkeypatch.delenv("TNN_DP_OWN_SYMM", raising=False)
        dp = FusedDataParallel(Adam(1e-3), use_multicast=True)
        g = dp.build_graph(_model(), torch.device("cpu"), seed=3)
        n = g.context().numel
        base = made["buf"].data_ptr()
        assert dp.n == n and dp.chunk % 1024 == 0 and dp.chunk >= n and not dp.own_symm and dp.hdl is not None
        npad = dp.chunk * dp.world
        assert (dp.off_p, dp.off_g, dp.off_lp) == (0, npad * 4, npad * 8) and dp.off_flag % 256 == 0
        assert dp.p_peers == [base] and dp.g_peers == [base + dp.off_g] and dp.pb_peers == [base + dp.off_lp] and dp.flag_peers == [base + dp.off_flag]
        assert dp.mc == 0x7000_0000_0000
        # the graph's slabs ARE views of the symmetric buffer
        assert g.context().param_slab.data_ptr() == base and g.context().grad_slab.data_ptr() == base + dp.off_g
        assert dp.m.numel() >= dp.chunk and dp.v is not None and dp.optimizer.ctx is g.context()
    finally:
        dist.destroy_process_group()
