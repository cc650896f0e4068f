#!/usr/bin/env python
"""Serialisation + communicator micro-benchmark (reference: benchmarks/serialization_benchmark.cpp,
tcp_communicator_benchmark.cpp): Message(Job{tensor}) archive / de-archive throughput and the
TCP bus throughput between two processes on localhost.  Runs on a CPU-only box."""
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit  # noqa: E402

import torch  # noqa: E402

from tnn_b200.distributed import CommandType, Endpoint, Job, Message, TCPCommunicator  # noqa: E402


def _echo(port, n_msgs, ready):
    c = TCPCommunicator("127.0.0.1", port, name="echo")
    c.start()
    ready.set()
    for _ in range(n_msgs):
        m = c.dequeue_input_message(timeout=60)
        host, p = m.sender.split(":")
        c.send_message(Message(CommandType.BACKWARD_COMPLETE), Endpoint.tcp(host, int(p)))
    time.sleep(0.2)
    c.stop()


if __name__ == "__main__":
    for mb in (1, 32, 128):
        t = torch.randn(mb << 18)          # mb MiB of fp32
        m = Message(CommandType.FORWARD_JOB, Job(t, 3))
        t0 = time.perf_counter()
        reps = max(2, 256 // mb)
        for _ in range(reps):
            data = m.serialize()
        ts = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            Message.deserialize(data)
        td = (time.perf_counter() - t0) / reps
        emit(bench="serialization", MB=mb, serialize_gbs=len(data) / ts / 1e9, deserialize_gbs=len(data) / td / 1e9)
    ctx = mp.get_context("spawn")
    port, n = 23000 + os.getpid() % 2000, 8
    ready = ctx.Event()
    p = ctx.Process(target=_echo, args=(port, n, ready), daemon=True)
    p.start()
    ready.wait(60)
    c = TCPCommunicator("127.0.0.1", 0, name="bench")
    c.start()
    peer = Endpoint.tcp("127.0.0.1", port)
    c.connect(peer)
    payload = Message(CommandType.FORWARD_JOB, Job(torch.randn(32 << 18), 0))      # 32 MiB
    nbytes = payload.byte_size()
    t0 = time.perf_counter()
    for _ in range(n):
        c.send_message(payload, peer)
        assert c.dequeue_input_message(timeout=60) is not None
    dt = (time.perf_counter() - t0) / n
    emit(bench="tcp_communicator", MB=nbytes / 2 ** 20, round_trip_ms=dt * 1e3, gbs=nbytes / dt / 1e9, bus=c.stats()["bus"])
    c.stop()
    p.join(10)
