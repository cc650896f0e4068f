#!/usr/bin/env python
"""Ping-pong bandwidth of the TCP message bus between two endpoints (reference:
benchmarks/tcp_communicator_benchmark.cpp:131-222): each side seeds 4 FORWARD_JOB messages carrying
a [256,32,32,128] fp32 tensor (128 MiB) and bounces whatever arrives back to its peer for
``--seconds``; bandwidth = received bytes / elapsed.

  terminal 1:  python benchmarks/tcp_communicator_bench.py --port 8001 --peer-port 8002
  terminal 2:  python benchmarks/tcp_communicator_bench.py --port 8002 --peer-port 8001
  one shot:    python benchmarks/tcp_communicator_bench.py --self-test [--mb 16 --seconds 3]

CPU-only; the data path is the native C++ bus (csrc/runtime/tcp_bus.hpp)."""
import argparse
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import emit  # noqa: E402

import torch  # noqa: E402

from tnn_b200.distributed import CommandType, Endpoint, Job, Message, RoCECommunicator, TCPCommunicator  # noqa: E402


def run_side(host, port, peer_host, peer_port, threads, seconds, shape, inflight=4, quiet=False, transport="tcp"):
    cls = RoCECommunicator if transport == "roce" else TCPCommunicator      # roce = shared-memory rendezvous (roce_communicator_benchmark.cpp)
    comm = cls(host, port, num_io_threads=threads, name=f"bench{port}")
    comm.start()
    peer = Endpoint.tcp(peer_host, peer_port)
    for attempt in range(120):
        try:
            if comm.connect(peer, timeout=2.0):
                break
        except RuntimeError:
            pass
        print("Retrying connection to peer...", file=sys.stderr)
        time.sleep(0.5)
    else:
        raise SystemExit(f"peer {peer_host}:{peer_port} never came up")
    master = torch.randn(*shape, generator=torch.Generator().manual_seed(123456))
    nbytes = master.numel() * 4
    for _ in range(inflight):
        comm.send_message(Message(CommandType.FORWARD_JOB, Job(master.clone(), 10)), peer)
    t0 = time.perf_counter()
    received = 0
    while time.perf_counter() - t0 < seconds:
        m = comm.dequeue_input_message(timeout=max(0.01, seconds - (time.perf_counter() - t0)))
        if m is None:
            continue
        if m.command_type != CommandType.FORWARD_JOB:
            continue
        if received == 0:
            assert torch.equal(m.payload.data, master), "payload corrupted in flight"
        received += 1
        try:
            comm.send_message(m, peer)
        except RuntimeError:
            # the two sides start their clocks up to one connect-retry interval apart: the peer's window may end (and its
            # sockets close) while this side still bounces -- that ends this side's measurement, it is not an error
            break
    dt = time.perf_counter() - t0
    time.sleep(0.3)                    # let the peer drain before the sockets close
    stats = comm.stats()
    comm.stop()
    res = dict(bench=f"{transport}_communicator", port=port, messages=received, bytes=received * nbytes, seconds=round(dt, 3),
               MBps=round(received * nbytes / dt / 1e6, 1), message_MB=round(nbytes / 2 ** 20, 1), io_threads=threads,
               packets_sent=(stats.get("bus") or {}).get("packets_sent"))
    if not quiet:
        print(f"Total messages received: {received}\nTotal bytes received: {received * nbytes} bytes\n"
              f"Total time taken: {dt:.2f} seconds\nBandwidth: {res['MBps']} MB/s")
    return res


def _child(port, peer_port, threads, seconds, shape, q, transport):
    q.put(run_side("127.0.0.1", port, "127.0.0.1", peer_port, threads, seconds, shape, quiet=True, transport=transport))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int)
    ap.add_argument("--peer-host", default="127.0.0.1")
    ap.add_argument("--peer-port", type=int)
    ap.add_argument("--num-threads", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--mb", type=int, default=128, help="message payload in MiB (reference: 128)")
    ap.add_argument("--self-test", action="store_true", help="spawn the peer locally")
    ap.add_argument("--transport", choices=["tcp", "roce"], default="tcp",
                    help="roce = the rendezvous transport (reference: roce_communicator_benchmark.cpp)")
    a = ap.parse_args()
    shape = (a.mb * 2, 32, 32, 128)                  # mb MiB of fp32; 128 -> the reference's [256,32,32,128]
    if a.self_test:
        p0 = 24000 + os.getpid() % 3000
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        child = ctx.Process(target=_child, args=(p0 + 1, p0, a.num_threads, a.seconds, shape, q, a.transport), daemon=True)
        child.start()
        mine = run_side("127.0.0.1", p0, "127.0.0.1", p0 + 1, a.num_threads, a.seconds, shape, quiet=True, transport=a.transport)
        other = q.get(timeout=a.seconds + 60)
        child.join(10)
        emit(**mine)
        emit(**other)
        return 0
    if not a.port or not a.peer_port:
        ap.error("--port and --peer-port are required (or --self-test)")
    print(f"Communicator Benchmark Configuration\nListen host: {a.host}\nListen port: {a.port}\nPeer host: {a.peer_host}\n"
          f"Peer port: {a.peer_port}\nWorker threads: {a.num_threads}")
    emit(**run_side(a.host, a.port, a.peer_host, a.peer_port, a.num_threads, a.seconds, shape, transport=a.transport))
    return 0


if __name__ == "__main__":
    sys.exit(main())
