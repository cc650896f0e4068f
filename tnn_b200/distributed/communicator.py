"""Communicators: the message bus between coordinator and workers.

Reference: include/distributed/communicator.hpp:30-196 (per-CommandType queues, dequeue scans
types in enum order so FORWARD_JOB beats BACKWARD_JOB beats everything else, notify callback,
IN_PROCESS short-circuit, profile counters), tcp_communicator.hpp, roce_communicator.hpp.

* ``InProcessCommunicator`` hands Message objects over by reference (tensors are NOT serialised:
  a GPU tensor stays on its device; used for the local worker and for one-process multi-GPU runs).
* ``TCPCommunicator`` serialises with the C++ archiver and ships packets through the C++
  ``TcpBus`` (multi-socket, 4 MiB slices, reassembly, priority dequeue).  Every send/receive is
  recorded as a COMMUNICATION event in the global profiler, as the reference does per packet.
* The RoCE family (``RoCECommunicator`` / ``RoCEWorker`` / ``RoCECoordinator``) is the one-sided rendezvous
  transport of ``rendezvous.py``: registered shared-memory slab, PREPARE / READY_TO_WRITE / write + immediate,
  control frames on the TCP bus.  (The tensor data path of the B200 engines is NVLink: parallel/pipeline.py and
  parallel/dp.py.)
"""
from __future__ import annotations

import atexit
import os
import threading
import time
from collections import deque

from ..ops import native
from ..utils.profiler import EventType, GlobalProfiler, Profiler
from .message import CommandType, Endpoint, Message


class Communicator:
    def __init__(self, name: str = "comm"):
        self.name = name
        self._queues: dict[int, deque] = {int(c): deque() for c in CommandType}
        self._cv = threading.Condition()
        self._callback = None
        self.sent_messages = 0
        self.received_messages = 0
        self.sent_bytes = 0
        self.received_bytes = 0
        self._endpoint: Endpoint | None = None

    # ---- identity
    def endpoint(self) -> Endpoint:
        if self._endpoint is None:
            self._endpoint = Endpoint.in_process(self)
        return self._endpoint

    # ---- receive side
    def set_callback(self, fn):
        self._callback = fn

    def enqueue_input_message(self, msg: Message):
        with self._cv:
            self._queues[int(msg.command_type)].append(msg)
            self.received_messages += 1
            self._cv.notify_all()
        if self._callback is not None:
            self._callback()

    def has_input_message(self) -> bool:
        with self._cv:
            return any(self._queues[k] for k in self._queues)

    def _pop_locked(self, types=None):
        keys = sorted(self._queues) if types is None else sorted(int(t) for t in types)
        for k in keys:
            q = self._queues[k]
            if q:
                return q.popleft()
        return None

    def dequeue_input_message(self, timeout: float | None = None, types=None) -> Message | None:
        """Highest-priority pending message (lowest CommandType value), waiting up to ``timeout``."""
        deadline = None if timeout is None else time.monotonic() + timeout
        with self._cv:
            while True:
                m = self._pop_locked(types)
                if m is not None:
                    return m
                if deadline is None:
                    self._cv.wait(0.5)
                else:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        return None
                    self._cv.wait(left)

    def pending(self, cmd: CommandType) -> int:
        with self._cv:
            return len(self._queues[int(cmd)])

    # ---- send side
    def connect(self, endpoint: Endpoint, timeout: float = 10.0) -> bool:
        return True

    def send_message(self, msg: Message, endpoint: Endpoint):
        if endpoint.type == Endpoint.IN_PROCESS:
            peer = endpoint.communicator()
            if peer is None:
                raise RuntimeError(f"{self.name}: in-process endpoint {endpoint} is gone")
            msg.sender = self.endpoint().id()
            self.sent_messages += 1
            peer.enqueue_input_message(msg)
            return
        self._send_remote(msg, endpoint)

    def _send_remote(self, msg: Message, endpoint: Endpoint):
        raise RuntimeError(f"{self.name}: no transport for {endpoint}")

    def start(self):
        return self

    def stop(self):
        pass

    def stats(self) -> dict:
        return dict(sent_messages=self.sent_messages, received_messages=self.received_messages,
                    sent_bytes=self.sent_bytes, received_bytes=self.received_bytes)


class InProcessCommunicator(Communicator):
    pass


class TCPCommunicator(Communicator):
    """TCP transport (also serves ROCE endpoints).  ``device`` = where received Job tensors land."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0, num_io_threads: int = 4, skts_per_endpoint: int = 4,
                 device=None, name: str = "tcp"):
        super().__init__(name)
        self.host = host
        self.device = device
        self._bus = native.require().TcpBus(int(port), int(skts_per_endpoint), 4 << 20, str(host),
                                            int(os.environ.get("TNN_TCP_MAX_MESSAGE", 8 << 30)))
        self._pump: threading.Thread | None = None
        self._running = False
        self._port = port
        self.num_io_threads = num_io_threads

    def start(self):
        if self._running:
            return self
        self._bus.start()
        self._port = self._bus.port()
        self._endpoint = Endpoint.tcp(self.host, self._port)
        self._running = True
        self._pump = threading.Thread(target=self._receive_loop, name=f"{self.name}-rx", daemon=True)
        self._pump.start()
        # a daemon thread parked inside the C++ bus (GIL released) must not outlive the interpreter
        atexit.register(self.stop)
        return self

    start_server = start

    def port(self) -> int:
        return self._port

    def endpoint(self) -> Endpoint:
        if self._endpoint is None or self._endpoint.type == Endpoint.IN_PROCESS:
            self._endpoint = Endpoint.tcp(self.host, self._port)
        return self._endpoint

    def connect(self, endpoint: Endpoint, timeout: float = 10.0) -> bool:
        if endpoint.type == Endpoint.IN_PROCESS:
            return True
        self._bus.connect(endpoint.params["host"], int(endpoint.params["port"]), int(timeout * 1000))
        return True

    def _key(self, endpoint: Endpoint) -> str:
        host = endpoint.params["host"]
        if host == "localhost":
            host = "127.0.0.1"
        return f"{host}:{endpoint.params['port']}"

    def _send_remote(self, msg: Message, endpoint: Endpoint):
        t0 = Profiler.now_ns()
        data = msg.serialize()
        t1 = Profiler.now_ns()
        key = self._key(endpoint)
        if not self._bus.has_peer(key):
            self.connect(endpoint)
        self._bus.send(key, data)
        t2 = Profiler.now_ns()
        self.sent_messages += 1
        self.sent_bytes += len(data)
        GlobalProfiler.add_event(EventType.COMMUNICATION, t0, t1, f"serialize {msg.command_type.name}", self.name)
        GlobalProfiler.add_event(EventType.COMMUNICATION, t1, t2, f"send {msg.command_type.name} {len(data)}B", self.name)

    def _receive_loop(self):
        while self._running:
            got = self._bus.recv(200)
            if got is None:
                continue
            sender, data = got
            t0 = Profiler.now_ns()
            try:
                msg = Message.deserialize(data, False, self.device, sender)
            except Exception as e:  # noqa: BLE001  (a corrupt frame must not kill the pump)
                msg = Message(CommandType.ERROR_REPORT, f"deserialise failed: {e!r}", sender)
            self.received_bytes += len(data)
            GlobalProfiler.add_event(EventType.COMMUNICATION, t0, Profiler.now_ns(),
                                     f"deserialize {msg.command_type.name} {len(data)}B", self.name)
            self.enqueue_input_message(msg)

    def stop(self):
        if not self._running:
            return
        self._running = False
        if self._pump is not None:
            self._pump.join(timeout=2)
        self._bus.stop()

    def stats(self) -> dict:
        d = super().stats()
        d.update({"bus": self._bus.stats()})
        return d


def __getattr__(name):
    # ``RoCECommunicator`` lives in rendezvous.py (which imports this module): resolve it lazily
    if name == "RoCECommunicator":
        from .rendezvous import RendezvousCommunicator
        return RendezvousCommunicator
    raise AttributeError(name)
