"""One-sided rendezvous transport: the shape of the reference's RoCE backend on a single node.

Reference: include/distributed/roce_communicator.hpp:39-391, roce_channel.hpp:17-418 and
include/device/ibv_allocator.hpp:31-266.  There every peer registers ONE host slab with the NIC, and a message
travels in three steps: the sender announces ``MSG_PREPARE(len, id)``, the receiver carves a block out of its
registered slab and answers ``MSG_READY_TO_WRITE(addr, rkey)``, the sender RDMA-writes the bytes straight into
that block and the write's immediate value (``id``) tells the receiver to deserialise; at most 16 messages per peer
are in flight.

There is no NIC between the GPUs of a B200 node, so the same protocol runs over memory every local process can
map: the registered slab is a POSIX shared-memory segment (``rkey`` = segment name), the one-sided write is a
memcpy into the peer's mapping, the immediate is a small control frame, and control frames ride the native TCP
bus.  Payload bytes therefore cross once through memory instead of through four loop-back sockets, and the
receiver deserialises in place.  Small messages (commands, acknowledgements) are sent inline, and a peer on a
different host -- recognised by its boot id -- gets the whole message inline over TCP, so the class is a drop-in
for ``TCPCommunicator`` anywhere.  ``RoCECommunicator`` is this class.

``SlabAllocator`` is the ``IbvAllocator`` analogue: best-fit over one registered range with coalescing frees.
"""
from __future__ import annotations

import mmap
import os
import struct
import threading
import uuid

from ..utils.profiler import EventType, GlobalProfiler, Profiler
from .communicator import TCPCommunicator
from .message import CommandType, Endpoint, Message

K_INLINE, K_PREPARE, K_READY, K_DONE, K_CANCEL = 0, 1, 2, 3, 4
_HDR = struct.Struct("<BQQ")          # kind, msg id, length / offset
_NO_BLOCK = (1 << 64) - 1             # offset value of a READY that grants nothing (message larger than the slab)


def _host_token() -> str:
    try:
        with open("/proc/sys/kernel/random/boot_id") as f:
            return f.read().strip()
    except OSError:
        import socket
        return socket.gethostname()


class _Segment:
    """A POSIX shared-memory segment mapped through /dev/shm directly: the creator unlinks it, peers only map it
    (``multiprocessing.shared_memory`` would make every attaching process's resource tracker try to unlink it too)."""

    def __init__(self, name: str, size: int = 0, create: bool = False):
        self.name, self.created = name, create
        path = os.path.join("/dev/shm", name)
        fd = os.open(path, (os.O_CREAT | os.O_EXCL | os.O_RDWR) if create else os.O_RDWR, 0o600)
        try:
            if create:
                os.ftruncate(fd, size)
            self.size = os.fstat(fd).st_size
            self._map = mmap.mmap(fd, self.size)
        finally:
            os.close(fd)
        self.buf = memoryview(self._map)

    def close(self):
        try:
            self.buf.release()
            self._map.close()
        except (BufferError, ValueError):
            pass

    def unlink(self):
        if self.created:
            try:
                os.unlink(os.path.join("/dev/shm", self.name))
            except FileNotFoundError:
                pass


class SlabAllocator:
    """Best-fit allocator over ``[0, size)`` with ``align``-byte blocks and coalescing frees (ibv_allocator.hpp:
    one slab registered once, blocks handed out as (offset, length))."""

    def __init__(self, size: int, align: int = 256):
        self.size, self.align = int(size), int(align)
        self._free = [(0, self.size)]               # sorted by offset, never adjacent
        self._live: dict[int, int] = {}
        self._lock = threading.Lock()
        self.peak = 0

    def allocate(self, nbytes: int) -> int | None:
        n = max(self.align, (int(nbytes) + self.align - 1) // self.align * self.align)
        with self._lock:
            best = None
            for i, (off, ln) in enumerate(self._free):
                if ln >= n and (best is None or ln < self._free[best][1]):
                    best = i
            if best is None:
                return None
            off, ln = self._free[best]
            if ln == n:
                self._free.pop(best)
            else:
                self._free[best] = (off + n, ln - n)
            self._live[off] = n
            self.peak = max(self.peak, self.allocated())
            return off

    def free(self, off: int):
        with self._lock:
            n = self._live.pop(off, None)
            if n is None:
                return
            lo, hi = off, off + n
            out = []
            for o, ln in self._free:
                if o + ln == lo:
                    lo = o
                elif o == hi:
                    hi = o + ln
                else:
                    out.append((o, ln))
            out.append((lo, hi - lo))
            out.sort()
            self._free = out

    def allocated(self) -> int:
        return sum(self._live.values())

    def largest_free(self) -> int:
        with self._lock:
            return max((ln for _, ln in self._free), default=0)


class RendezvousCommunicator(TCPCommunicator):
    """Drop-in ``TCPCommunicator`` whose large messages go through the PREPARE / READY_TO_WRITE / write+immediate
    rendezvous into the receiver's registered shared-memory slab."""

    MAX_IN_FLIGHT = 16                 # roce_channel.hpp: at most 16 rendezvous per peer

    def __init__(self, host: str = "127.0.0.1", port: int = 0, num_io_threads: int = 4, skts_per_endpoint: int = 4, device=None,
                 name: str = "rdv", slab_bytes: int = 256 << 20, inline_bytes: int = 16 << 10):
        super().__init__(host, port, num_io_threads, skts_per_endpoint, device, name)
        self.slab_bytes, self.inline_bytes = int(slab_bytes), int(inline_bytes)
        self._shm: _Segment | None = None
        self._slab: SlabAllocator | None = None
        self._peers_shm: dict[str, _Segment] = {}
        self._pending: dict[int, tuple[bytes, str, str]] = {}       # id -> (data, peer key, command name)
        self._waiting: list[tuple[str, int, int]] = []              # PREPAREs that did not fit yet
        self._sem: dict[str, threading.Semaphore] = {}
        self._next_id = 1
        self._lock = threading.Lock()
        self._token = _host_token()
        self.rendezvous_sent = self.rendezvous_received = self.inline_sent = self.fallback_inline = 0

    # ------------------------------------------------------------------ lifecycle
    def start(self):
        if self._running:
            return self
        self._shm = _Segment(f"tnn_{uuid.uuid4().hex[:16]}", self.slab_bytes, create=True)
        self._slab = SlabAllocator(self.slab_bytes)
        return super().start()

    start_server = start

    def endpoint(self) -> Endpoint:
        e = super().endpoint()
        return Endpoint(Endpoint.ROCE, host=e.params["host"], port=e.params["port"], device_name="shm", gid_index=-1)

    def stop(self):
        was = self._running
        super().stop()
        if not was:
            return
        for s in self._peers_shm.values():
            s.close()
        self._peers_shm.clear()
        if self._shm is not None:
            self._shm.unlink()
            self._shm.close()
            self._shm = None

    # ------------------------------------------------------------------ send side
    def _frame(self, key: str, kind: int, a: int = 0, b: int = 0, tail: bytes = b""):
        if not self._bus.has_peer(key):
            host, port = key.rsplit(":", 1)
            self._bus.connect(host, int(port), 10000)
        self._bus.send(key, _HDR.pack(kind, a, b) + tail)

    def _send_remote(self, msg: Message, endpoint: Endpoint):
        t0 = Profiler.now_ns()
        data = msg.serialize()
        t1 = Profiler.now_ns()
        key = self._key(endpoint)
        self.sent_messages += 1
        self.sent_bytes += len(data)
        GlobalProfiler.add_event(EventType.COMMUNICATION, t0, t1, f"serialize {msg.command_type.name}", self.name)
        if len(data) <= self.inline_bytes:
            self._frame(key, K_INLINE, 0, len(data), data)
            self.inline_sent += 1
            return
        sem = self._sem.setdefault(key, threading.Semaphore(self.MAX_IN_FLIGHT))
        sem.acquire()
        with self._lock:
            mid = self._next_id
            self._next_id += 1
            self._pending[mid] = (data, key, msg.command_type.name)
        self._frame(key, K_PREPARE, mid, len(data), self._token.encode())

    def _attach(self, name: str) -> _Segment:
        s = self._peers_shm.get(name)
        if s is None:
            s = self._peers_shm[name] = _Segment(name)
        return s

    def _on_ready(self, sender: str, mid: int, off: int, tail: bytes):
        with self._lock:
            item = self._pending.pop(mid, None)
        if item is None:
            return
        data, key, cname = item
        token, _, shm_name = tail.decode().partition("|")
        t0 = Profiler.now_ns()
        try:
            if token != self._token:
                raise OSError("peer is on another host")
            peer = self._attach(shm_name)
            peer.buf[off:off + len(data)] = data                        # the one-sided write
            self._frame(key, K_DONE, mid, off, struct.pack("<Q", len(data)))   # ... and its immediate
            self.rendezvous_sent += 1
        except (OSError, ValueError, FileNotFoundError):
            self._frame(key, K_CANCEL, mid, off)
            self._frame(key, K_INLINE, 0, len(data), data)
            self.fallback_inline += 1
        finally:
            self._sem[key].release()
        GlobalProfiler.add_event(EventType.COMMUNICATION, t0, Profiler.now_ns(), f"write {cname} {len(data)}B", self.name)

    # ------------------------------------------------------------------ receive side
    def _grant(self, sender: str, mid: int, length: int) -> bool:
        off = self._slab.allocate(length) if length <= self.slab_bytes else None
        if off is None:
            if length > self.slab_bytes:                # can never fit: make the sender fall back to inline
                self._frame(sender, K_READY, mid, _NO_BLOCK, b"-|")
                return True
            return False
        self._frame(sender, K_READY, mid, off, f"{self._token}|{self._shm.name}".encode())
        return True

    def _retry_waiting(self):
        still = []
        for sender, mid, length in self._waiting:
            if not self._grant(sender, mid, length):
                still.append((sender, mid, length))
        self._waiting = still

    def _deliver(self, data, sender: str):
        t0 = Profiler.now_ns()
        try:
            msg = Message.deserialize(data, False, self.device, sender)
        except Exception as e:  # noqa: BLE001  (a corrupt frame must not kill the pump)
            msg = Message(CommandType.ERROR_REPORT, f"deserialise failed: {e!r}", sender)
        self.received_bytes += len(data)
        GlobalProfiler.add_event(EventType.COMMUNICATION, t0, Profiler.now_ns(),
                                 f"deserialize {msg.command_type.name} {len(data)}B", self.name)
        self.enqueue_input_message(msg)

    def _receive_loop(self):
        while self._running:
            got = self._bus.recv(200)
            if got is None:
                continue
            sender, frame = got
            if len(frame) < _HDR.size:
                continue
            kind, a, b = _HDR.unpack_from(frame, 0)
            tail = frame[_HDR.size:]
            if kind == K_INLINE:
                self._deliver(tail, sender)
            elif kind == K_PREPARE:
                if not self._grant(sender, a, b):
                    self._waiting.append((sender, a, b))
            elif kind == K_READY:
                # the copy runs on a helper thread so that a long write does not stall incoming control frames
                threading.Thread(target=self._on_ready, args=(sender, a, b, tail), daemon=True).start()
            elif kind == K_DONE:
                (length,) = struct.unpack("<Q", tail[:8])
                self._deliver(bytes(self._shm.buf[b:b + length]), sender)
                self._slab.free(b)
                self.rendezvous_received += 1
                self._retry_waiting()
            elif kind == K_CANCEL:
                if b != _NO_BLOCK:
                    self._slab.free(b)
                    self._retry_waiting()

    def stats(self) -> dict:
        d = super().stats()
        d.update(rendezvous_sent=self.rendezvous_sent, rendezvous_received=self.rendezvous_received, inline_sent=self.inline_sent,
                 fallback_inline=self.fallback_inline, slab_bytes=self.slab_bytes,
                 slab_peak=self._slab.peak if self._slab is not None else 0)
        return d
