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

GPU tensors take the DEVICE rendezvous (the GPUDirect-RDMA analogue): a communicator that owns a CUDA device registers
a second slab in device memory (``cudaMalloc`` + ``cudaIpcGetMemHandle``).  A job whose tensor lives on a GPU is
announced with ``G_PREPARE(nbytes)``; the receiver carves a block out of its device slab and answers
``G_READY(offset, ipc handle)``; the sender opens the handle once, copies the tensor device-to-device straight into the
block (NVLink / NVSwitch between the GPUs of one node, one hop, no host staging), and ``G_DONE`` carries the small
metadata record (command, micro-batch id, dtype, shape).  The receiver hands the worker a tensor that IS the block
(zero-copy view; the block returns to the slab when the tensor dies and the consumer stream has passed that point).
Peers on other hosts, CPU-only peers and messages that do not fit fall back to the host path above.
"""
from __future__ import annotations

import mmap
import os
import struct
import threading
import uuid

import torch

from ..core.dtype import DType, from_torch, to_torch
from ..ops import native
from ..utils.profiler import EventType, GlobalProfiler, Profiler
from .communicator import TCPCommunicator
from .message import CommandType, Endpoint, Job, Message

K_INLINE, K_PREPARE, K_READY, K_DONE, K_CANCEL = 0, 1, 2, 3, 4
K_GPREPARE, K_GREADY, K_GDONE = 5, 6, 7          # device rendezvous (CUDA-IPC slab)
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


class _DeviceBlock:
    """A block of the device slab exposed through ``__cuda_array_interface__`` (bytes): ``torch.as_tensor`` wraps it
    without a copy and keeps this object alive; when the last view dies the block goes back to the slab -- after the
    consumer's stream has passed the point of the release (an event recorded at release time)."""

    def __init__(self, owner: "RendezvousCommunicator", ptr: int, off: int, nbytes: int):
        self._owner, self._off = owner, off
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr + off, False), "version": 3,
                                         "strides": None}

    def __del__(self):
        try:
            self._owner._release_device_block(self._off)
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


_GMETA = struct.Struct("<HQIIQ")      # command, micro-batch id, dtype code, ndim, payload bytes (dims follow as u64)


class RendezvousCommunicator(TCPCommunicator):
    """Drop-in ``TCPCommunicator`` whose large messages go through the PREPARE / READY_TO_WRITE / write+immediate
    rendezvous into the receiver's registered shared-memory slab."""

    MAX_IN_FLIGHT = 16                 # roce_channel.hpp: at most 16 rendezvous per peer

    def __init__(self, host: str = "127.0.0.1", port: int = 0, num_io_threads: int = 4, skts_per_endpoint: int = 4, device=None,
                 name: str = "rdv", slab_bytes: int = 256 << 20, inline_bytes: int = 16 << 10, device_slab_bytes: int | None = None):
        super().__init__(host, port, num_io_threads, skts_per_endpoint, device, name)
        self.slab_bytes, self.inline_bytes = int(slab_bytes), int(inline_bytes)
        # device rendezvous (CUDA-IPC slab): on when this communicator owns a CUDA device, TNN_DEVICE_RENDEZVOUS=0 disables
        env = os.environ.get("TNN_DEVICE_SLAB_MB")
        self.device_slab_bytes = int(device_slab_bytes) if device_slab_bytes is not None else (int(env) << 20 if env else 256 << 20)
        self._dptr = 0
        self._dhandle = b""
        self._dslab: SlabAllocator | None = None
        self._dindex = -1
        self._dreleased: list[tuple[object, int]] = []            # (event, offset) blocks waiting for the consumer stream
        self._dopened: dict[bytes, int] = {}                        # peer ipc handle -> mapped base pointer
        self._pending_gpu: dict[int, tuple] = {}
        self.device_rendezvous_sent = self.device_rendezvous_received = self.device_fallback_host = 0
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
        self._start_device_slab()
        return super().start()

    def _start_device_slab(self):
        if self.device is None or os.environ.get("TNN_DEVICE_RENDEZVOUS", "1") == "0" or self.device_slab_bytes <= 0:
            return
        dev = torch.device(self.device)
        if dev.type != "cuda" or not torch.cuda.is_available():
            return
        self._dindex = dev.index if dev.index is not None else torch.cuda.current_device()
        torch.cuda.init()
        self._dptr, self._dhandle = native.require().cuda_ipc_alloc(self._dindex, self.device_slab_bytes)
        self._dslab = SlabAllocator(self.device_slab_bytes, align=512)

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
        if self._dptr:
            C = native.require()
            for base in self._dopened.values():
                if base:
                    C.cuda_ipc_close(self._dindex, base)
            self._dopened.clear()
            # blocks still referenced by live tensors keep their memory valid only until here: the slab is released
            C.cuda_ipc_free(self._dindex, self._dptr)
            self._dptr, self._dslab = 0, None
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
        if self._dslab is not None and isinstance(msg.payload, Job) and getattr(msg.payload.data, "is_cuda", False):
            t = msg.payload.data
            nbytes = t.numel() * t.element_size()
            if nbytes > self.inline_bytes and t.dtype != torch.bool:
                return self._send_device(msg, endpoint, nbytes)
        return self._send_host(msg, endpoint)

    # ------------------------------------------------------------------ device rendezvous, send side
    def _send_device(self, msg: Message, endpoint: Endpoint, nbytes: int):
        t = msg.payload.data.contiguous()
        torch.cuda.current_stream(t.device).synchronize()           # the producer of the tensor has finished
        key = self._key(endpoint)
        dims = list(t.shape)
        meta = _GMETA.pack(int(msg.command_type), int(msg.payload.mb_id), int(from_torch(t.dtype)), len(dims), nbytes) + \
            struct.pack(f"<{len(dims)}Q", *dims)
        sem = self._sem.setdefault(key, threading.Semaphore(self.MAX_IN_FLIGHT))
        sem.acquire()
        with self._lock:
            mid = self._next_id
            self._next_id += 1
            self._pending_gpu[mid] = (t, meta, key, msg, endpoint)
        self.sent_messages += 1
        self.sent_bytes += nbytes
        self._frame(key, K_GPREPARE, mid, nbytes, f"{self._token}|{os.getpid()}".encode())

    def _on_gready(self, sender: str, mid: int, off: int, tail: bytes):
        with self._lock:
            item = self._pending_gpu.pop(mid, None)
        if item is None:
            return
        t, meta, key, msg, endpoint = item
        t0 = Profiler.now_ns()
        ok = False
        try:
            if off != _NO_BLOCK:
                pid, _, rest = tail.partition(b"|")
                ptr, _, handle = rest.partition(b"|")
                if int(pid) == os.getpid():
                    base = int(ptr)                                # same process: the slab pointer is directly usable
                else:
                    base = self._dopened.get(handle)
                    if base is None:
                        base = self._dopened[handle] = native.require().cuda_ipc_open(t.device.index, bytes.fromhex(handle.decode()))
                native.require().cuda_copy_d2d(t.device.index, base + off, t.data_ptr(), t.numel() * t.element_size())   # the one-sided write
                self._frame(key, K_GDONE, mid, off, meta)                                                               # ... and its immediate
                self.device_rendezvous_sent += 1
                ok = True
        except Exception:  # noqa: BLE001  (cannot map the peer's slab: give the block back, use the host path)
            if off != _NO_BLOCK:
                self._frame(key, K_GDONE, mid, off, b"")            # empty metadata = cancelled
        finally:
            self._sem[key].release()
        if not ok:
            self.device_fallback_host += 1
            self.sent_messages -= 1
            self.sent_bytes -= t.numel() * t.element_size()
            self._send_host(msg, endpoint)
        GlobalProfiler.add_event(EventType.COMMUNICATION, t0, Profiler.now_ns(),
                                 f"device write {msg.command_type.name} {t.numel() * t.element_size()}B", self.name)

    # ------------------------------------------------------------------ device rendezvous, receive side
    def _release_device_block(self, off: int):
        """Called when the last tensor view of a block dies: free it once the consumer stream has passed this point."""
        ev = None
        try:
            if self._dslab is not None and torch.cuda.is_available():
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self._dindex))
        except Exception:  # noqa: BLE001
            ev = None
        with self._lock:
            self._dreleased.append((ev, off))

    def _reap_device_blocks(self):
        with self._lock:
            done = [(e, o) for e, o in self._dreleased if e is None or e.query()]
            self._dreleased = [(e, o) for e, o in self._dreleased if not (e is None or e.query())]
        if self._dslab is not None:
            for _, o in done:
                self._dslab.free(o)

    def _ggrant(self, sender: str, mid: int, nbytes: int, tail: bytes):
        token = tail.partition(b"|")[0].decode()
        off = None
        if self._dslab is not None and token == self._token and nbytes <= self.device_slab_bytes:
            self._reap_device_blocks()
            off = self._dslab.allocate(nbytes)
        if off is None:
            self._frame(sender, K_GREADY, mid, _NO_BLOCK, b"")     # no device slab / other host / full: sender uses the host path
            return
        self._frame(sender, K_GREADY, mid, off, f"{os.getpid()}|{self._dptr}|{self._dhandle.hex()}".encode())

    def _on_gdone(self, sender: str, off: int, meta: bytes):
        if self._dslab is None:
            return
        if len(meta) < _GMETA.size:                                 # cancelled by the sender
            self._dslab.free(off)
            return
        t0 = Profiler.now_ns()
        cmd, mb, dt, nd, nbytes = _GMETA.unpack_from(meta, 0)
        dims = struct.unpack_from(f"<{nd}Q", meta, _GMETA.size)
        block = _DeviceBlock(self, self._dptr, off, nbytes)
        raw = torch.as_tensor(block, device=torch.device("cuda", self._dindex))     # zero-copy: the tensor IS the slab block
        t = raw.view(to_torch(DType(dt))).reshape(dims)
        self.received_bytes += nbytes
        self.device_rendezvous_received += 1
        GlobalProfiler.add_event(EventType.COMMUNICATION, t0, Profiler.now_ns(), f"device deliver {CommandType(cmd).name} {nbytes}B", self.name)
        self.enqueue_input_message(Message(CommandType(cmd), Job(t, int(mb)), sender))

    # ------------------------------------------------------------------ host path
    def _send_host(self, msg: Message, endpoint: Endpoint):
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
            elif kind == K_GPREPARE:
                self._ggrant(sender, a, b, tail)
            elif kind == K_GREADY:
                threading.Thread(target=self._on_gready, args=(sender, a, b, tail), daemon=True).start()
            elif kind == K_GDONE:
                self._on_gdone(sender, b, tail)

    def stats(self) -> dict:
        d = super().stats()
        d.update(rendezvous_sent=self.rendezvous_sent, rendezvous_received=self.rendezvous_received, inline_sent=self.inline_sent,
                 fallback_inline=self.fallback_inline, slab_bytes=self.slab_bytes,
                 device_rendezvous_sent=self.device_rendezvous_sent, device_rendezvous_received=self.device_rendezvous_received,
                 device_fallback_host=self.device_fallback_host, device_slab_bytes=self.device_slab_bytes if self._dslab is not None else 0,
                 device_slab_peak=self._dslab.peak if self._dslab is not None else 0,
                 slab_peak=self._slab.peak if self._slab is not None else 0)
        return d
