"""Device layer: Device / DeviceManager / Flow / allocators.

Reference: include/device/{device,device_manager,flow,task,dptr,iallocator,pool_allocator,
del_allocator_v2}.hpp.  The B200 build keeps the API surface (``DeviceManager.getInstance()
.getDevice("GPU:0")``, flows keyed by handle, Pool / DEL allocators with ``flip/set_side/
reserve/clear``) but maps it onto CUDA streams + events and slab tensors:

* ``Flow`` = a CUDA stream (compute / comm / copy) with event-based ordering; kernels of this
  package always launch on the *current* torch stream, so ``with flow:`` scopes them.
* ``DELAllocator`` = the native double-ended offset allocator (csrc/runtime/allocator.hpp)
  backed by uint8 slab tensors; ``allocate`` returns a typed view, ``flip`` ping-pongs sides so
  consecutive layer outputs never alias.
* ``PoolAllocator`` = size-bucketed block cache with the reference's "reuse if <= 4x" rule.
"""
from __future__ import annotations

import enum
import math
import threading

import torch

from ..ops import native


class DeviceType(enum.Enum):
    CPU = "CPU"
    GPU = "GPU"
    NULL_DEVICE = "NULL"


class Flow:
    """Stream abstraction (flow.hpp:18-72).  Handle 0 is the default compute flow."""

    def __init__(self, device: "Device", handle: int = 0):
        self.device, self.handle = device, handle
        self.stream = None
        if device.type == DeviceType.GPU:
            self.stream = torch.cuda.current_stream(device.torch) if handle == 0 else torch.cuda.Stream(device.torch)

    def synchronize(self):
        if self.stream is not None:
            self.stream.synchronize()

    def record_event(self):
        if self.stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def wait_event(self, ev):
        if self.stream is not None and ev is not None:
            self.stream.wait_event(ev)

    def wait_flow(self, other: "Flow"):
        self.wait_event(other.record_event())

    def __enter__(self):
        if self.stream is not None:
            self._ctx = torch.cuda.stream(self.stream)
            self._ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.stream is not None:
            self._ctx.__exit__(*a)


class Task:
    """Handle on a unit of asynchronous work (reference: include/device/task.hpp -- ``ready()`` / ``sync()`` returning
    an error status; ``create_cpu_task`` / ``create_cuda_task``).  A CPU task runs on the native thread pool's Python
    side (a ``concurrent.futures`` future); a GPU task is whatever ``fn`` enqueued on its flow plus the event
    recorded behind it, so ``sync()`` waits for exactly that work and not for the whole device."""

    def __init__(self, future=None, event=None, flow: Flow | None = None, result=None, error: BaseException | None = None):
        self._future, self._event, self._flow = future, event, flow
        self._result, self._error = result, error

    def ready(self) -> bool:
        if self._error is not None:
            return True
        if self._future is not None:
            return self._future.done()
        if self._event is not None:
            return bool(self._event.query())
        return True

    def sync(self) -> BaseException | None:
        """Block until the work is done; returns None on success or the exception that the work raised."""
        if self._future is not None:
            try:
                self._result = self._future.result()
            except BaseException as e:  # noqa: BLE001
                self._error = e
            self._future = None
        if self._event is not None:
            self._event.synchronize()
        return self._error

    def result(self):
        err = self.sync()
        if err is not None:
            raise err
        return self._result


_TASK_POOL = None


def create_cpu_task(fn, *args, **kw) -> Task:
    """Run ``fn(*args)`` on a host worker thread (task.hpp create_cpu_task)."""
    global _TASK_POOL
    if _TASK_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        import os as _os
        _TASK_POOL = ThreadPoolExecutor(max_workers=max(2, min(8, _os.cpu_count() or 2)), thread_name_prefix="tnn-task")
    return Task(future=_TASK_POOL.submit(fn, *args, **kw))


def create_cuda_task(flow: Flow, fn, *args, **kw) -> Task:
    """Enqueue ``fn(*args)`` on ``flow`` (its kernels go to the flow's stream) and return a task bound to an event
    recorded right behind it (task.hpp create_cuda_task).  On a CPU flow the call simply runs."""
    try:
        with flow:
            res = fn(*args, **kw)
        return Task(event=flow.record_event(), flow=flow, result=res)
    except BaseException as e:  # noqa: BLE001
        return Task(error=e)


class Device:
    def __init__(self, type: DeviceType, index: int = 0):
        self.type, self.index = type, index
        self.torch = torch.device("cuda", index) if type == DeviceType.GPU else torch.device("cpu")
        self._flows: dict[int, Flow] = {}

    def getFlow(self, handle: int = 0) -> Flow:
        if handle not in self._flows:
            self._flows[handle] = Flow(self, handle)
        return self._flows[handle]

    get_flow = getFlow

    def name(self) -> str:
        return f"{self.type.value}:{self.index}"

    def total_memory(self) -> int:
        if self.type == DeviceType.GPU:
            return torch.cuda.get_device_properties(self.index).total_memory
        from ..utils.hardware_info import HardwareInfo
        return HardwareInfo().total_ram_kb * 1024

    def available_memory(self) -> int:
        """Free bytes right now (device.hpp getAvailableMemory)."""
        if self.type == DeviceType.GPU:
            return torch.cuda.mem_get_info(self.index)[0]
        from ..utils.hardware_info import HardwareInfo
        return HardwareInfo().available_ram_kb * 1024

    getTotalMemory, getAvailableMemory, getName = total_memory, available_memory, name

    def allocate(self, nbytes: int, pinned: bool = False) -> torch.Tensor:
        if self.type == DeviceType.GPU:
            return torch.empty(nbytes, dtype=torch.uint8, device=self.torch)
        return torch.empty(nbytes, dtype=torch.uint8, pin_memory=pinned and torch.cuda.is_available())

    def __repr__(self):
        return f"Device({self.name()})"


class DeviceManager:
    """Singleton that discovers CPU:0 and all GPUs (device_manager.cpp:22-62,128-152)."""
    _inst = None
    _lock = threading.Lock()

    def __init__(self):
        self.devices: dict[str, Device] = {"CPU:0": Device(DeviceType.CPU, 0)}
        if torch.cuda.is_available():
            for i in range(torch.cuda.device_count()):
                self.devices[f"GPU:{i}"] = Device(DeviceType.GPU, i)

    @classmethod
    def getInstance(cls) -> "DeviceManager":
        with cls._lock:
            if cls._inst is None:
                cls._inst = DeviceManager()
        return cls._inst

    get_instance = getInstance

    def getDevice(self, spec) -> Device:
        if isinstance(spec, DeviceType):
            spec = f"{spec.value}:0"
        spec = spec.upper()
        if ":" not in spec:
            spec += ":0"
        if spec.startswith("CUDA"):
            spec = "GPU" + spec[4:]
        if spec not in self.devices:
            raise KeyError(f"no device {spec}; available: {sorted(self.devices)}")
        return self.devices[spec]

    get_device = getDevice

    def getHost(self) -> Device:
        return self.devices["CPU:0"]

    def getGPU(self, i: int = 0) -> Device:
        return self.getDevice(f"GPU:{i}")

    def available(self) -> list[str]:
        return sorted(self.devices)

    getAvailableDeviceIDs = available

    def hasDevice(self, spec) -> bool:
        try:
            self.getDevice(spec)
            return True
        except KeyError:
            return False

    has_device = hasDevice


def getHost() -> Device:
    return DeviceManager.getInstance().getHost()


def getGPU(i: int = 0) -> Device:
    return DeviceManager.getInstance().getGPU(i)


_DT_SIZE = {torch.float32: 4, torch.bfloat16: 2, torch.float16: 2, torch.float64: 8, torch.int32: 4, torch.int64: 8,
            torch.uint8: 1, torch.bool: 1, torch.int8: 1, torch.int16: 2}


class DELAllocator:
    """Double-ended slab allocator (del_allocator_v2.hpp:73-336) over native offset bookkeeping."""

    def __init__(self, device: Device | torch.device | str = "cpu", slab_bytes: int = 64 << 20, align: int = 256):
        if not isinstance(device, Device):
            d = torch.device(device)
            device = Device(DeviceType.GPU, d.index or 0) if d.type == "cuda" else Device(DeviceType.CPU, 0)
        self.device = device
        self._core = native.require().DELAllocator(slab_bytes, align)
        self._align = int(align)
        self._slabs: list[torch.Tensor] = []
        self._live: dict[int, tuple[int, int]] = {}

    @staticmethod
    def create(device="cpu", flow=None, **kw) -> "DELAllocator":
        return DELAllocator(device, **kw)

    def _back(self):
        while len(self._slabs) < self._core.num_slabs():
            n = self._core.slab_size(len(self._slabs))
            raw = self.device.allocate(n + self._align)           # base aligned so offsets stay aligned
            skew = (-raw.data_ptr()) % self._align
            self._slabs.append(raw[skew:skew + n])

    def allocate(self, shape, dtype=torch.float32) -> torch.Tensor:
        n = int(math.prod(shape)) if len(shape) else 1
        nbytes = max(1, n * _DT_SIZE[dtype])
        slab, off, size = self._core.allocate(nbytes)
        self._back()
        t = self._slabs[slab][off:off + n * _DT_SIZE[dtype]].view(dtype).view(*shape)
        self._live[t.data_ptr()] = (slab, off)
        return t

    def free(self, t: torch.Tensor):
        key = t.data_ptr()
        if key in self._live:
            slab, off = self._live.pop(key)
            self._core.free(slab, off)

    def flip(self):
        self._core.flip()

    def set_side(self, right: bool):
        self._core.set_side(bool(right))

    def side(self) -> bool:
        return self._core.side()

    def reserve(self, nbytes: int):
        self._core.reserve(int(nbytes))
        self._back()

    def clear(self):
        self._core.clear()
        self._live.clear()

    def total_allocated(self) -> int:
        return self._core.total_allocated()

    def peak_allocated(self) -> int:
        return self._core.peak_allocated()

    def total_reserved(self) -> int:
        return self._core.total_reserved()


class PoolAllocator:
    """Size-keyed free list; a cached block is reused when <= 4x the request (pool_allocator.hpp:104)."""
    _instances: dict = {}

    def __init__(self, device: Device | torch.device | str = "cpu"):
        if not isinstance(device, Device):
            d = torch.device(device)
            device = Device(DeviceType.GPU, d.index or 0) if d.type == "cuda" else Device(DeviceType.CPU, 0)
        self.device = device
        self._core = native.require().PoolAllocatorCore()
        self._blocks: dict[int, torch.Tensor] = {}
        self._owner: dict[int, int] = {}

    @classmethod
    def instance(cls, device="cpu", flow=None) -> "PoolAllocator":
        key = str(device.name() if isinstance(device, Device) else device)
        if key not in cls._instances:
            cls._instances[key] = PoolAllocator(device)
        return cls._instances[key]

    def allocate(self, shape, dtype=torch.float32) -> torch.Tensor:
        n = int(math.prod(shape)) if len(shape) else 1
        nbytes = max(256, (n * _DT_SIZE[dtype] + 255) // 256 * 256)
        bid = self._core.acquire(nbytes)
        if bid < 0:
            bid = self._core.add(nbytes)
            self._blocks[bid] = self.device.allocate(nbytes)
        t = self._blocks[bid][:n * _DT_SIZE[dtype]].view(dtype).view(*shape)
        self._owner[t.data_ptr()] = bid
        return t

    def free(self, t: torch.Tensor):
        bid = self._owner.pop(t.data_ptr(), None)
        if bid is not None:
            self._core.release(bid)

    def total_bytes(self) -> int:
        return self._core.total_bytes()

    def cached_bytes(self) -> int:
        return self._core.cached_bytes()


def HostAllocator():
    return PoolAllocator.instance("cpu")


def GPUAllocator(i: int = 0):
    return PoolAllocator.instance(f"cuda:{i}")
