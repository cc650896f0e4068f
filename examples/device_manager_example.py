#!/usr/bin/env python
"""Lists the devices the manager discovered, their memory, and does an allocate / touch / free
round on the host and on the first GPU (reference: examples/device_manager_example.cpp:10-91)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tnn_b200.core.device import DeviceManager, DeviceType  # noqa: E402


def main():
    print("=== Device Manager Example ===")
    manager = DeviceManager.getInstance()
    ids = manager.getAvailableDeviceIDs()
    print(f"Found {len(ids)} device(s):")
    for dev_id in ids:
        dev = manager.getDevice(dev_id)
        print(f"  Device {dev_id}: {dev.getName()} (Type: {dev.type.name})")
        print(f"    Memory - Total: {dev.getTotalMemory() >> 20} MB, Available: {dev.getAvailableMemory() >> 20} MB")
    if manager.hasDevice("CPU:0"):
        print("Testing allocation on CPU device...")
        n = 1 << 20
        buf = manager.getHost().allocate(n)
        buf.fill_(0xAA)
        ok = int(buf[0]) == 0xAA and int(buf[-1]) == 0xAA
        print(f"  Successfully allocated {n} bytes\n  Memory access test {'passed' if ok else 'failed'}")
        del buf
        print("  Successfully deallocated memory")
    gpus = [i for i in ids if manager.getDevice(i).type == DeviceType.GPU]
    if not gpus:
        print("No GPU devices available for testing")
        return 0
    dev = manager.getDevice(gpus[0])
    print(f"Testing allocation on GPU device {gpus[0]}...")
    n = 1 << 30
    buf = dev.allocate(n)
    buf.fill_(0x55)
    dev.getFlow().synchronize()
    print(f"  Successfully allocated {n} bytes on GPU (first byte {int(buf[0]):#x})")
    del buf
    print("  Successfully deallocated GPU memory")
    return 0


if __name__ == "__main__":
    sys.exit(main())
