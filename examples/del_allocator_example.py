#!/usr/bin/env python
"""Double-ended slab allocator walk-through (reference: examples/test_del.cpp:8-35): reserve one
slab, allocate from the left end, flip and allocate from the right end, release blocks and watch a
request that fits the coalesced hole reuse it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tnn_b200.core.device import DELAllocator  # noqa: E402


def main():
    device = "cuda" if torch.cuda.is_available() else "cpu"
    a = DELAllocator.create(device, slab_bytes=1 << 20)
    a.reserve(1 << 20)
    base = None

    def where(name, t):
        nonlocal base
        base = t.data_ptr() if base is None else min(base, t.data_ptr())
        print(f"  {name}: offset {t.data_ptr() - base:>8} .. {t.data_ptr() - base + t.numel() * t.element_size():>8}")

    p1 = a.allocate((10 * 1024,), torch.uint8)
    p2 = a.allocate((20 * 1024,), torch.uint8)
    a.flip()                                   # subsequent blocks come from the right end of the slab
    p3 = a.allocate((15 * 1024,), torch.uint8)
    p4 = a.allocate((25 * 1024,), torch.uint8)
    print("left side grows up, right side grows down:")
    for n, t in (("ptr1", p1), ("ptr2", p2), ("ptr3", p3), ("ptr4", p4)):
        where(n, t)
    assert p2.data_ptr() > p1.data_ptr() and p3.data_ptr() > p4.data_ptr() > p2.data_ptr()
    print(f"allocated {a.total_allocated()} B of {a.total_reserved()} B reserved")
    for t in (p2, p3, p4):
        a.free(t)
    print(f"after releasing ptr2..ptr4: {a.total_allocated()} B live")
    a.flip()
    px = a.allocate((30 * 1024,), torch.uint8)
    where("ptrX", px)
    assert px.data_ptr() == p2.data_ptr(), "the released left-side block is reused"
    print(f"peak {a.peak_allocated()} B, slabs {a._core.num_slabs()}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
