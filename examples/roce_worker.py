#!/usr/bin/env python
"""Pipeline-stage worker process (reference: examples/tcp_worker.cpp:39-141, roce_worker.cpp).

  python examples/roce_worker.py <port> [--gpu] [--gpu-index N] [--host 127.0.0.1] [--io-threads 4]

Waits for a coordinator to send CONFIG_TRANSFER, then serves FORWARD/BACKWARD/UPDATE commands
until SHUTDOWN.  ``--device/--gid-index`` of the reference RoCE worker are accepted and ignored: the
transport is the shared-memory rendezvous of tnn_b200/distributed/rendezvous.py (no NIC to pick).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tnn_b200.distributed import RoCEWorker  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("port", type=int)
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--gpu-index", type=int, default=0)
    ap.add_argument("--io-threads", type=int, default=4)
    ap.add_argument("--num-threads", type=int, default=8)
    ap.add_argument("--device", default="")
    ap.add_argument("--gid-index", type=int, default=-1)
    a = ap.parse_args()
    import torch
    torch.set_num_threads(a.num_threads)
    w = RoCEWorker(a.host, a.port, a.gpu, a.gpu_index, a.io_threads)
    print(f"[worker] listening on {a.host}:{a.port} ({'GPU ' + str(a.gpu_index) if a.gpu else 'CPU'})", flush=True)
    w.start()
    print("[worker] shutdown", flush=True)
