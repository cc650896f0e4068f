#!/usr/bin/env python
"""Pipeline-stage worker process on the rendezvous (RoCE-protocol) transport (reference: examples/roce_worker.cpp).

  python examples/roce_worker.py <port> [--gpu] [--gpu-index N] [--device NIC] [--gid-index G]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _pipeline_cli import worker_main  # noqa: E402

if __name__ == "__main__":
    worker_main("roce")
