#!/usr/bin/env python
"""Pipeline-stage worker process on the TCP bus (reference: examples/tcp_worker.cpp:39-141).

  python examples/tcp_worker.py <port> [--gpu] [--gpu-index N] [--host 127.0.0.1] [--io-threads 4]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _pipeline_cli import worker_main  # noqa: E402

if __name__ == "__main__":
    worker_main("tcp")
