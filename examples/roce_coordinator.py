#!/usr/bin/env python
"""The same coordinator over the rendezvous (RoCE-protocol) transport (reference: examples/roce_coordinator.cpp:20-167):
PREPARE / READY / WRITE+IMM over registered slabs, see tnn_b200/distributed/rendezvous.py.

  WORKER1_PORT=9001 python examples/roce_coordinator.py --synthetic --proportions 1,2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _pipeline_cli import coordinator_main  # noqa: E402

if __name__ == "__main__":
    coordinator_main("roce")
