#!/usr/bin/env python
"""Pipeline-parallel training driven by a coordinator over the TCP bus (reference: examples/tcp_coordinator.cpp:27-108).
Endpoints from the environment like the reference: COORDINATOR_HOST/PORT, LOCAL_WORKER_HOST/PORT/POSITION, WORKER1_HOST/PORT, ...

  WORKER1_PORT=9001 python examples/tcp_coordinator.py --synthetic --proportions 1,2
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _pipeline_cli import coordinator_main  # noqa: E402

if __name__ == "__main__":
    coordinator_main("tcp")
