"""Coordinator / worker distributed runtime (reference: include/distributed/)."""
from .communicator import Communicator, InProcessCommunicator, TCPCommunicator  # noqa: F401
from .rendezvous import RendezvousCommunicator, SlabAllocator  # noqa: F401
RoCECommunicator = RendezvousCommunicator
from .coordinator import (Coordinator, CoordinatorConfig, DistributedError, NetworkCoordinator,  # noqa: F401
                          RoCECoordinator)
from .message import CommandType, Endpoint, Job, Message, ParallelMode, StageConfig  # noqa: F401
from .train import split_microbatches, train_model, train_semi_async_epoch, validate_semi_async_epoch  # noqa: F401
from .worker import InProcessWorker, RoCEWorker, TCPWorker, Worker  # noqa: F401
