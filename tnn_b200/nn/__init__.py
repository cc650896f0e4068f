from .layer import Layer, LayerFactory, ParamDescriptor, ParameterizedLayer, SISOLayer, StatelessLayer, register  # noqa: F401
from .layers import *  # noqa: F401,F403
from .blocks import Sequential, MSequential, ResidualBlock, AttentionBlock, FlashAttentionBlock, SDPALayer  # noqa: F401
from .builder import LayerBuilder  # noqa: F401
from .graph import Graph, GraphBuilder, GraphContext, GraphExecutor, graph_from_sequential  # noqa: F401
from .loss import LossFactory, CrossEntropyLoss, compute_class_corrects  # noqa: F401
from .optimizers import OptimizerFactory, SGD, Adam  # noqa: F401
from .schedulers import SchedulerFactory  # noqa: F401
from .reducer import Reducer, ReducerBuilder, ReductionStrategy, SeqReduction  # noqa: F401,E402
