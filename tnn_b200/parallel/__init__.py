from .partitioner import (FTDPartitioner, InputPartition, NaiveDataPartitioner, NaivePartitionerConfig,  # noqa: F401
                          NaivePipelinePartitioner, Partitioner, SeqPartition, WeightedPipelinePartitioner)
