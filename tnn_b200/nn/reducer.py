"""Graph reduction passes (reference: include/nn/reducer.hpp:9-41, reduction_strategy.hpp,
reductions_impl/seq_reduction.hpp:14-150).

``SeqReduction`` walks a ``GraphBuilder`` and merges every maximal chain of single-input / single-output
operations whose intermediate tensors have no other consumer into ONE ``Sequential`` op node.  Besides
removing executor hops, this is what lets the in-Sequential fusions of this framework see the chain
(BatchNorm + Dropout apply, shortcut add in the tail GEMM epilogue, side-stream weight gradients joined
per Sequential, ...), which the executor cannot do across separate graph nodes.
"""
from __future__ import annotations

from .blocks import Sequential
from .graph import Edge, GraphBuilder


class ReductionStrategy:
    def reduce(self, builder: GraphBuilder) -> None:
        raise NotImplementedError


class SeqReduction(ReductionStrategy):
    def reduce(self, builder: GraphBuilder) -> None:
        edges = builder.topo_sort()
        if len(edges) <= 1:
            return
        producer = {}                     # io uid -> edge that writes it
        n_consumers: dict[str, int] = {}  # io uid -> number of edges that read it
        for e in edges:
            for o in e.outputs:
                producer[o] = e
            for i in e.inputs:
                n_consumers[i] = n_consumers.get(i, 0) + 1
        consumer_of = {}                  # io uid -> the single edge reading it
        for e in edges:
            for i in e.inputs:
                if n_consumers[i] == 1:
                    consumer_of[i] = e

        def fusable(e: Edge) -> bool:
            return len(e.inputs) == 1 and len(e.outputs) == 1

        used: set[int] = set()
        chains: list[list[Edge]] = []
        for e in edges:
            if id(e) in used or not fusable(e):
                continue
            # only start at a chain head: the producer of my input is not itself a fusable single-consumer link
            p = producer.get(e.inputs[0])
            if p is not None and fusable(p) and n_consumers.get(e.inputs[0], 0) == 1 and id(p) not in used:
                continue
            chain = [e]
            used.add(id(e))
            cur = e
            while True:
                out = cur.outputs[0]
                nxt = consumer_of.get(out)
                if nxt is None or n_consumers.get(out, 0) != 1 or not fusable(nxt) or id(nxt) in used:
                    break
                chain.append(nxt)
                used.add(id(nxt))
                cur = nxt
            if len(chain) > 1:
                chains.append(chain)
            else:
                used.discard(id(e))
        if not chains:
            return
        drop = {id(e) for c in chains for e in c}
        new_edges = [e for e in builder.edges if id(e) not in drop]
        for k, chain in enumerate(chains):
            layers = [builder.op_nodes.pop(e.op).layer for e in chain]
            inner = {e.outputs[0] for e in chain[:-1]}
            for uid in inner:                                   # intermediate tensors disappear from the graph
                builder.io_nodes.pop(uid, None)
            seq = Sequential(layers, f"seq_reduced_{k}")
            node = builder.add_layer(seq, uid=f"seq_reduced_{k}")
            new_edges.append(Edge(chain[0].inputs, chain[-1].outputs, node.uid))
        builder.edges = new_edges


class Reducer:
    def __init__(self, strategies: list[ReductionStrategy]):
        self.strategies = list(strategies)

    def reduce(self, builder: GraphBuilder) -> GraphBuilder:
        for s in self.strategies:
            s.reduce(builder)
        return builder


class ReducerBuilder:
    def __init__(self):
        self._s: list[ReductionStrategy] = []

    def add_strategy(self, s: ReductionStrategy) -> "ReducerBuilder":
        self._s.append(s)
        return self

    def seq_reduction(self) -> "ReducerBuilder":
        return self.add_strategy(SeqReduction())

    def build(self) -> Reducer:
        return Reducer(self._s)
