"""Graph IR: IONode / OpNode / Edge / GraphBuilder / GraphContext / Graph / GraphExecutor.

Reference: include/nn/{node,io_node,op_node,edge,graph_builder,graph_context,graph,
graph_executor}.hpp.  GraphContext packs ALL trainable parameters into one fp32 slab, all
gradients into another (256-byte aligned sub-spans; ``zero_grads`` is a single fill) and -- on a
GPU -- all bf16 compute copies into a third; this contiguous layout is what the flat / fused
data-parallel optimizer kernels operate on.  Non-trainable buffers (BatchNorm running stats)
live in a fourth slab and are serialised with the parameters.
"""
from __future__ import annotations

import io
import json
import struct
from collections import defaultdict, deque

import torch

from ..core import tensor_io
from ..core.config import TConfig
from .layer import Layer, LayerFactory

ALIGN_ELEMS = 64  # 256 bytes of fp32


class IONode:
    def __init__(self, uid: str):
        self.uid = uid

    def get_config(self) -> TConfig:
        """NodeConfig of the reference (src/nn/io_node.cpp:5-11): TConfig type "io_node", parameter uid."""
        return TConfig("", "io_node", uid=self.uid)


class OpNode:
    def __init__(self, uid: str, layer: Layer):
        self.uid = uid
        self.layer = layer

    def get_config(self) -> TConfig:
        """src/nn/op_node.cpp:6-15: TConfig type "op_node", parameters uid / name / layer_config (json)."""
        c = TConfig("", "op_node")
        return c.set("uid", self.uid).set("name", self.layer.name).set("layer_config", self.layer.get_config())


class Edge:
    def __init__(self, inputs: list[str], outputs: list[str], op: str):
        self.inputs, self.outputs, self.op = list(inputs), list(outputs), op

    def get_config(self):
        """include/nn/graph.hpp:93-113: plain JSON {producers: [io uid], consumers: [io uid], op_node: uid}."""
        return {"producers": self.inputs, "consumers": self.outputs, "op_node": self.op}


class GraphBuilder:
    def __init__(self, name: str = "graph"):
        self.name = name
        self.io_nodes: dict[str, IONode] = {}
        self.op_nodes: dict[str, OpNode] = {}
        self.edges: list[Edge] = []

    def io(self, uid: str) -> IONode:
        if uid not in self.io_nodes:
            self.io_nodes[uid] = IONode(uid)
        return self.io_nodes[uid]

    def add_layer(self, layer: Layer, uid: str | None = None) -> OpNode:
        uid = uid or f"op{len(self.op_nodes)}_{layer.name}"
        if uid in self.op_nodes:
            raise ValueError(f"duplicate op node {uid}")
        node = OpNode(uid, layer)
        self.op_nodes[uid] = node
        return node

    def add_edge(self, inputs, outputs, op) -> "GraphBuilder":
        ins = [i.uid if isinstance(i, IONode) else i for i in inputs]
        outs = [o.uid if isinstance(o, IONode) else o for o in outputs]
        opu = op.uid if isinstance(op, OpNode) else op
        for u in ins + outs:
            self.io(u)
        if opu not in self.op_nodes:
            raise KeyError(f"unknown op node {opu}")
        self.edges.append(Edge(ins, outs, opu))
        return self

    def topo_sort(self) -> list[Edge]:
        """Kahn's algorithm over edges; raises on cycles and on io nodes with two producers
        (graph_builder.hpp:28-112).  The ordering itself is computed by the native planner
        (csrc/runtime/planner.hpp::topo_order); ``_topo_sort_py`` is the same algorithm kept as the test oracle."""
        from ..ops import native
        C = native.load()
        if C is None or not hasattr(C, "plan_topo_order"):
            return self._topo_sort_py()
        ids: dict[str, int] = {}

        def nid(u):
            return ids.setdefault(u, len(ids))

        ins = [[nid(u) for u in e.inputs] for e in self.edges]
        outs = [[nid(u) for u in e.outputs] for e in self.edges]
        try:
            order = C.plan_topo_order(ins, outs)
        except (ValueError, RuntimeError) as err:
            msg = str(err)
            for u, i in ids.items():                       # report the io node by its name, not its dense id
                msg = msg.replace(f"io node {i} ", f"io node {u} ")
            raise ValueError(msg) from None
        return [self.edges[i] for i in order]

    def _topo_sort_py(self) -> list[Edge]:
        producers = {}
        for e in self.edges:
            for o in e.outputs:
                if o in producers:
                    raise ValueError(f"io node {o} has two producers")
                producers[o] = e
        indeg = {id(e): 0 for e in self.edges}
        consumers = defaultdict(list)
        for e in self.edges:
            for i in e.inputs:
                if i in producers:
                    indeg[id(e)] += 1
                    consumers[id(producers[i])].append(e)
        q = deque(e for e in self.edges if indeg[id(e)] == 0)
        order = []
        while q:
            e = q.popleft()
            order.append(e)
            for c in consumers[id(e)]:
                indeg[id(c)] -= 1
                if indeg[id(c)] == 0:
                    q.append(c)
        if len(order) != len(self.edges):
            raise ValueError("graph has a cycle")
        return order

    def compile(self, device="cpu", seed: int | None = None, alloc_fn=None) -> "Graph":
        return Graph(self, torch.device(device), seed, alloc_fn)


def _plan_slab(sizes: list[int]) -> list[int]:
    """Element offsets of ``sizes`` packed with ALIGN_ELEMS alignment; the last entry is the slab length."""
    from ..ops import native
    C = native.load()
    if C is not None and hasattr(C, "plan_slab"):
        return list(C.plan_slab([int(n) for n in sizes], ALIGN_ELEMS))
    out, cur = [], 0
    for n in sizes:
        out.append(cur)
        cur += (n + ALIGN_ELEMS - 1) // ALIGN_ELEMS * ALIGN_ELEMS
    return out + [cur]


class GraphContext:
    def __init__(self, layers: list[Layer], device: torch.device, alloc_fn=None):
        self.device = device
        self.layers = layers
        alloc = alloc_fn or (lambda n, dt: torch.zeros(n, dtype=dt, device=device))
        train, frozen = [], []
        for layer in layers:
            for d in layer.param_descriptors():
                n = 1
                for s in d.shape:
                    n *= s
                (train if d.trainable else frozen).append((layer, d, n))
        # aligned back-to-back offsets from the native planner (csrc/runtime/planner.hpp::plan_slab)
        toff, boffs = _plan_slab([n for _, _, n in train]), _plan_slab([n for _, _, n in frozen])
        entries = [(layer, d, o, n) for (layer, d, n), o in zip(train, toff)]
        buffers = [(layer, d, o, n) for (layer, d, n), o in zip(frozen, boffs)]
        off, boff = toff[-1], boffs[-1]
        self.numel = off
        self.param_slab = alloc(max(off, ALIGN_ELEMS), torch.float32)
        self.grad_slab = alloc(max(off, ALIGN_ELEMS), torch.float32)
        self.lp_slab = alloc(max(off, ALIGN_ELEMS), torch.bfloat16) if device.type == "cuda" else None
        self.buffer_slab = torch.zeros(max(boff, ALIGN_ELEMS), dtype=torch.float32, device=device)
        self.grad_slab.zero_()
        self.entries = entries
        self.buffers = buffers
        for layer, d, o, n in entries:
            pv = self.param_slab[o:o + n].view(d.shape)
            pv.copy_(layer.params[d.name])
            gv = self.grad_slab[o:o + n].view(d.shape)
            lv = self.lp_slab[o:o + n].view(d.shape) if self.lp_slab is not None else None
            layer.rebind(d.name, pv, gv, lv)
        for layer, d, o, n in buffers:
            bv = self.buffer_slab[o:o + n].view(d.shape)
            bv.copy_(layer.params[d.name])
            layer.rebind(d.name, bv, None, None)
        self.refresh_lp()

    def refresh_lp(self):
        if self.lp_slab is not None:
            from ..ops import native
            native.require().cast_f32_to_bf16(self.param_slab, self.lp_slab)
            for layer in self.layers:
                layer.mark_lp_fresh()
        else:
            for layer in self.layers:
                layer.mark_params_dirty()

    def parameters(self):
        """(qualified name, tensor) in registration order: trainable params then buffers, per layer."""
        out = []
        by_layer = defaultdict(list)
        for layer, d, o, n in self.entries + self.buffers:
            by_layer[id(layer)].append((layer, d))
        seen = set()
        for layer in self.layers:
            if id(layer) in seen:
                continue
            seen.add(id(layer))
            for lyr, d in by_layer.get(id(layer), []):
                out.append((f"{lyr.name}.{d.name}", lyr.params[d.name]))
        return out

    def zero_grads(self):
        self.grad_slab.zero_()

    def num_parameters(self) -> int:
        return sum(n for _, _, _, n in self.entries)


class Graph:
    def __init__(self, builder: GraphBuilder, device: torch.device, seed=None, alloc_fn=None):
        self.name = builder.name
        self.io_nodes = builder.io_nodes
        self.op_nodes = builder.op_nodes
        self.edges = builder.topo_sort()
        self._device = device
        flat = []
        for node in self.op_nodes.values():
            if not node.layer._initialized:
                node.layer.init(device, seed)
            else:
                node.layer.to(device)
            flat += node.layer.all_layers()
        self._flat_layers = flat
        self.ctx = GraphContext(flat, device, alloc_fn)
        self.training = True

    # -- accessors (reference: graph.hpp:18-191)
    def context(self) -> GraphContext:
        return self.ctx

    def device(self) -> torch.device:
        return self._device

    def get_layers(self) -> list[Layer]:
        return [n.layer for n in self.op_nodes.values()]

    def set_training(self, flag: bool):
        self.training = flag
        for l in self.get_layers():
            l.set_training(flag)

    def set_seed(self, seed: int):
        for l in self.get_layers():
            l.set_seed(seed)

    def get_config(self) -> TConfig:
        """GraphConfig in the reference schema (include/nn/graph.hpp:78-117): a TConfig of type "graph" whose
        parameters op_nodes / io_nodes / edges are json arrays of NodeConfig::to_json() / edge objects."""
        return TConfig(self.name, "graph",
                       op_nodes=[n.get_config() for n in self.op_nodes.values()],
                       io_nodes=[n.get_config() for n in self.io_nodes.values()],
                       edges=[e.get_config() for e in self.edges])

    @staticmethod
    def create_from_config(cfg, device="cpu", seed=None, alloc_fn=None) -> "Graph":
        """Accepts the reference schema (graph.hpp:128-169) as TConfig / its JSON, and the flat round-1 form
        ({name, op_nodes:[{uid,name,layer_config}], io_nodes:[{uid}], edges:[{inputs,outputs,op}]})."""
        if isinstance(cfg, dict) and "parameters" in cfg:
            cfg = TConfig.from_json(cfg)
        if isinstance(cfg, TConfig):
            name = cfg.name or "graph"
            ios = [(n.get("uid") if isinstance(n, TConfig) else n["uid"]) for n in cfg.get("io_nodes", [])]
            ops = [(n.get("uid"), n.get("layer_config")) if isinstance(n, TConfig) else (n["uid"], n["layer_config"])
                   for n in cfg.get("op_nodes", [])]
            edges = cfg.get("edges", [])
        else:
            name = cfg.get("name", "graph")
            ios = [n["uid"] for n in cfg["io_nodes"]]
            ops = [(n["uid"], n["layer_config"]) for n in cfg["op_nodes"]]
            edges = cfg["edges"]
        b = GraphBuilder(name)
        for u in ios:
            b.io(u)
        for uid, lc in ops:
            b.add_layer(LayerFactory.create(lc if isinstance(lc, TConfig) else TConfig.from_json(lc)), uid)
        for e in edges:
            b.add_edge(e.get("producers", e.get("inputs")), e.get("consumers", e.get("outputs")), e.get("op_node", e.get("op")))
        return b.compile(device, seed, alloc_fn)

    # -- checkpoint: pretty JSON config immediately followed by tensor records (graph.hpp:119-183)
    def save_state(self, f, extra_sections: dict | None = None):
        close = False
        if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
            f = open(f, "wb")
            close = True
        try:
            f.write(json.dumps(self.get_config().to_json(), indent=4).encode())
            for _, t in self.ctx.parameters():
                tensor_io.save_tensor(f, t)
            if extra_sections:
                tensor_io.write_trailer(f, extra_sections)
        finally:
            if close:
                f.close()

    def load_state_tensors(self, f):
        for name, t in self.ctx.parameters():
            tensor_io.load_into(f, t)
        self.ctx.refresh_lp()

    @staticmethod
    def load_state(f, device="cpu", alloc_fn=None, return_extra=False):
        close = False
        if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
            f = open(f, "rb")
            close = True
        try:
            data = f.read()
        finally:
            if close:
                f.close()
        cfg, end = json.JSONDecoder().raw_decode(data.decode("latin-1"))
        g = Graph.create_from_config(cfg, device, None, alloc_fn)
        bio = io.BytesIO(data[end:])
        g.load_state_tensors(bio)
        extra = tensor_io.read_trailer(bio)
        return (g, extra) if return_extra else g


class GraphExecutor:
    """Runs the compiled edge list (graph_executor.hpp:15-163).  Inputs/outputs are named packs."""

    def __init__(self, graph: Graph, allocator=None):
        self.graph = graph
        self.allocator = allocator

    def forward(self, inputs: dict, outputs: dict | None = None, mb_id: int = 0) -> dict:
        vals = dict(inputs)
        for e in self.graph.edges:
            layer = self.graph.op_nodes[e.op].layer
            ins = [vals[i] for i in e.inputs]
            out = layer.forward(ins if len(ins) > 1 else ins[0], mb_id)
            if isinstance(out, (list, tuple)):
                for o, t in zip(e.outputs, out):
                    vals[o] = t
            else:
                vals[e.outputs[0]] = out
        produced = {o for e in self.graph.edges for o in e.outputs}
        consumed = {i for e in self.graph.edges for i in e.inputs}
        result = {k: vals[k] for k in produced - consumed}
        if outputs is not None:
            outputs.update(result)
        return result

    def backward(self, grads: dict, grad_inputs: dict | None = None, mb_id: int = 0) -> dict:
        g = dict(grads)
        for e in reversed(self.graph.edges):
            layer = self.graph.op_nodes[e.op].layer
            gouts = [g.get(o) for o in e.outputs]
            if all(x is None for x in gouts):
                continue
            gin = layer.backward(gouts if len(gouts) > 1 else gouts[0], mb_id)
            if not isinstance(gin, (list, tuple)):
                gin = [gin]
            for i, t in zip(e.inputs, gin):
                if t is None:
                    continue
                g[i] = t if g.get(i) is None else g[i] + t
        produced = {o for e in self.graph.edges for o in e.outputs}
        result = {k: v for k, v in g.items() if k not in produced and k not in grads}
        if grad_inputs is not None:
            grad_inputs.update(result)
        return result


def graph_from_sequential(seq: Layer, device="cpu", seed=None, name=None, alloc_fn=None) -> Graph:
    """input -> [model] -> output, the shape every zoo model uses (example_models.hpp:49-84)."""
    b = GraphBuilder(name or seq.name)
    b.io("input")
    b.io("output")
    node = b.add_layer(seq, "model")
    b.add_edge(["input"], ["output"], node)
    return b.compile(device, seed, alloc_fn)
