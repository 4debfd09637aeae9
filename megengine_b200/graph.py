"""Graph-level use of the B200 operators: what the reference's `src/opr` ConvBias node,
`gopt` fuse passes and fastrun do around megdnn, reduced to the two pieces the ConvBias path needs.

  AlgoChooser   the PROFILE strategy of src/rdnn/impl/algo_chooser.cpp:1156-1195 (AlgoChooser<Opr>::setup_algo ->
                choose_by_profile): enumerate get_all_algorithms_info(), time each candidate on the real layouts
                with its own workspace, cache the winner per (operator type, param, layouts) -- the in-process
                analogue of the persistent fastrun cache -- and pin it through execution_policy.
  Sequential    a static inference graph in the order `gopt` leaves it after fusing conv + bias + activation
                (+ residual z) into ConvBias (src/gopt/impl/inference.cpp:2140-2273): ConvBias / Pooling /
                Elemwise / TypeCvt nodes over pre-allocated tensors, one stream, optional whole-graph CUDA graph
                capture (launch-bound small-batch inference replays as one graph launch).

The reference's graph runtime itself (src/core, src/opr, src/gopt: ~300 kLoC with generated sources and external
dependencies) cannot be built in this image; tests/test_graph_gpu.py therefore proves the composition end to end
against the chain of oracles instead: a QuantizedS8 ResNet-style network, every algorithm chosen by profiling,
bit-exact from the input image to the pooled features.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import _capi as C
from .conv_bias import (AlgorithmInfo, ConvBiasForward, ExecutionPolicy, Handle, MegDNNError, TensorLayout, TensorND,
                        empty_like_dtype, megdnn_assert)
from .nn_ops import ElemwiseForward, ElemwiseMultiType, PoolingForward, TypeCvt


class AlgoChooser:
    """fastrun for ConvBiasForward on the B200 handle."""

    _cache: Dict[tuple, AlgorithmInfo] = {}

    def __init__(self, handle: Handle, warmup: int = 2, runs: int = 5):
        self.handle, self.warmup, self.runs = handle, warmup, runs
        self.profiled = 0

    @staticmethod
    def _key(opr: ConvBiasForward, lays):
        p = opr.param
        return ("ConvBiasForward", p.nonlineMode, p.mode, p.sparse, p.format, p.pad_h, p.pad_w, p.stride_h, p.stride_w,
                p.dilate_h, p.dilate_w, p.compute_mode) + tuple(ConvBiasForward._layout_key(l) for l in lays)

    def profile(self, opr: ConvBiasForward, src, flt, bias, z, dst):
        """time every available algorithm (choose_by_profile); returns [(AlgorithmInfo, ms)] sorted by time."""
        lays = (src.layout, flt.layout, bias.layout, z.layout, dst.layout)
        saved = opr.execution_policy
        out = []
        try:
            for info in opr.get_all_algorithms_info(*lays):
                opr.execution_policy = ExecutionPolicy(info)
                nbytes = opr.get_workspace_in_bytes(*lays)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self.handle.device) if nbytes else None
                for _ in range(self.warmup):
                    opr.exec(src, flt, bias, z, dst, None, ws)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                stream = self.handle.stream
                e0.record(stream)
                for _ in range(self.runs):
                    opr.exec(src, flt, bias, z, dst, None, ws)
                e1.record(stream)
                e1.synchronize()
                out.append((info, e0.elapsed_time(e1) / self.runs))
        finally:
            opr.execution_policy = saved
        megdnn_assert(len(out) > 0, "no usable conv_bias fwd algorithm")
        self.profiled += 1
        return sorted(out, key=lambda t: t[1])

    def setup_algo(self, opr: ConvBiasForward, src, flt, bias, z, dst) -> AlgorithmInfo:
        """AlgoChooser<Opr>::setup_algo with strategy PROFILE: cached winner, else profile now."""
        key = self._key(opr, (src.layout, flt.layout, bias.layout, z.layout, dst.layout))
        best = AlgoChooser._cache.get(key)
        if best is None:
            best = self.profile(opr, src, flt, bias, z, dst)[0][0]
            AlgoChooser._cache[key] = best
        opr.execution_policy = ExecutionPolicy(best)
        return best


@dataclass
class Node:
    kind: str                     # "conv_bias" | "pooling" | "elemwise" | "elemwise_q" | "typecvt"
    opr: object
    inputs: List[str]             # names of the tensors it reads (conv_bias: src, filter, bias or "", z or "")
    output: str
    workspace: Optional[torch.Tensor] = None
    algo: Optional[AlgorithmInfo] = None


class Sequential:
    """static inference graph over named tensors; exec() enqueues every node on the handle's stream."""

    def __init__(self, handle: Handle, fastrun: bool = True):
        self.handle = handle
        self.tensors: Dict[str, TensorND] = {}
        self.nodes: List[Node] = []
        self.chooser = AlgoChooser(handle) if fastrun else None
        self._graph: Optional[torch.cuda.CUDAGraph] = None

    # ---- tensors
    def add_tensor(self, name: str, t: TensorND) -> TensorND:
        self.tensors[name] = t
        return t

    def new_tensor(self, name: str, layout: TensorLayout) -> TensorND:
        t = torch.empty(tuple(layout.shape), dtype=layout.dtype.torch_dtype, device=self.handle.device)
        return self.add_tensor(name, TensorND(t, TensorLayout(tuple(layout.shape), layout.dtype)))

    def _get(self, name: str, absent_dtype=None) -> TensorND:
        return self.tensors[name] if name else empty_like_dtype(absent_dtype)

    # ---- nodes
    def conv_bias(self, out: str, src: str, flt: str, bias: str = "", z: str = "", dst_dtype=None, **param) -> TensorND:
        opr = ConvBiasForward(self.handle)
        for k, v in param.items():
            megdnn_assert(hasattr(opr.param, k), "unknown ConvBias param " + k)
            setattr(opr.param, k, v)
        s, f = self.tensors[src], self.tensors[flt]
        b = self._get(bias, dst_dtype)
        lay = opr.deduce_layout(s.layout, f.layout, b.layout, TensorLayout((), dst_dtype), dst_dtype)
        d = self.new_tensor(out, lay)
        zt = self._get(z, lay.dtype)
        node = Node("conv_bias", opr, [src, flt, bias, z], out)
        if self.chooser is not None:
            node.algo = self.chooser.setup_algo(opr, s, f, b, zt, d)
        nbytes = opr.get_workspace_in_bytes(s.layout, f.layout, b.layout, zt.layout, d.layout)
        node.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.handle.device) if nbytes else None
        self.nodes.append(node)
        return d

    def pooling(self, out: str, src: str, **param) -> TensorND:
        opr = PoolingForward(self.handle)
        for k, v in param.items():
            megdnn_assert(hasattr(opr.param, k), "unknown Pooling param " + k)
            setattr(opr.param, k, v)
        d = self.new_tensor(out, opr.deduce_layout(self.tensors[src].layout))
        self.nodes.append(Node("pooling", opr, [src], out))
        return d

    def elemwise(self, out: str, mode: int, a: str, b: str = "", dst_dtype=None) -> TensorND:
        ta = self.tensors[a]
        quant = ta.layout.dtype.enumv == C.DT_QS8
        opr = ElemwiseMultiType(self.handle) if quant else ElemwiseForward(self.handle)
        opr.mode = mode
        d = self.new_tensor(out, TensorLayout(tuple(ta.layout.shape), dst_dtype or ta.layout.dtype))
        self.nodes.append(Node("elemwise_q" if quant else "elemwise", opr, [a] + ([b] if b else []), out))
        return d

    def typecvt(self, out: str, src: str, dst_dtype) -> TensorND:
        d = self.new_tensor(out, TensorLayout(tuple(self.tensors[src].layout.shape), dst_dtype))
        self.nodes.append(Node("typecvt", TypeCvt(self.handle), [src], out))
        return d

    # ---- execution
    def _enqueue(self):
        T = self.tensors
        for n in self.nodes:
            if n.kind == "conv_bias":
                s, f, b, z = n.inputs
                d = T[n.output]
                n.opr.exec(T[s], T[f], self._get(b, d.layout.dtype), self._get(z, d.layout.dtype), d, None, n.workspace)
            elif n.kind == "pooling":
                n.opr.exec(T[n.inputs[0]], T[n.output])
            elif n.kind in ("elemwise", "elemwise_q"):
                n.opr.exec([T[i] for i in n.inputs], T[n.output])
            elif n.kind == "typecvt":
                n.opr.exec(T[n.inputs[0]], T[n.output])
            else:
                raise MegDNNError("unknown node kind " + n.kind)

    def exec(self):
        if self._graph is not None:
            self._graph.replay()
        else:
            self._enqueue()

    def capture(self):
        """record the whole network into one CUDA graph (every node only enqueues kernels on the stream)."""
        self._enqueue()   # warm: plans, tensor maps and function attributes exist before capture
        torch.cuda.synchronize(self.handle.device)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.handle.device)
        saved = self.handle._stream
        with torch.cuda.stream(side):
            self.handle._stream = side
            try:
                g.capture_begin()
                self._enqueue()
                g.capture_end()
            finally:
                self.handle._stream = saved
        self._graph = g
        return g
