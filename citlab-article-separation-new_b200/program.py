"""Lowering of a frozen GraphDef (GraphIR) to the fused op program the C engine executes.

This is the "import" half of the reference's ``load_graph``
(net_post_processing_helper.py:36-53: ``tf.import_graph_def``).  It is *op-type driven*, never
name driven: node names of the shipped ``.pb`` files are unknown.

Fusions performed (all are exact re-associations of what TF computes op by op):
  * ``Conv2D -> BiasAdd [-> Identity] [-> Add(residual)] [-> Relu]``  -> one ARU_OP_CONV
    (layers.py:191-247; residual block ARU_v1.py:212-227 - the residual operand is the
    *pre-activation* BiasAdd output of ``conv1``, so that conv exports both values);
  * ``Conv2DBackpropInput(3x3, s2) -> BiasAdd -> Relu``               -> ARU_OP_DECONV (layers.py:342-367);
  * ``ConcatV2(axis=3)``                                              -> no op: producers write channel
    slices of one buffer (ARU_v1.py:264);
  * ``upsample_simple`` x A -> ``ConcatV2 -> Softmax -> Split -> Mul -> AddN`` -> one ARU_OP_COMBINE
    (ARU_v1.py:115,137,145-153; the ones-filter transposed conv of layers.py:716-720 is recognised
    by its constant all-ones filter with k == stride and never run as a GEMM);
  * ``Shape / StridedSlice / Pack`` shape plumbing is resolved symbolically ("same spatial dims as
    tensor T") and re-evaluated by the engine for every (n, h, w).
Anything else raises ``UnsupportedGraphError`` - there is no slow path to fall back to.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Dict, List, Optional, Tuple

import numpy as np

from .graphdef import GraphIR, Node

MAX_SCALES = 8
OP_CONV, OP_DECONV, OP_MAXPOOL, OP_AVGPOOL, OP_COMBINE, OP_UPSUM, OP_COPY = 1, 2, 3, 4, 5, 6, 7
ACT_NONE, ACT_RELU, ACT_SOFTMAX, ACT_SIGMOID = 0, 1, 2, 3
ROLE_TMP, ROLE_INPUT, ROLE_OUTPUT = 0, 1, 2


class UnsupportedGraphError(RuntimeError):
    pass


# ctypes mirrors of include/aru_b200.h ------------------------------------------------------
class CView(ctypes.Structure):
    _fields_ = [("buf", ctypes.c_int32), ("ch_off", ctypes.c_int32), ("ch", ctypes.c_int32)]


class COp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("ksize", ctypes.c_int32), ("stride", ctypes.c_int32),
                ("act", ctypes.c_int32), ("n_scales", ctypes.c_int32), ("like_buf", ctypes.c_int32),
                ("w_off", ctypes.c_int64), ("b_off", ctypes.c_int64),
                ("inp", CView), ("out", CView), ("out_pre", CView), ("res", CView),
                ("att", CView * MAX_SCALES), ("det", CView * MAX_SCALES),
                ("up_att", ctypes.c_int32 * MAX_SCALES), ("up_det", ctypes.c_int32 * MAX_SCALES)]


class CBuffer(ctypes.Structure):
    _fields_ = [("channels", ctypes.c_int32), ("role", ctypes.c_int32)]


class CGraphDesc(ctypes.Structure):
    _fields_ = [("magic", ctypes.c_uint32), ("abi_version", ctypes.c_uint32),
                ("n_buffers", ctypes.c_int32), ("n_ops", ctypes.c_int32), ("n_weights", ctypes.c_int64),
                ("buffers", ctypes.POINTER(CBuffer)), ("ops", ctypes.POINTER(COp)),
                ("weights", ctypes.POINTER(ctypes.c_float))]


@dataclasses.dataclass(frozen=True)
class View:
    buf: int
    ch_off: int
    ch: int

    def c(self) -> CView:
        return CView(self.buf, self.ch_off, self.ch)


NOVIEW = View(-1, 0, 0)


@dataclasses.dataclass
class Op:
    kind: int
    ksize: int = 0
    stride: int = 0
    act: int = ACT_NONE
    like_buf: int = -1
    w_off: int = -1
    b_off: int = -1
    inp: View = NOVIEW
    out: View = NOVIEW
    out_pre: View = NOVIEW
    res: View = NOVIEW
    att: Tuple[View, ...] = ()
    det: Tuple[View, ...] = ()
    up_att: Tuple[int, ...] = ()
    up_det: Tuple[int, ...] = ()
    name: str = ""

    def c(self) -> COp:
        o = COp()
        o.kind, o.ksize, o.stride, o.act = self.kind, self.ksize, self.stride, self.act
        o.n_scales, o.like_buf, o.w_off, o.b_off = len(self.att), self.like_buf, self.w_off, self.b_off
        o.inp, o.out, o.out_pre, o.res = self.inp.c(), self.out.c(), self.out_pre.c(), self.res.c()
        for i in range(MAX_SCALES):
            o.att[i] = (self.att[i] if i < len(self.att) else NOVIEW).c()
            o.det[i] = (self.det[i] if i < len(self.det) else NOVIEW).c()
            o.up_att[i] = self.up_att[i] if i < len(self.up_att) else 0
            o.up_det[i] = self.up_det[i] if i < len(self.up_det) else 0
        return o


class Program:
    """Buffers + ops + weight blob; ``desc()`` yields the C struct aru_create() takes."""

    def __init__(self):
        self.buffers: List[List[int]] = []     # [channels, role]
        self.ops: List[Op] = []
        self._weights: List[np.ndarray] = []
        self._n_weights = 0
        self._weight_index: Dict[str, int] = {}
        self.n_class = 0
        self.output_buf = -1
        self.input_buf = -1
        self.tensor_of_node: Dict[str, View] = {}   # debugging / per-layer parity: GraphDef node -> view

    def new_buffer(self, channels: int, role: int = ROLE_TMP) -> int:
        self.buffers.append([channels, role])
        return len(self.buffers) - 1

    def add_weights(self, key: str, arr: np.ndarray) -> int:
        if key in self._weight_index:
            return self._weight_index[key]
        off = self._n_weights
        flat = np.ascontiguousarray(arr, np.float32).reshape(-1)
        self._weights.append(flat)
        self._n_weights += flat.size
        self._weight_index[key] = off
        return off

    @property
    def weights(self) -> np.ndarray:
        return np.concatenate(self._weights) if self._weights else np.zeros(0, np.float32)

    def desc(self):
        """Returns (CGraphDesc, keepalive) - keep the second value referenced while the desc is used."""
        bufs = (CBuffer * len(self.buffers))(*[CBuffer(c, r) for c, r in self.buffers])
        ops = (COp * len(self.ops))(*[o.c() for o in self.ops])
        w = self.weights
        d = CGraphDesc(0x31555241, 1, len(self.buffers), len(self.ops), w.size, bufs, ops,
                       w.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return d, (bufs, ops, w)


class _Lowering:
    def __init__(self, ir: GraphIR, in_name: str, out_name: str):
        self.ir = ir
        self.in_name, self.out_name = in_name, out_name
        self.p = Program()
        self.memo: Dict[Tuple[str, int], View] = {}
        self.dest: Dict[Tuple[str, int], View] = {}
        self.cons = ir.consumers()
        self.conv_ops: Dict[str, Op] = {}      # BiasAdd node name -> emitted conv op (for out_pre export)
        self.concat_parts: Dict[str, tuple] = {}
        self.concat_done = set()

    # ---- helpers ---------------------------------------------------------------------------
    def node(self, edge) -> Node:
        return self.ir[edge[0]]

    def skip_identity(self, edge):
        return self.ir.resolve_identity(edge)

    def consumers_through_identity(self, name: str) -> List[str]:
        out = []
        for c in self.cons[name]:
            if self.ir[c].op == "Identity":
                out.extend(self.consumers_through_identity(c))
            else:
                out.append(c)
        return out

    def const(self, edge) -> np.ndarray:
        v = self.ir.const_value(edge)
        if v is None:
            raise UnsupportedGraphError(f"expected a constant at {edge[0]}")
        return v

    def channels(self, edge) -> int:
        """Static channel count of a float tensor."""
        name, idx = self.skip_identity(edge)
        nd = self.ir[name]
        if nd.op == "Placeholder":
            return 1
        if nd.op == "Conv2D":
            return int(self.const(nd.inputs[1]).shape[3])
        if nd.op == "Conv2DBackpropInput":
            return int(self.const(nd.inputs[1]).shape[2])
        if nd.op == "ConcatV2":
            return sum(self.channels(e) for e in nd.inputs[:-1])
        if nd.op == "Split":
            return self.channels(nd.inputs[1]) // int(nd.attrs["num_split"])
        if nd.op in ("Mul", "Add", "AddV2", "AddN"):
            return max(self.channels(e) for e in nd.inputs)
        if nd.op in ("BiasAdd", "Relu", "MaxPool", "AvgPool", "Softmax", "Sigmoid", "Elu"):
            return self.channels(nd.inputs[0])
        raise UnsupportedGraphError(f"cannot infer channels of {nd.op} ({name})")

    def alloc(self, edge, ch: int) -> View:
        """Output location for the value of `edge`: a pre-assigned concat slice or a fresh buffer."""
        edge = self.skip_identity(edge)
        if edge in self.dest:
            v = self.dest.pop(edge)
            assert v.ch == ch
            return v
        return View(self.p.new_buffer(ch), 0, ch)

    def like_of_shape(self, edge) -> Tuple[str, int]:
        """Resolve an int32 output_shape tensor to 'spatial dims of float tensor T'."""
        name, _ = self.skip_identity(edge)
        nd = self.ir[name]
        if nd.op == "Shape":
            return self.skip_identity(nd.inputs[0])
        if nd.op == "Pack":                       # tf.stack([shape[0], shape[1], shape[2], C])
            srcs = set()
            for e in nd.inputs[:3]:
                s, _ = self.skip_identity(e)
                sn = self.ir[s]
                if sn.op != "StridedSlice":
                    raise UnsupportedGraphError(f"Pack input {s} is not a shape slice")
                sh, _ = self.skip_identity(sn.inputs[0])
                if self.ir[sh].op != "Shape":
                    raise UnsupportedGraphError(f"StridedSlice {s} does not slice a Shape")
                srcs.add(self.skip_identity(self.ir[sh].inputs[0]))
            if len(srcs) != 1:
                raise UnsupportedGraphError("output_shape mixes dims of several tensors")
            return srcs.pop()
        raise UnsupportedGraphError(f"unsupported output_shape producer {nd.op} ({name})")

    # ---- pre-pass: concat destinations -----------------------------------------------------
    def assign_concat_slices(self, order: List[str]):
        for name in order:
            nd = self.ir[name]
            if nd.op != "ConcatV2":
                continue
            axis = int(self.const(nd.inputs[-1]))
            if axis not in (3, -1):
                raise UnsupportedGraphError(f"ConcatV2 axis {axis} ({name})")
            parts = nd.inputs[:-1]
            chs = [self.channels(e) for e in parts]
            if all(c == 1 for c in chs):
                continue                           # attention concat: consumed by the COMBINE pattern
            if any(c % 8 for c in chs[:-1]):
                raise UnsupportedGraphError(f"ConcatV2 {name}: channel offsets must be multiples of 8, got {chs}")
            buf = self.p.new_buffer(sum(chs))
            off = 0
            for e, c in zip(parts, chs):
                key = self.skip_identity(e)
                if key not in self.dest:
                    self.dest[key] = View(buf, off, c)
                off += c
            self.concat_parts[name] = (buf, list(zip(parts, chs)))

    # ---- main recursion -----------------------------------------------------------------------
    def lower(self, edge) -> View:
        edge = self.skip_identity(edge)
        name, idx = edge
        nd = self.ir[name]
        if nd.op == "ConcatV2" and name in self.concat_parts and name not in self.concat_done:
            return self.lower_concat(name)
        if edge in self.memo:
            return self.memo[edge]
        op = nd.op
        if op == "Placeholder":
            if name != self.in_name:
                raise UnsupportedGraphError(f"unexpected second input {name}")
            v = View(self.p.new_buffer(1, ROLE_INPUT), 0, 1)
            self.p.input_buf = v.buf
        elif op == "Relu":
            v = self.lower_relu(edge)
        elif op == "BiasAdd":
            v = self.lower_conv_like(edge, ACT_NONE)
        elif op in ("MaxPool", "AvgPool"):
            v = self.lower_pool(edge)
        elif op == "AddN":
            v = self.lower_combine(edge)
        elif op == "Conv2DBackpropInput":
            v = self.lower_upsum(edge)
        elif op in ("Softmax", "Sigmoid"):
            raise UnsupportedGraphError(f"{op} ({name}) outside the output head / attention pattern")
        else:
            raise UnsupportedGraphError(f"unsupported op {op} ({name})")
        self.memo[edge] = v
        self.p.tensor_of_node[name] = v
        return v

    def lower_concat(self, name: str) -> View:
        nd = self.ir[name]
        buf, parts = self.concat_parts[name]
        self.concat_done.add(name)
        off = 0
        for e, c in parts:
            want = View(buf, off, c)
            got = self.lower(e)
            if got != want:                          # produced elsewhere first: materialise a copy
                self.p.ops.append(Op(OP_COPY, inp=got, out=want, name=name + "/copy"))
            off += c
        v = View(buf, 0, off)
        self.memo[(name, 0)] = v
        self.p.tensor_of_node[name] = v
        return v

    def lower_relu(self, edge) -> View:
        nd = self.node(edge)
        src = self.skip_identity(nd.inputs[0])
        sn = self.ir[src[0]]
        if sn.op == "BiasAdd":
            return self.lower_conv_like(src, ACT_RELU, relu_edge=edge)
        if sn.op in ("Add", "AddV2"):
            return self.lower_residual(edge, src)
        raise UnsupportedGraphError(f"Relu over {sn.op} ({sn.name}) is not a fusable pattern")

    def _conv_parts(self, bias_edge):
        bn = self.node(bias_edge)
        prod = self.skip_identity(bn.inputs[0])
        pn = self.ir[prod[0]]
        bias = self.const(bn.inputs[1])
        return bn, pn, bias

    def lower_conv_like(self, bias_edge, act: int, relu_edge=None, res: Optional[View] = None,
                        out_edge=None) -> View:
        """BiasAdd(Conv2D | Conv2DBackpropInput) with an optional fused activation / residual.
        `out_edge` is the graph edge whose value the op's main output holds."""
        bn, pn, bias = self._conv_parts(bias_edge)
        bname = bn.name
        if pn.op == "Conv2D":
            w = self.const(pn.inputs[1])
            kh, kw, cin, cout = w.shape
            self._check_conv_attrs(pn)
            if kh != kw or kh not in (3, 4):
                raise UnsupportedGraphError(f"Conv2D {pn.name}: kernel {kh}x{kw} not supported")
            if bname in self.conv_ops:
                # the same conv value requested a second time, through another consumer
                op = self.conv_ops[bname]
                if res is not None or op.res.buf >= 0:
                    raise UnsupportedGraphError(f"conv {bname}: residual-fused conv consumed twice")
                if act == ACT_NONE:                 # want the pre-activation value (ARU_v1.py:214 orig_x)
                    if op.act == ACT_NONE:
                        return op.out
                    if op.out_pre.buf < 0:
                        op.out_pre = self.alloc(bias_edge, int(cout))
                    return op.out_pre
                if act == ACT_RELU:                 # want relu(value)
                    if op.act == ACT_RELU:
                        return op.out
                    if op.act == ACT_NONE:          # emitted as pre-activation first: export both
                        op.out_pre, op.act = op.out, ACT_RELU
                        op.out = self.alloc(relu_edge, int(cout))
                        return op.out
                raise UnsupportedGraphError(f"conv {bname} consumed through two different activations")
            x = self.lower(pn.inputs[0])
            if x.ch != cin:
                raise UnsupportedGraphError(f"Conv2D {pn.name}: input has {x.ch} channels, filter wants {cin}")
            key = self.skip_identity(pn.inputs[1])[0]
            op = Op(OP_CONV, ksize=int(kh), act=act, inp=x, name=bname,
                    w_off=self.p.add_weights(key, w),
                    b_off=self.p.add_weights(self.skip_identity(bn.inputs[1])[0], bias))
            if res is not None:
                op.res = res
            op.out = self.alloc(out_edge or relu_edge or bias_edge, int(cout))
            self.p.ops.append(op)
            self.conv_ops[bname] = op
            return op.out
        if pn.op == "Conv2DBackpropInput":
            w = self.const(pn.inputs[1])
            kh, kw, cout, cin = w.shape
            s = pn.attrs["strides"]
            if not (kh == kw == 3 and s == [1, 2, 2, 1] and pn.attrs.get("padding") == "SAME"):
                raise UnsupportedGraphError(f"Conv2DBackpropInput {pn.name}: only 3x3 stride 2 SAME is supported")
            like = self.lower(self.like_of_shape(pn.inputs[0]))
            x = self.lower(pn.inputs[2])
            if x.ch != cin:
                raise UnsupportedGraphError(f"deconv {pn.name}: channel mismatch")
            op = Op(OP_DECONV, ksize=3, stride=2, act=act, inp=x, like_buf=like.buf, name=bname,
                    w_off=self.p.add_weights(self.skip_identity(pn.inputs[1])[0], w),
                    b_off=self.p.add_weights(self.skip_identity(bn.inputs[1])[0], bias))
            op.out = self.alloc(relu_edge or bias_edge, int(cout))
            self.p.ops.append(op)
            return op.out
        raise UnsupportedGraphError(f"BiasAdd over {pn.op} ({pn.name})")

    def _check_conv_attrs(self, pn: Node):
        a = pn.attrs
        if a.get("padding") != "SAME" or a.get("strides", [1, 1, 1, 1]) != [1, 1, 1, 1] or \
                a.get("data_format", "NHWC") != "NHWC" or a.get("dilations", [1, 1, 1, 1]) != [1, 1, 1, 1]:
            raise UnsupportedGraphError(f"Conv2D {pn.name}: only NHWC / SAME / stride 1 / dilation 1")

    def lower_residual(self, relu_edge, add_edge) -> View:
        """Relu(Add(a, b)) where one side is a conv consumed only here (ARU_v1.py:225-226)."""
        an = self.node(add_edge)
        sides = [self.skip_identity(e) for e in an.inputs]
        main = None
        for i, s in enumerate(sides):
            sn = self.ir[s[0]]
            if sn.op == "BiasAdd" and self.ir[self.skip_identity(sn.inputs[0])[0]].op == "Conv2D" \
                    and len(self.consumers_through_identity(s[0])) == 1 and s[0] not in self.conv_ops:
                main = i
                break
        if main is None or len(self.cons[an.name]) != 1:
            raise UnsupportedGraphError(f"Add {an.name}: not a fusable residual pattern")
        res = self.lower(sides[1 - main])
        return self.lower_conv_like(sides[main], ACT_RELU, res=res, out_edge=relu_edge)

    def lower_pool(self, edge) -> View:
        nd = self.node(edge)
        a = nd.attrs
        if a.get("ksize") != [1, 2, 2, 1] or a.get("strides") != [1, 2, 2, 1] or a.get("padding") != "SAME":
            raise UnsupportedGraphError(f"{nd.op} {nd.name}: only 2x2 stride 2 SAME")
        x = self.lower(nd.inputs[0])
        out = self.alloc(edge, x.ch)
        self.p.ops.append(Op(OP_MAXPOOL if nd.op == "MaxPool" else OP_AVGPOOL, ksize=2, stride=2, inp=x, out=out,
                             name=nd.name))
        return out

    def _match_upsum(self, edge):
        """Conv2DBackpropInput with a constant all-ones [up,up,C,C] filter and stride up (layers.py:716-720).
        Returns (source edge, up, like edge) or None."""
        name, _ = self.skip_identity(edge)
        nd = self.ir[name]
        if nd.op != "Conv2DBackpropInput":
            return None
        w = self.ir.const_value(nd.inputs[1])
        if w is None:
            return None
        up = int(w.shape[0])
        if w.shape[0] != w.shape[1] or nd.attrs["strides"] != [1, up, up, 1] or not np.all(w == 1.0) \
                or w.shape[2] != w.shape[3]:
            return None
        if nd.attrs.get("padding", "SAME") != "SAME":
            raise UnsupportedGraphError(f"upsample {name}: padding {nd.attrs.get('padding')}")
        return nd.inputs[2], up, self.like_of_shape(nd.inputs[0])

    def lower_upsum(self, edge) -> View:
        m = self._match_upsum(edge)
        if m is None:
            raise UnsupportedGraphError(f"bare Conv2DBackpropInput {edge[0]} (no BiasAdd, not an all-ones upsample)")
        src, up, like = m
        x = self.lower(src)
        lk = self.lower(like)
        out = self.alloc(edge, x.ch)
        self.p.ops.append(Op(OP_UPSUM, stride=up, inp=x, out=out, like_buf=lk.buf, name=edge[0]))
        return out

    def lower_combine(self, edge) -> View:
        """AddN_k( Mul(det_k, Split(Softmax(ConcatV2(att_0..att_{A-1})))[k]) ), ARU_v1.py:145-153."""
        nd = self.node(edge)
        atts, dets, ups_a, ups_d = {}, {}, {}, {}
        like = None
        n_scales = None
        for e in nd.inputs:
            mn = self.node(self.skip_identity(e))
            if mn.op != "Mul":
                raise UnsupportedGraphError(f"AddN {nd.name}: operand {mn.name} is {mn.op}, expected Mul")
            sides = [self.skip_identity(x) for x in mn.inputs]
            k_side = [i for i, s in enumerate(sides) if self.ir[s[0]].op == "Split"]
            if len(k_side) != 1:
                raise UnsupportedGraphError(f"Mul {mn.name}: expected exactly one Split operand")
            sp_edge = sides[k_side[0]]
            sp = self.ir[sp_edge[0]]
            k = sp_edge[1]
            sm = self.node(self.skip_identity(sp.inputs[1]))
            if sm.op != "Softmax":
                raise UnsupportedGraphError(f"Split {sp.name} is not fed by Softmax")
            cc = self.node(self.skip_identity(sm.inputs[0]))
            if cc.op != "ConcatV2" or int(self.const(cc.inputs[-1])) not in (3, -1):
                raise UnsupportedGraphError(f"Softmax {sm.name} is not fed by a channel concat")
            parts = cc.inputs[:-1]
            if n_scales is None:
                n_scales = len(parts)
                if n_scales != int(sp.attrs["num_split"]) or n_scales > MAX_SCALES:
                    raise UnsupportedGraphError("attention split / concat arity mismatch")
            m = self._match_upsum(parts[k])
            if m is None or self.channels(m[0]) != 1:
                raise UnsupportedGraphError(f"attention map {k} is not a 1-channel upsample_simple")
            atts[k], ups_a[k], lk = m
            like = like or lk
            det_edge = sides[1 - k_side[0]]
            md = self._match_upsum(det_edge)
            if md is None:
                dets[k], ups_d[k] = det_edge, 1
            else:
                dets[k], ups_d[k], _ = md
        if sorted(atts) != list(range(n_scales)):
            raise UnsupportedGraphError("attention scales are not a complete 0..A-1 set")
        ch = max(self.channels(dets[k]) for k in range(n_scales))
        det_views = [self.lower(dets[k]) for k in range(n_scales)]
        att_views = [self.lower(atts[k]) for k in range(n_scales)]
        lk = self.lower(like)
        out = self.alloc(edge, ch)
        self.p.ops.append(Op(OP_COMBINE, like_buf=lk.buf, out=out, att=tuple(att_views), det=tuple(det_views),
                             up_att=tuple(ups_a[k] for k in range(n_scales)),
                             up_det=tuple(ups_d[k] for k in range(n_scales)), name=nd.name))
        return out

    # ---- entry ---------------------------------------------------------------------------------
    def run(self) -> Program:
        ir = self.ir
        if self.in_name not in ir or self.out_name not in ir:
            raise KeyError(f"The name '{self.in_name}:0' or '{self.out_name}:0' refers to a Tensor which does "
                           f"not exist in the graph")            # same failure class as tf get_tensor_by_name
        order = ir.topo_order([self.out_name])
        for n in order:
            if ir[n].op in ("Enter", "Exit", "Merge", "Switch", "NextIteration", "LoopCond", "FusedBatchNorm"):
                raise UnsupportedGraphError(f"control-flow / batch-norm op {ir[n].op} ({n}) is not supported")
        self.assign_concat_slices(order)
        on = ir[self.out_name]
        act = {"Softmax": ACT_SOFTMAX, "Sigmoid": ACT_SIGMOID, "Identity": ACT_NONE}.get(on.op)
        if act is None:
            raise UnsupportedGraphError(f"output op {on.op}: expected Softmax / Sigmoid / Identity")
        head = self.skip_identity(on.inputs[0] if on.op != "Identity" else (self.out_name, 0))
        hn = self.ir[head[0]]
        if hn.op != "BiasAdd":
            raise UnsupportedGraphError(f"output head is {hn.op}; expected conv + BiasAdd")
        bn, pn, bias = self._conv_parts(head)
        n_class = int(bias.shape[0])
        out_buf = self.p.new_buffer(n_class, ROLE_OUTPUT)
        self.dest[head] = View(out_buf, 0, n_class)
        self.lower_conv_like(head, act)
        self.p.n_class, self.p.output_buf = n_class, out_buf
        self.p.tensor_of_node[self.out_name] = View(out_buf, 0, n_class)
        return self.p


def lower_graph(ir: GraphIR, in_name: str = "inImg", out_name: str = "output") -> Program:
    return _Lowering(ir, in_name, out_name).run()
