"""Frozen TensorFlow GraphDef I/O without TensorFlow.

Replaces the parsing half of the reference's ``load_graph``
(article_separation/image_segmentation/net_post_processing/net_post_processing_helper.py:36-53):
the reference does ``GraphDef.ParseFromString`` + ``tf.import_graph_def``; here the
protobuf is parsed with ``tensorboard.compat.proto`` (a pure-protobuf copy of the TF
schema) into a small typed IR that the lowering pass (``program.py``) and the CPU oracle
(``oracle/aru_oracle.py``) both consume.

Because the shipped ``separator_detection_net.pb`` / ``heading_detection_net.pb`` are
stripped from the reference tree (``.MISSING_LARGE_BLOBS``), this module also holds a
*writer* that emits a synthetic frozen ARU-Net with exactly the node structure TF1
produces for ``article_separation/backbones/ARU_v1.py:62-294`` and
``article_separation/gnn/model/graph_util/layers.py`` (conv2d :191-247, deconv2d
:342-367, pools :526-544, upsample_simple :716-720), with the pinned I/O tensor names
``inImg:0`` / ``output:0`` (net_post_processing_helper.py:69-70).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from tensorboard.compat.proto import graph_pb2, types_pb2
from tensorboard.util import tensor_util

__all__ = ["Node", "GraphIR", "parse_graphdef", "load_pb", "GraphBuilder", "build_aru_graphdef",
           "aru_conv_macs"]


# --------------------------------------------------------------------------------------
# IR
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Node:
    name: str
    op: str
    inputs: List[Tuple[str, int]]           # (producer name, output index); control deps dropped
    attrs: Dict[str, object]
    value: Optional[np.ndarray] = None       # Const payload


class GraphIR:
    """Name-indexed view of a frozen graph. Edges are (name, output_index) pairs."""

    def __init__(self, nodes: Sequence[Node]):
        self.nodes: Dict[str, Node] = {n.name: n for n in nodes}
        self.order: List[str] = [n.name for n in nodes]

    def __getitem__(self, name: str) -> Node:
        return self.nodes[name]

    def __contains__(self, name: str) -> bool:
        return name in self.nodes

    def consumers(self) -> Dict[str, List[str]]:
        out: Dict[str, List[str]] = {n: [] for n in self.nodes}
        for n in self.nodes.values():
            for src, _ in n.inputs:
                out[src].append(n.name)
        return out

    def resolve_identity(self, edge: Tuple[str, int]) -> Tuple[str, int]:
        """Follow Identity chains (``<var>/read`` nodes of frozen variables)."""
        name, idx = edge
        while self.nodes[name].op == "Identity":
            name, idx = self.nodes[name].inputs[0]
        return name, idx

    def const_value(self, edge: Tuple[str, int]) -> Optional[np.ndarray]:
        name, _ = self.resolve_identity(edge)
        n = self.nodes[name]
        return n.value if n.op == "Const" else None

    def topo_order(self, outputs: Sequence[str]) -> List[str]:
        seen, order = set(), []
        stack = [(o, False) for o in outputs]
        while stack:
            name, done = stack.pop()
            if done:
                order.append(name)
                continue
            if name in seen:
                continue
            seen.add(name)
            stack.append((name, True))
            for src, _ in reversed(self.nodes[name].inputs):
                if src not in seen:
                    stack.append((src, False))
        return order


def _attr_to_py(a):
    kind = a.WhichOneof("value")
    if kind == "s":
        return a.s.decode("utf-8", "replace")
    if kind == "i":
        return int(a.i)
    if kind == "f":
        return float(a.f)
    if kind == "b":
        return bool(a.b)
    if kind == "type":
        return int(a.type)
    if kind == "shape":
        return [d.size for d in a.shape.dim]
    if kind == "list":
        if a.list.i:
            return [int(v) for v in a.list.i]
        if a.list.f:
            return [float(v) for v in a.list.f]
        if a.list.s:
            return [v.decode() for v in a.list.s]
        return []
    if kind == "tensor":
        return None  # decoded separately
    return None


def _parse_edge(s: str) -> Optional[Tuple[str, int]]:
    if s.startswith("^"):
        return None  # control dependency
    if ":" in s:
        name, idx = s.rsplit(":", 1)
        return name, int(idx)
    return s, 0


def parse_graphdef(data: bytes) -> GraphIR:
    gd = graph_pb2.GraphDef()
    gd.ParseFromString(data)
    nodes = []
    for nd in gd.node:
        inputs = [e for e in (_parse_edge(s) for s in nd.input) if e is not None]
        attrs = {k: _attr_to_py(v) for k, v in nd.attr.items()}
        value = None
        if nd.op == "Const":
            value = tensor_util.make_ndarray(nd.attr["value"].tensor)
        nodes.append(Node(nd.name, nd.op, inputs, attrs, value))
    return GraphIR(nodes)


def load_pb(path: str) -> GraphIR:
    with open(path, "rb") as f:
        return parse_graphdef(f.read())


# --------------------------------------------------------------------------------------
# Writer (synthetic frozen graphs)
# --------------------------------------------------------------------------------------
class GraphBuilder:
    """Emits NodeDefs the way TF1 python ops + ``convert_variables_to_constants`` do."""

    def __init__(self):
        self.gd = graph_pb2.GraphDef()
        self._names = set()

    def unique(self, name: str) -> str:
        if name not in self._names:
            self._names.add(name)
            return name
        i = 1
        while f"{name}_{i}" in self._names:
            i += 1
        self._names.add(f"{name}_{i}")
        return f"{name}_{i}"

    def _node(self, name, op, inputs=(), unique=True, **attrs):
        name = self.unique(name) if unique else name
        nd = self.gd.node.add()
        nd.name, nd.op = name, op
        nd.input.extend(inputs)
        for k, v in attrs.items():
            a = nd.attr[k]
            if k in ("T", "dtype", "Tidx", "out_type", "Index", "Tshape"):
                a.type = v
            elif isinstance(v, bool):
                a.b = v
            elif isinstance(v, int):
                a.i = v
            elif isinstance(v, float):
                a.f = v
            elif isinstance(v, (bytes, str)):
                a.s = v if isinstance(v, bytes) else v.encode()
            elif isinstance(v, (list, tuple)):
                a.list.i.extend(int(x) for x in v)
            else:
                raise TypeError((k, v))
        return name

    F = types_pb2.DT_FLOAT
    I = types_pb2.DT_INT32

    def placeholder(self, name, shape):
        nd = self.gd.node.add()
        nd.name, nd.op = name, "Placeholder"
        self._names.add(name)
        nd.attr["dtype"].type = self.F
        for s in shape:
            nd.attr["shape"].shape.dim.add().size = -1 if s is None else s
        return name

    def const(self, name, arr, splat=False):
        name = self.unique(name)
        arr = np.asarray(arr)
        nd = self.gd.node.add()
        nd.name, nd.op = name, "Const"
        dt = self.F if arr.dtype.kind == "f" else self.I
        nd.attr["dtype"].type = dt
        t = nd.attr["value"].tensor
        t.dtype = dt
        for s in arr.shape:
            t.tensor_shape.dim.add().size = s
        if splat or arr.ndim == 0:  # scalars / tf.constant(1.0, shape=[...]): one value (+ full shape)
            if dt == self.F:
                t.float_val.append(float(arr.flat[0]))
            else:
                t.int_val.append(int(arr.flat[0]))
        else:
            t.tensor_content = arr.astype(np.float32 if dt == self.F else np.int32).tobytes()
        return name

    def variable(self, name, arr):
        """A frozen tf.get_variable: Const + Identity '<name>/read'."""
        c = self.const(name, arr)
        return self._node(c + "/read", "Identity", [c], T=self.F)

    # thin op wrappers --------------------------------------------------------------
    def conv2d(self, name, x, w):
        return self._node(name, "Conv2D", [x, w], T=self.F, strides=[1, 1, 1, 1], padding="SAME",
                          data_format="NHWC", dilations=[1, 1, 1, 1], use_cudnn_on_gpu=True)

    def bias_add(self, name, x, b):
        return self._node(name, "BiasAdd", [x, b], T=self.F, data_format="NHWC")

    def relu(self, name, x):
        return self._node(name, "Relu", [x], T=self.F)

    def identity(self, name, x, unique=True):
        return self._node(name, "Identity", [x], unique=unique, T=self.F)

    def add(self, name, a, b):
        return self._node(name, "Add", [a, b], T=self.F)

    def pool(self, name, op, x):
        return self._node(name, op, [x], T=self.F, ksize=[1, 2, 2, 1], strides=[1, 2, 2, 1],
                          padding="SAME", data_format="NHWC")

    def shape(self, name, x):
        return self._node(name, "Shape", [x], T=self.F, out_type=self.I)

    def strided_slice_scalar(self, name, shp, i):
        b = self.const(name + "/stack", np.array([i], np.int32))
        e = self.const(name + "/stack_1", np.array([i + 1], np.int32))
        s = self.const(name + "/stack_2", np.array([1], np.int32))
        return self._node(name, "StridedSlice", [shp, b, e, s], T=self.I, Index=self.I, begin_mask=0,
                          end_mask=0, ellipsis_mask=0, new_axis_mask=0, shrink_axis_mask=1)

    def pack(self, name, xs):
        return self._node(name, "Pack", xs, T=self.I, N=len(xs), axis=0)

    def conv2d_transpose(self, name, out_shape, w, x, stride):
        return self._node(name, "Conv2DBackpropInput", [out_shape, w, x], T=self.F,
                          strides=[1, stride, stride, 1], padding="SAME", data_format="NHWC",
                          dilations=[1, 1, 1, 1], use_cudnn_on_gpu=True)

    def concat(self, name, xs, axis=3):
        ax = self.const(name + "/axis", np.array(axis, np.int32))
        return self._node(name, "ConcatV2", list(xs) + [ax], T=self.F, N=len(xs), Tidx=self.I)

    def softmax(self, name, x, unique=True):
        return self._node(name, "Softmax", [x], unique=unique, T=self.F)

    def sigmoid(self, name, x, unique=True):
        return self._node(name, "Sigmoid", [x], unique=unique, T=self.F)

    def split(self, name, x, num, axis=3):
        ax = self.const(name + "/split_dim", np.array(axis, np.int32))
        return self._node(name, "Split", [ax, x], T=self.F, num_split=num)

    def mul(self, name, a, b):
        return self._node(name, "Mul", [a, b], T=self.F)

    def add_n(self, name, xs):
        return self._node(name, "AddN", xs, T=self.F, N=len(xs))

    def serialize(self) -> bytes:
        return self.gd.SerializeToString(deterministic=True)


def build_aru_graphdef(graph: str = "ARU", scale_space_num: int = 5, num_scales_att: int = 3,
                       feat_root: int = 8, res_depth: int = 3, n_class: int = 2, channels: int = 1,
                       output: str = "softmax", seed: int = 0, logit_gain: float = 1.0,
                       logit_bias: Optional[Sequence[float]] = None) -> bytes:
    """Synthetic frozen ARU-/RU-Net following ARU_v1.py:62-294 node for node.

    Weights ~ N(0, sqrt(2/(kh*kw*cin+cout))), biases 0.1 (layers.py:223-239, :344-358).
    ``logit_gain`` / ``logit_bias`` rescale the 4x4 classifier so the synthetic separator
    mask is neither all-on nor all-off at the 0.05 threshold.
    """
    assert graph in ("RU", "ARU")
    rng = np.random.default_rng(seed)
    b = GraphBuilder()
    variables: Dict[str, str] = {}

    def var(scope, shape, kind):
        key = scope
        if key in variables:
            return variables[key]
        if kind == "w":
            std = np.sqrt(2.0 / (shape[0] * shape[1] * shape[2] + shape[3]))
            arr = rng.normal(0.0, std, size=shape).astype(np.float32)
        else:
            arr = np.full(shape, 0.1, np.float32)
        variables[key] = b.variable(scope, arr)
        return variables[key]

    def conv2d(x, scope, k, cin, cout, act):
        w = var(f"{scope}/weights", (k, k, cin, cout), "w")
        bi = var(f"{scope}/biases", (cout,), "b")
        y = b.conv2d(f"{scope}/conv", x, w)
        y = b.bias_add(f"{scope}/preActivation", y, bi)
        if act == "relu":
            y = b.relu(f"{scope}/activation", y)
        else:
            y = b.identity(f"{scope}/activation", y)
        return y

    def deconv2d(x, scope, cout, cin, like):
        w = var(f"{scope}/weights", (3, 3, cout, cin), "w")
        bi = var(f"{scope}/bias", (cout,), "b")
        shp = b.shape(scope.rsplit("/", 1)[0] + "/Shape", like)  # tf.shape(dw_h_conv), ARU_v1.py:256
        y = b.conv2d_transpose(f"{scope}/conv", shp, w, x, 2)
        y = b.bias_add(f"{scope}/preActivation", y, bi)
        return b.relu(f"{scope}/activation", y)

    def upsample_simple(x, scope, shape_out, up, ncls):
        ones = b.const(f"{scope}/Const", np.ones((up, up, ncls, ncls), np.float32), splat=True)
        return b.conv2d_transpose(f"{scope}/conv2d_transpose", shape_out, ones, x, up)

    def res_block(x, scope, cin, cout):
        x = conv2d(x, f"{scope}/conv1", 3, cin, cout, None)
        orig = x
        x = b.relu(f"{scope}/activation", x)
        for r in range(res_depth):
            x = conv2d(x, f"{scope}/convR_{r}", 3, cout, cout, "relu" if r < res_depth - 1 else None)
        x = b.add(f"{scope}/add", x, orig)
        return b.relu(f"{scope}/activation_1", x)

    def det_cnn(x, root):
        last, act = channels, feat_root
        skips = {}
        for layer in range(scale_space_num):
            scope = f"{root}/unet_down_{layer}"
            skips[layer] = res_block(x, scope, last, act)
            x = b.pool(f"{scope}/pool", "MaxPool", skips[layer]) if layer < scale_space_num - 1 else skips[layer]
            last, act = act, act * 2
        act = last // 2
        for layer in range(scale_space_num - 2, -1, -1):
            scope = f"{root}/unet_up_{layer}"
            de = deconv2d(x, f"{scope}/deconv", act, last, skips[layer])
            conc = b.concat(f"{scope}/concat", [skips[layer], de])
            x = res_block(conc, scope, 2 * act, act)
            last, act = act, act // 2
        return x

    def att_cnn(x, root):
        scope = f"{root}/attPart"
        c = conv2d(x, f"{scope}/conv1", 4, channels, 12, "relu")
        c = b.pool(f"{scope}/pool1", "MaxPool", c)
        c = conv2d(c, f"{scope}/conv2", 4, 12, 16, "relu")
        c = b.pool(f"{scope}/pool2", "MaxPool", c)
        c = conv2d(c, f"{scope}/conv3", 4, 16, 32, "relu")
        c = b.pool(f"{scope}/pool3", "MaxPool", c)
        return conv2d(c, f"{scope}/conv4", 4, 32, 1, "relu")

    x = b.placeholder("inImg", [None, None, None, channels])
    img_shape = b.shape("aru_net/misc/Shape", x)
    dims = [b.strided_slice_scalar(f"aru_net/misc/strided_slice", img_shape, i) for i in range(3)]
    froot = b.const("aru_net/misc/stack/3", np.array(feat_root, np.int32))
    o_shape = b.pack("aru_net/misc/stack", dims + [froot])

    use_att = graph == "ARU"
    scales = {0: x}
    att_maps = []
    if use_att:
        for sc in range(1, num_scales_att):
            scales[sc] = b.pool("aru_net/attMapG/avg_pool2d", "AvgPool", scales[sc - 1])
        up = 8
        for sc in range(num_scales_att):
            a = att_cnn(scales[sc], "aru_net/attMapG")
            shp = b.shape("aru_net/attMapG/Shape", x)
            att_maps.append(upsample_simple(a, "aru_net/attMapG/up", shp, up, 1))
            up *= 2
    det_maps = [det_cnn(x, "aru_net/featMapG")]
    if use_att:
        up = 1
        for sc in range(1, num_scales_att):
            d = det_cnn(scales[sc], "aru_net/featMapG")
            up *= 2
            det_maps.append(upsample_simple(d, "aru_net/featMapG/up", o_shape, up, feat_root))
        all_att = b.concat("aru_net/logit/concat", att_maps)
        sm = b.softmax("aru_net/logit/Softmax", all_att)
        parts = b.split("aru_net/logit/split", sm, num_scales_att)
        prods = [b.mul("aru_net/logit/Mul", det_maps[sc], parts if sc == 0 else f"{parts}:{sc}")
                 for sc in range(num_scales_att)]
        fmap = b.add_n("aru_net/logit/AddN", prods)
    else:
        fmap = det_maps[0]

    # classifier (ARU_v1.py:158-160): 4x4 conv, identity activation, then tf.identity 'logits'
    cls_scope = "aru_net/logit/class"
    std = np.sqrt(2.0 / (16 * feat_root + n_class))
    wc = rng.normal(0.0, std, size=(4, 4, feat_root, n_class)).astype(np.float32) * np.float32(logit_gain)
    bc = np.full((n_class,), 0.1, np.float32) if logit_bias is None else np.asarray(logit_bias, np.float32)
    w = b.variable(f"{cls_scope}/weights", wc)
    bi = b.variable(f"{cls_scope}/biases", bc)
    y = b.conv2d(f"{cls_scope}/conv", fmap, w)
    y = b.bias_add(f"{cls_scope}/preActivation", y, bi)
    y = b.identity(f"{cls_scope}/activation", y)
    logits = b.identity("aru_net/logit/logits", y)
    if output == "softmax":
        b.softmax("output", logits, unique=False)
    elif output == "sigmoid":
        b.sigmoid("output", logits, unique=False)
    else:
        b.identity("output", logits, unique=False)
    return b.serialize()


def aru_conv_macs(h: int, w: int, scale_space_num: int = 5, num_scales_att: int = 3, feat_root: int = 8,
                  res_depth: int = 3, n_class: int = 2, graph: str = "ARU") -> int:
    """Exact MAC count of one page (TF SAME/ceil semantics; transposed convs counted without
    zero insertion; ones-filter upsampling counted as data movement) - BASELINE.md section 2."""
    def cdiv(a, b):
        return -(-a // b)

    def det(hh, ww):
        macs, last, act = 0, 1, feat_root
        dims = []
        for layer in range(scale_space_num):
            dims.append((hh, ww))
            macs += hh * ww * 9 * (last * act + res_depth * act * act)
            if layer < scale_space_num - 1:
                hh, ww = cdiv(hh, 2), cdiv(ww, 2)
            last, act = act, act * 2
        act = last // 2
        for layer in range(scale_space_num - 2, -1, -1):
            macs += hh * ww * 9 * last * act  # deconv at input resolution
            hh, ww = dims[layer]
            macs += hh * ww * 9 * (2 * act * act + res_depth * act * act)
            last, act = act, act // 2
        return macs

    def att(hh, ww):
        macs = hh * ww * 16 * 1 * 12
        hh, ww = cdiv(hh, 2), cdiv(ww, 2)
        macs += hh * ww * 16 * 12 * 16
        hh, ww = cdiv(hh, 2), cdiv(ww, 2)
        macs += hh * ww * 16 * 16 * 32
        hh, ww = cdiv(hh, 2), cdiv(ww, 2)
        return macs + hh * ww * 16 * 32

    total, hh, ww = 0, h, w
    nsc = num_scales_att if graph == "ARU" else 1
    for sc in range(nsc):
        total += det(hh, ww)
        if graph == "ARU":
            total += att(hh, ww)
        hh, ww = cdiv(hh, 2), cdiv(ww, 2)
    return total + h * w * 16 * feat_root * n_class
