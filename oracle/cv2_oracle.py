"""CPU ORACLE B (test infrastructure): OpenCV's TensorFlow importer on the same GraphDef bytes.

``cv2.dnn.readNetFromTensorflow`` is an independent third-party executor of frozen TF
graphs; it pins Oracle A (``aru_oracle.py``) because the reference's own TF1 run of
net_post_processing_helper.py:56-72 is unavailable here.

OpenCV limitations worked around by *rewriting a copy* of the graph (never the product's view):
  * ``Conv2DBackpropInput`` needs a Const ``output_shape`` -> Shape/StridedSlice/Pack are folded
    for one (n,h,w);
  * broadcast ``Mul`` [1,H,W,8]x[1,H,W,1] (ARU_v1.py:152) fails to import -> the 1-channel
    operand is replaced by a ConcatV2 of 8 copies (same values).
"""
from __future__ import annotations

import numpy as np
from tensorboard.compat.proto import graph_pb2, types_pb2
from tensorboard.util import tensor_util

from .aru_oracle import Oracle


def fold_for_cv2(pb_bytes: bytes, h: int, w: int, n: int = 1) -> bytes:
    orc = Oracle(pb_bytes)
    orc.run(np.zeros((n, h, w, 1), np.float32))          # records every int-valued tensor
    ints = orc.int_values
    # channel counts per float tensor are needed to spot the broadcast Mul
    gd = graph_pb2.GraphDef()
    gd.ParseFromString(pb_bytes)
    out = graph_pb2.GraphDef()
    ir = orc.ir

    def n_channels(edge):
        name, idx = ir.resolve_identity(edge)
        nd = ir[name]
        if nd.op == "Split":
            return n_channels(nd.inputs[1]) // nd.attrs["num_split"]
        if nd.op in ("Conv2D",):
            return ir.const_value(nd.inputs[1]).shape[3]
        if nd.op == "Conv2DBackpropInput":
            return ir.const_value(nd.inputs[1]).shape[2]
        if nd.op == "ConcatV2":
            return sum(n_channels(e) for e in nd.inputs[:-1])
        if nd.op == "Placeholder":
            return 1
        return n_channels(nd.inputs[0])

    live = set(ir.topo_order([orc.out_name]))
    for nd in gd.node:
        if nd.name not in live:
            continue                                          # dead shape plumbing (e.g. o_shape in RU graphs)
        if nd.name in ints and nd.op in ("Shape", "StridedSlice", "Pack"):
            c = out.node.add()
            c.name, c.op = nd.name, "Const"
            c.attr["dtype"].type = types_pb2.DT_INT32
            c.attr["value"].tensor.CopyFrom(tensor_util.make_tensor_proto(np.asarray(ints[nd.name], np.int32)))
            continue
        if nd.op == "Mul":
            ca, cb = (n_channels(e) for e in ir[nd.name].inputs)
            if ca != cb:
                small = 1 if cb < ca else 0
                reps = max(ca, cb) // min(ca, cb)
                ax = out.node.add()
                ax.name, ax.op = nd.name + "/bc_axis", "Const"
                ax.attr["dtype"].type = types_pb2.DT_INT32
                ax.attr["value"].tensor.CopyFrom(tensor_util.make_tensor_proto(np.array(3, np.int32)))
                cc = out.node.add()
                cc.name, cc.op = nd.name + "/bc", "ConcatV2"
                cc.input.extend([nd.input[small]] * reps + [ax.name])
                cc.attr["N"].i = reps
                cc.attr["T"].type = types_pb2.DT_FLOAT
                cc.attr["Tidx"].type = types_pb2.DT_INT32
                m = out.node.add()
                m.CopyFrom(nd)
                m.input[small] = cc.name
                continue
        c = out.node.add()
        c.CopyFrom(nd)
        if nd.op == "Const" and ir[nd.name].value.ndim > 0 and not nd.attr["value"].tensor.tensor_content:
            # splat-encoded tf.constant(1.0, shape=[up,up,C,C]) (layers.py:717): OpenCV wants full content
            c.attr["value"].tensor.CopyFrom(tensor_util.make_tensor_proto(ir[nd.name].value))
        if nd.op == "Placeholder":
            del c.attr["shape"].shape.dim[:]
            for s in (n, h, w, 1):
                c.attr["shape"].shape.dim.add().size = s
    return out.SerializeToString()


def run_cv2(pb_bytes: bytes, image: np.ndarray, fetch: str = "output") -> np.ndarray:
    """image [H,W] float in [0,1] -> float32 [H,W,C] computed by cv2.dnn."""
    import cv2
    h, w = image.shape
    net = cv2.dnn.readNetFromTensorflow(np.frombuffer(fold_for_cv2(pb_bytes, h, w), np.uint8))
    net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
    net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)
    net.setInput(np.ascontiguousarray(image.astype(np.float32))[None, None])   # NCHW blob
    out = net.forward(fetch)
    return np.ascontiguousarray(out[0].transpose(1, 2, 0))
