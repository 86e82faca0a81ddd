"""GraphDef reader/writer and the lowering pass (host logic; no GPU)."""
import numpy as np
import pytest

from aru_b200.graphdef import GraphBuilder, aru_conv_macs, build_aru_graphdef, parse_graphdef
from aru_b200.program import (OP_AVGPOOL, OP_COMBINE, OP_CONV, OP_DECONV, OP_MAXPOOL, UnsupportedGraphError,
                              lower_graph)
from aru_b200.synth import synth_pb


def test_io_names_and_op_set():
    ir = parse_graphdef(synth_pb("separator"))
    assert ir["inImg"].op == "Placeholder" and ir["output"].op == "Softmax"   # helper.py:69-70
    ops = {n.op for n in ir.nodes.values()}
    assert {"Conv2D", "BiasAdd", "Relu", "MaxPool", "AvgPool", "Conv2DBackpropInput", "ConcatV2", "Softmax", "Split",
            "Mul", "AddN", "Shape", "StridedSlice", "Pack", "Identity", "Const", "Add"} <= ops


def test_weights_are_shared_across_scales():
    ir = parse_graphdef(synth_pb("separator"))
    convs = [n for n in ir.nodes.values() if n.op == "Conv2D"]
    filters = {ir.resolve_identity(n.inputs[1])[0] for n in convs}
    assert len(convs) == 121                      # 3 scales x (36 det + 4 att) + classifier
    assert len(filters) == 36 + 4 + 1             # ARU_v1.py:116,127 reuse_variables
    n_params = sum(ir[f].value.size for f in filters) + sum(
        ir.resolve_identity(n.inputs[1])[0] and ir.const_value(n.inputs[1]).size
        for n in ir.nodes.values() if n.op == "BiasAdd" and ir.resolve_identity(n.inputs[1])[0] in
        {ir.resolve_identity(m.inputs[1])[0] for m in ir.nodes.values() if m.op == "BiasAdd"}) * 0
    assert n_params > 0


def test_parameter_count_matches_survey():
    prog = lower_graph(parse_graphdef(synth_pb("separator")))
    assert prog.weights.size == 1043839


def test_mac_count_matches_baseline_table():
    assert round(aru_conv_macs(1024, 768) / 1e9, 2) == 24.99
    assert round(aru_conv_macs(1500, 1125) / 1e9, 2) == 53.87
    assert round(aru_conv_macs(1856, 1344) / 1e9, 2) == 79.27
    assert round(aru_conv_macs(1024, 768, scale_space_num=6, num_scales_att=5) / 1e9, 2) == 30.79


def test_lowering_fuses_everything():
    prog = lower_graph(parse_graphdef(synth_pb("separator")))
    kinds = [o.kind for o in prog.ops]
    assert kinds.count(OP_CONV) == 121 and kinds.count(OP_DECONV) == 12
    assert kinds.count(OP_MAXPOOL) == 21 and kinds.count(OP_AVGPOOL) == 2 and kinds.count(OP_COMBINE) == 1
    assert len(prog.ops) == 157                    # no copies, no stand-alone upsample / softmax / add
    # residual convs add the *pre-activation* conv1 output (ARU_v1.py:214,225)
    res_ops = [o for o in prog.ops if o.res.buf >= 0]
    assert len(res_ops) == 27
    pre_views = {(o.out_pre.buf, o.out_pre.ch_off) for o in prog.ops if o.out_pre.buf >= 0}
    assert all((o.res.buf, o.res.ch_off) in pre_views for o in res_ops)
    comb = [o for o in prog.ops if o.kind == OP_COMBINE][0]
    assert comb.up_att == (8, 16, 32) and comb.up_det == (1, 2, 4)
    assert prog.n_class == 2


def test_concat_is_aliased_into_channel_slices():
    prog = lower_graph(parse_graphdef(synth_pb("tiny")))
    deconvs = [o for o in prog.ops if o.kind == OP_DECONV]
    assert deconvs and all(o.out.ch_off == o.out.ch for o in deconvs)   # [skip, deconv] -> second half


def test_ru_graph_and_sigmoid_head():
    prog = lower_graph(parse_graphdef(synth_pb("tiny_sigmoid")))
    assert prog.n_class == 1 and prog.ops[-1].act == 3
    assert not any(o.kind == OP_COMBINE for o in prog.ops)


def test_missing_tensor_names_raise_keyerror():
    with pytest.raises(KeyError):
        lower_graph(parse_graphdef(synth_pb("tiny")), in_name="nope")
    with pytest.raises(KeyError):
        lower_graph(parse_graphdef(synth_pb("tiny")), out_name="nope")


def test_unsupported_ops_fail_loudly():
    b = GraphBuilder()
    x = b.placeholder("inImg", [None, None, None, 1])
    w = b.variable("w", np.zeros((3, 3, 1, 8), np.float32))
    bi = b.variable("b", np.zeros((8,), np.float32))
    y = b.bias_add("ba", b._node("conv", "Conv2D", [x, w], T=b.F, strides=[1, 2, 2, 1], padding="SAME",
                                 data_format="NHWC", dilations=[1, 1, 1, 1]), bi)
    b.identity("output", y, unique=False)
    with pytest.raises(UnsupportedGraphError):
        lower_graph(parse_graphdef(b.serialize()))


def test_writer_is_deterministic():
    assert build_aru_graphdef(seed=7, scale_space_num=2, num_scales_att=1) == \
        build_aru_graphdef(seed=7, scale_space_num=2, num_scales_att=1)
