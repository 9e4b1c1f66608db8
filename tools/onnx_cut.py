"""Wire-level ONNX sub-graph cutter (test tooling; no `onnx` package in this image).

Purpose: an EXTERNAL pin for the oracle.  OpenCV's DNN module (cv2.dnn, 4.13, an independent ONNX executor that none of this
repo's code shares a line with) rejects the reference's five shipped graphs as a whole (dynamic H x W inputs, ConstantOfShape,
Einsum shapes).  It does accept static-shape sub-graphs, so this tool re-emits a sub-graph of a reference .onnx file:

  * nodes are copied as RAW BYTES from the source file (never re-encoded), in their original order, restricted to the
    ancestors of the requested output tensors and cut at the requested input tensors;
  * initializers are copied as raw bytes, only the ones the kept nodes reference;
  * graph inputs / outputs are rewritten as float tensors with the static shapes given by the caller.

Nothing on the product path imports this.  Used by tools/make_cv2dnn_golden.py.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import onnx_reader as R  # noqa: E402


def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(fn, payload):
    """length-delimited field"""
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + bytes(payload)


def _vi(fn, v):
    return _enc_varint((fn << 3) | 0) + _enc_varint(v)


def _value_info(name, shape, elem_type=1):
    dims = b"".join(_ld(1, _vi(1, d)) for d in shape)          # TensorShapeProto.dim{dim_value}
    tensor = _vi(1, elem_type) + _ld(2, dims)                     # TypeProto.Tensor{elem_type, shape}
    return _ld(1, name.encode()) + _ld(2, _ld(1, tensor))         # ValueInfoProto{name, type{tensor_type}}


def _reencode_fp16_rounded(raw):
    """TensorProto bytes -> the same tensor with fp32 weights rounded to fp16-representable values (rule of tools/make_golden.py:
    float32, ndim >= 2, more than 16 elements, not an exporter-folded 'onnx::Mul' constant), stored as raw fp32 data."""
    import numpy as np
    name, arr = R.parse_tensor(raw)
    if not (arr.dtype == np.float32 and arr.ndim >= 2 and arr.size > 16 and "onnx::Mul" not in name):
        return raw
    arr = arr.astype(np.float16).astype(np.float32)
    out = b"".join(_vi(1, d) for d in arr.shape) + _vi(2, 1) + _ld(8, name.encode()) + _ld(9, arr.tobytes())
    return out


def cut(src, dst, inputs, outputs, int_inputs=(), fp16_weights=False):
    """inputs: {tensor name: static shape}; outputs: {tensor name: static shape or None}.
    Returns the list of kept nodes."""
    with open(src, "rb") as fh:
        buf = memoryview(fh.read())
    model_other, graph_buf = [], None
    for fn, wt, v in R._fields(buf):
        if fn == 7:
            graph_buf = v
        elif wt == 2:
            model_other.append(_ld(fn, v))
        elif wt == 0:
            model_other.append(_vi(fn, v))
    raw_nodes, raw_init, graph_name = [], {}, b"cut"
    for gfn, gwt, gv in R._fields(graph_buf):
        if gfn == 1:
            raw_nodes.append((R.parse_node(gv), bytes(gv)))
        elif gfn == 5:
            name = ""
            for tfn, twt, tv in R._fields(gv):
                if tfn == 8:
                    name = bytes(tv).decode()
            raw_init[name] = bytes(gv)
    producer = {}
    for i, (n, _) in enumerate(raw_nodes):
        for o in n.outputs:
            producer[o] = i
    keep, stack = set(), [o for o in outputs]
    seen = set()
    while stack:
        t = stack.pop()
        if t in seen or t in inputs or t in raw_init or t == "":
            continue
        seen.add(t)
        if t not in producer:
            raise KeyError("tensor %r has no producer and is not a declared input" % t)
        i = producer[t]
        if i in keep:
            continue
        keep.add(i)
        stack.extend(raw_nodes[i][0].inputs)
    used_init = set()
    for i in keep:
        for t in raw_nodes[i][0].inputs:
            if t in raw_init and t not in inputs:
                used_init.add(t)
    g = bytearray()
    for i in sorted(keep):
        g += _ld(1, raw_nodes[i][1])
    g += _ld(2, graph_name)
    for t in sorted(used_init):
        g += _ld(5, _reencode_fp16_rounded(raw_init[t]) if fp16_weights else raw_init[t])
    for name, shape in inputs.items():
        g += _ld(11, _value_info(name, shape, 7 if name in int_inputs else 1))
    for name, shape in outputs.items():
        g += _ld(12, _value_info(name, shape or []))
    with open(dst, "wb") as fh:
        fh.write(b"".join(model_other) + _ld(7, bytes(g)))
    return [raw_nodes[i][0] for i in sorted(keep)]


if __name__ == "__main__":
    g = R.load(sys.argv[1])
    for n in g.nodes[: int(sys.argv[2]) if len(sys.argv) > 2 else 80]:
        print(n.op, n.name, list(n.inputs), "->", list(n.outputs))
