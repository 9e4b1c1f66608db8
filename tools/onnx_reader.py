"""Minimal ONNX protobuf wire-format reader (no `onnx` package in this image).

Test/tooling infrastructure only: used by tools/make_weights.py and
tools/onnx_interp.py to read the reference's shipped model files
(/root/reference/output/*.onnx), which are the arithmetic specification of the
hot path.  Nothing on the product path imports this.

Field numbers follow the public onnx.proto3 schema (ModelProto.graph = 7,
GraphProto.node = 1 / initializer = 5 / input = 11 / output = 12, ...).
"""
import struct
import numpy as np


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            break
        shift += 7
    return result, pos


def _fields(buf):
    """Yield (field_number, wire_type, value) over a serialized message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield fn, wt, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(buf):
    out = []
    pos = 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(_signed(v))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64,
           9: np.bool_, 10: np.float16, 11: np.float64}


def parse_tensor(buf):
    dims, dtype, raw, name = [], 1, None, ""
    float_data, int32_data, int64_data, double_data = [], [], [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            dims += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fn == 2:
            dtype = v
        elif fn == 4:
            if wt == 2:
                float_data += list(struct.unpack("<%df" % (len(v) // 4), v))
            else:
                float_data.append(struct.unpack("<f", v)[0])
        elif fn == 5:
            int32_data += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fn == 7:
            int64_data += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fn == 8:
            name = bytes(v).decode()
        elif fn == 9:
            raw = bytes(v)
        elif fn == 10:
            double_data += list(struct.unpack("<%dd" % (len(v) // 8), v))
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt).copy()
    elif float_data:
        arr = np.array(float_data, dtype=np_dt)
    elif int64_data:
        arr = np.array(int64_data, dtype=np_dt)
    elif int32_data:
        arr = np.array(int32_data, dtype=np_dt)
    elif double_data:
        arr = np.array(double_data, dtype=np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    arr = arr.reshape(dims) if dims else (arr.reshape(()) if arr.size == 1 else arr)
    return name, arr


def parse_attribute(buf):
    name, atype = "", 0
    f = i = s = t = None
    floats, ints, strings = [], [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            name = bytes(v).decode()
        elif fn == 2:
            f = struct.unpack("<f", v)[0]
        elif fn == 3:
            i = _signed(v)
        elif fn == 4:
            s = bytes(v)
        elif fn == 5:
            t = parse_tensor(v)[1]
        elif fn == 7:
            floats += list(struct.unpack("<%df" % (len(v) // 4), v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fn == 8:
            ints += _packed_varints(v) if wt == 2 else [_signed(v)]
        elif fn == 9:
            strings.append(bytes(v))
        elif fn == 20:
            atype = v
    # AttributeType: FLOAT=1 INT=2 STRING=3 TENSOR=4 FLOATS=6 INTS=7 STRINGS=8
    if atype == 1:
        val = f
    elif atype == 2:
        val = i
    elif atype == 3:
        val = s.decode()
    elif atype == 4:
        val = t
    elif atype == 6:
        val = floats
    elif atype == 7:
        val = ints
    elif atype == 8:
        val = [x.decode() for x in strings]
    else:  # untyped (old exporters): pick whatever is present
        val = t if t is not None else (ints or floats or i if i is not None else f)
    return name, val


class Node:
    __slots__ = ("op", "name", "inputs", "outputs", "attrs")

    def __init__(self):
        self.op, self.name, self.inputs, self.outputs, self.attrs = "", "", [], [], {}


def parse_node(buf):
    n = Node()
    for fn, wt, v in _fields(buf):
        if fn == 1:
            n.inputs.append(bytes(v).decode())
        elif fn == 2:
            n.outputs.append(bytes(v).decode())
        elif fn == 3:
            n.name = bytes(v).decode()
        elif fn == 4:
            n.op = bytes(v).decode()
        elif fn == 5:
            k, val = parse_attribute(v)
            n.attrs[k] = val
    return n


def _value_info_name(buf):
    for fn, wt, v in _fields(buf):
        if fn == 1:
            return bytes(v).decode()
    return ""


class Graph:
    def __init__(self):
        self.nodes, self.init, self.inputs, self.outputs = [], {}, [], []


def load(path):
    with open(path, "rb") as fh:
        buf = memoryview(fh.read())
    g = Graph()
    for fn, wt, v in _fields(buf):
        if fn == 7:  # ModelProto.graph
            for gfn, gwt, gv in _fields(v):
                if gfn == 1:
                    g.nodes.append(parse_node(gv))
                elif gfn == 5:
                    name, arr = parse_tensor(gv)
                    g.init[name] = arr
                elif gfn == 11:
                    g.inputs.append(_value_info_name(gv))
                elif gfn == 12:
                    g.outputs.append(_value_info_name(gv))
    g.inputs = [i for i in g.inputs if i not in g.init]
    return g


if __name__ == "__main__":
    import sys
    from collections import Counter
    g = load(sys.argv[1])
    print("inputs", g.inputs, "outputs", g.outputs, "nodes", len(g.nodes), "init", len(g.init))
    print(Counter(n.op for n in g.nodes))
