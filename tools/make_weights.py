"""One-time conversion: reference ONNX initialisers -> weights/*.afw (AIRFEW01 containers).

Run in the authoring container only (needs /root/reference/output/*.onnx):
    python -m tools.make_weights

The reference JIT-builds TensorRT engines from the ONNX files at start-up
(src/plnet.cpp:24-196, src/super_point.cpp:18-85, src/light_glue.cpp:24-118,
src/super_glue.cpp:26-130); this build replaces that with a plain weight
container that both the CUDA library (airslam_b200/csrc/weights.cc) and the
oracle (oracle/weights.py) read.  Nothing is folded that the graph has not
already folded.  Tensors with >= 2 dims are stored as fp16 (the operand type
of the tcgen05 kernels, same as the reference's kFP16 engines), everything else
fp32.  Layouts are the graph's own, except MatMul weights ([in,out]) which are
transposed to [out,in] so that every matrix is "output-major".

Container: 16-byte header {"AIRFEW01", u32 count, u32 0}, count x 160-byte
entries {char name[120]; u32 dtype (0 f32, 1 f16, 2 i32); u32 ndim; u32 dims[4];
u64 offset; u64 nbytes}, then 64-byte-aligned payloads.
"""
import os
import struct
import sys
import numpy as np

from . import onnx_reader as R

REF = "/root/reference/output"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")


def write_container(path, tensors):
    """tensors: list of (name, np.ndarray) with dtype float32 / float16 / int32."""
    count = len(tensors)
    table = 16 + 160 * count
    off = (table + 63) // 64 * 64
    entries, blobs = [], []
    for name, arr in tensors:
        arr = np.ascontiguousarray(arr)
        dt = {np.dtype(np.float32): 0, np.dtype(np.float16): 1, np.dtype(np.int32): 2}[arr.dtype]
        dims = list(arr.shape) if arr.ndim else [1]
        assert len(dims) <= 4 and len(name) < 120
        nb = arr.nbytes
        entries.append(struct.pack("<120sII4IQQ", name.encode(), dt, len(dims), *(dims + [1] * (4 - len(dims))), off, nb))
        blobs.append((off, arr.tobytes()))
        off = (off + nb + 63) // 64 * 64
    with open(path, "wb") as fh:
        fh.write(struct.pack("<8sII", b"AIRFEW01", count, 0))
        for e in entries:
            fh.write(e)
        for o, b in blobs:
            fh.seek(o)
            fh.write(b)
        fh.truncate(off)
    print("wrote", path, count, "tensors", off, "bytes")


def _store(arr):
    arr = np.asarray(arr)
    if arr.dtype == np.float32 and arr.ndim >= 2:
        assert np.abs(arr).max() < 6e4
        return arr.astype(np.float16)
    if arr.dtype in (np.float64,):
        return arr.astype(np.float32)
    if arr.dtype == np.int64:
        return arr.astype(np.int32)
    return arr


def _node_path(name):
    """'/backbone/stack1/conv1a/Conv' -> 'stack1.conv1a' ; '/a/ffn/ffn.0/MatMul' -> 'a.ffn.0'."""
    parts = [p for p in name.strip("/").split("/")[:-1]]
    out = []
    for i, p in enumerate(parts):
        if i + 1 < len(parts) and parts[i + 1].startswith(p + "."):
            continue  # 'ffn' followed by 'ffn.0', 'heads.0' followed by 'heads.0.0'
        out.append(p)
    return ".".join(out)


def convert_superpoint():
    g = R.load(os.path.join(REF, "superpoint_v1_sim_int32.onnx"))
    t = [("sp." + k, _store(v)) for k, v in g.init.items() if v.dtype == np.float32 and v.size > 16]
    write_container(os.path.join(OUT, "superpoint.afw"), t)


def convert_plnet():
    g = R.load(os.path.join(REF, "plnet_s0.onnx"))
    t = []
    for n in g.nodes:
        if n.op != "Conv":
            continue
        path = _node_path(n.name)
        path = path.replace("backbone.point_detector.", "pd.").replace("backbone.", "")
        t.append(("plnet." + path + ".weight", _store(g.init[n.inputs[1]])))
        if len(n.inputs) > 2:
            t.append(("plnet." + path + ".bias", _store(g.init[n.inputs[2]])))
    g1 = R.load(os.path.join(REF, "plnet_s1.onnx"))
    for n in g1.nodes:
        if n.op != "Gemm":
            continue
        assert n.attrs.get("transB", 0) == 1 and n.attrs.get("alpha", 1.0) == 1.0
        path = _node_path(n.name)
        t.append(("plnet.s1." + path + ".weight", _store(g1.init[n.inputs[1]])))
        if len(n.inputs) > 2:
            t.append(("plnet.s1." + path + ".bias", _store(g1.init[n.inputs[2]])))
    t.append(("plnet.s1.tspan", _store(g1.init["onnx::Mul_1141"].reshape(-1))))
    t.append(("plnet.s1.tspan_c", _store(g1.init["onnx::Mul_1142"].reshape(-1))))
    write_container(os.path.join(OUT, "plnet.afw"), t)


def convert_lightglue():
    g = R.load(os.path.join(REF, "superpoint_lightglue.onnx"))
    ident = {n.outputs[0]: n.inputs[0] for n in g.nodes if n.op == "Identity"}
    seen, t = set(), []
    for k, v in g.init.items():
        if k.startswith("transformers.") or k.startswith("log_assignment."):
            t.append(("lg." + k, _store(v)))
    for n in g.nodes:
        if n.op != "MatMul":
            continue
        w = ident.get(n.inputs[1], n.inputs[1])
        if w not in g.init or w in seen:
            continue
        seen.add(w)
        path = _node_path(n.name)
        t.append(("lg." + path + ".weight", _store(np.ascontiguousarray(g.init[w].T))))
    names = [x[0] for x in t]
    assert len(set(names)) == len(names)
    write_container(os.path.join(OUT, "lightglue.afw"), t)


def convert_superglue(kind):
    g = R.load(os.path.join(REF, "superglue_%s_sim_int32.onnx" % kind))
    t, seen = [], set()
    kenc_idx = [0, 3, 6, 9]
    kenc_n = 0
    layer = 0
    for n in g.nodes:
        if n.op != "Conv":
            continue
        w, b = n.inputs[1], n.inputs[2]
        if w in seen:
            continue
        seen.add(w)
        if not w[0].isdigit():
            base = w[: -len(".weight")]
            if base.startswith("gnn.layers.") and base.endswith("mlp.3"):
                layer = int(base.split(".")[2]) + 1
        elif g.init[w].shape[1] != 512:
            base = "kenc.encoder.%d" % kenc_idx[kenc_n]
            kenc_n += 1
        else:
            base = "gnn.layers.%d.mlp.0" % layer
        t.append(("sg." + base + ".weight", _store(g.init[w][:, :, 0])))
        t.append(("sg." + base + ".bias", _store(g.init[b])))
    t.append(("sg.bin_score", np.asarray(g.init["bin_score"], dtype=np.float32).reshape(1)))
    names = [x[0] for x in t]
    assert len(set(names)) == len(names), names
    write_container(os.path.join(OUT, "superglue_%s.afw" % kind), t)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["superpoint", "plnet", "lightglue", "superglue_indoor", "superglue_outdoor"]
    for w in which:
        if w == "superpoint":
            convert_superpoint()
        elif w == "plnet":
            convert_plnet()
        elif w == "lightglue":
            convert_lightglue()
        elif w.startswith("superglue_"):
            convert_superglue(w.split("_")[1])
